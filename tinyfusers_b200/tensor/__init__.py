"""`tinyfusers.tensor` — the package name the reference's model files import `Tensor` from
(`from ..tensor.tensor import Tensor`, e.g. vision/resnet.py:4); the class lives in `storage/tensor.py`."""
