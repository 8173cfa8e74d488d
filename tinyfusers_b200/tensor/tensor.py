"""Alias of `tinyfusers_b200.storage.tensor` under the import path the reference's model files use
(reference: `from ..tensor.tensor import Tensor` in vision/resnet.py:4, attention/attention.py:8, vae/decoder.py:6)."""
from ..storage.tensor import Tensor  # noqa: F401
