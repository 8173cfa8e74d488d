# reference: tinyfusers/native/__init__.py:1-3 exports the ctypes singletons `cuda, cudart, nvrtc, cublas`; they stay importable
# (lazy: nothing is loaded until one of their entry points is called) next to the one the B200 path uses, `b200`.
from .b200.ops import b200  # noqa: F401
from .cuda.ops import cuda, cudart  # noqa: F401
from .nvrtc.ops import nvrtc  # noqa: F401
from .cublas.ops import cublas  # noqa: F401
