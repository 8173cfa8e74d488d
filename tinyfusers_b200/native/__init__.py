# reference: tinyfusers/native/__init__.py:1-3 exports the ctypes singletons; here the single B200 one.
from .b200.ops import b200  # noqa: F401
