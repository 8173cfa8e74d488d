"""Lazy ctypes bindings over the NVIDIA driver-side libraries, kept for API compatibility with the reference's
`tinyfusers.native` singletons (`cuda, cudart, nvrtc, cublas`: native/__init__.py:1-3, native/*/ops.py).

None of them is on the B200 hot path - that is `tinyfusers_b200.native.b200` (libtinyfusers_b200.so) - so nothing is loaded at
import: a library is opened on the first call of one of its entry points (sonames tried in order) and a missing library
raises RuntimeError there. Calling convention as in the reference: every method returns the library's integer status; the
arguments listed in `byref` are passed by reference, floats in `fscalars` as pointers to a c_float (cuBLAS alpha / beta)."""
import ctypes


class LazyLibrary:
    _sonames = ()
    # method name -> (exported symbol, indices of arguments passed by reference, indices of float scalars passed by pointer)
    _methods = {}

    def __init__(self):
        self._dll = None

    @property
    def dll(self):
        if self._dll is None:
            errors = []
            for name in self._sonames:
                try:
                    self._dll = ctypes.CDLL(name)
                    break
                except OSError as exc:
                    errors.append(str(exc))
            if self._dll is None:
                raise RuntimeError(f"{type(self).__name__}: none of {self._sonames} could be loaded: {errors[-1] if errors else ''}")
        return self._dll

    def __getattr__(self, name):
        spec = type(self)._methods.get(name)
        if spec is None:
            raise AttributeError(name)
        symbol, byref, fscalars = spec

        def call(*args):
            fn = getattr(self.dll, symbol)
            fn.restype = ctypes.c_int
            conv = []
            for i, a in enumerate(args):
                if i in byref:
                    a = ctypes.byref(a)
                elif i in fscalars:
                    a = ctypes.byref(ctypes.c_float(a))
                elif isinstance(a, str):
                    a = a.encode()
                conv.append(a)
            return fn(*conv)
        call.__name__ = name
        return call
