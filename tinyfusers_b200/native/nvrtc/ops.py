"""`nvrtc` singleton (reference: tinyfusers/native/nvrtc/ops.py:3-69), lazy: see native/_lazy.py. The B200 kernels are compiled
ahead of time for sm_100a (tinyfusers_b200/csrc/build.py); nothing on the path compiles at run time."""
import ctypes

from .._lazy import LazyLibrary


class _nvrtcProgram(ctypes.Structure):
    pass


class Nvrtc(LazyLibrary):
    _sonames = ("libnvrtc.so.12", "libnvrtc.so")
    _methods = {
        "nvrtcGetCUBINSize": ("nvrtcGetCUBINSize", (), ()),
        "nvrtcGetCUBIN": ("nvrtcGetCUBIN", (), ()),
        "nvrtcGetPTXSize": ("nvrtcGetPTXSize", (), ()),
        "nvrtcGetPTX": ("nvrtcGetPTX", (), ()),
    }

    def nvrtcCreateProgram(self, prog, code_str):
        fn = self.dll.nvrtcCreateProgram
        fn.restype = ctypes.c_uint32
        return fn(ctypes.byref(prog), code_str.encode(), b"<null>", 0, None, None)

    def nvrtcCompileProgram(self, prog, compile_options):
        fn = self.dll.nvrtcCompileProgram
        fn.restype = ctypes.c_uint32
        opts = (ctypes.c_char_p * len(compile_options))(*[o if isinstance(o, bytes) else o.encode() for o in compile_options])
        return fn(prog, len(compile_options), opts)


nvrtc = Nvrtc()
for _i, _n in enumerate(("SUCCESS", "ERROR_OUT_OF_MEMORY", "ERROR_PROGRAM_CREATION_FAILURE", "ERROR_INVALID_INPUT",
                         "ERROR_INVALID_PROGRAM", "ERROR_INVALID_OPTION", "ERROR_COMPILATION", "ERROR_BUILTIN_OPERATION_FAILURE",
                         "ERROR_NO_NAME_EXPRESSIONS_AFTER_COMPILATION", "ERROR_NO_LOWERED_NAMES_BEFORE_COMPILATION",
                         "ERROR_NAME_EXPRESSION_NOT_VALID", "ERROR_INTERNAL_ERROR", "ERROR_TIME_FILE_WRITE_FAILED")):
    setattr(nvrtc, "NVRTC_" + _n, _i)
nvrtc.nvrtcProgram = ctypes.POINTER(_nvrtcProgram)
nvrtc.nvrtcResult = ctypes.c_uint32
