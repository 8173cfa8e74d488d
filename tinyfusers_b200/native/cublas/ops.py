"""`cublas` singleton (reference: tinyfusers/native/cublas/ops.py:3-70), lazy: see native/_lazy.py. The B200 path never calls
cuBLAS (north_star: no cuDNN / cuBLAS on the hot path); this exists so code written against `tinyfusers.native` still imports."""
import ctypes

from .._lazy import LazyLibrary


class Cublas(LazyLibrary):
    _sonames = ("libcublas.so.12", "libcublas.so")
    _methods = {
        "cublasCreate": ("cublasCreate_v2", (0,), ()),
        "cublasDestroy": ("cublasDestroy_v2", (), ()),
        # (handle, transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc[, batchCount]): alpha / beta by host pointer
        "cublasSgemm": ("cublasSgemm_v2", (), (6, 11)),
        "cublasSgemmBatched": ("cublasSgemmBatched", (), (6, 11)),
    }


cublas = Cublas()
for _i, _n in enumerate(("SUCCESS", "NOT_INITIALIZED", "ALLOC_FAILED", "INVALID_VALUE", "ARCH_MISMATCH", "MAPPING_ERROR",
                         "EXECUTION_FAILED", "INTERNAL_ERROR", "NOT_SUPPORTED", "LICENSE_ERROR")):
    setattr(cublas, "CUBLAS_STATUS_" + _n, _i)
cublas.CUBLAS_OP_N, cublas.CUBLAS_OP_T, cublas.CUBLAS_OP_C = 0, 1, 2
cublas.cublasHandle_t = ctypes.c_void_p
