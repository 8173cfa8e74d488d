"""`cuda` / `cudart` singletons (reference: tinyfusers/native/cuda/ops.py:3-98), lazy: see native/_lazy.py."""
import ctypes

from .._lazy import LazyLibrary


class struct_CUmod_st(ctypes.Structure):
    pass


class struct_CUctx_st(ctypes.Structure):
    pass


class struct_cuFunction(ctypes.Structure):
    pass


class struct_CUstream_st(ctypes.Structure):
    pass


class Cuda(LazyLibrary):
    _sonames = ("libcuda.so.1", "libcuda.so")
    _methods = {
        "cuInit": ("cuInit", (), ()),
        "cuCtxCreate_v2": ("cuCtxCreate_v2", (), ()),
        "cuModuleLoadData": ("cuModuleLoadData", (0,), ()),
        "cuModuleGetFunction": ("cuModuleGetFunction", (0,), ()),
        "cuLaunchKernel": ("cuLaunchKernel", (), ()),
    }


class Cudart(LazyLibrary):
    _sonames = ("libcudart.so.12", "libcudart.so")
    _methods = {
        "cudaMalloc": ("cudaMalloc", (0,), ()),
        "cudaMemcpy": ("cudaMemcpy", (), ()),
        "cudaFree": ("cudaFree", (), ()),
        "cudaDeviceGetAttribute": ("cudaDeviceGetAttribute", (), ()),
    }


cuda = Cuda()
cudart = Cudart()

cuda.CUmodule = ctypes.POINTER(struct_CUmod_st)
cuda.CUcontext = ctypes.POINTER(struct_CUctx_st)
cuda.CUfunction = ctypes.POINTER(struct_cuFunction)
cuda.CUstream = ctypes.POINTER(struct_CUstream_st)
cuda.handler = ctypes.c_uint32
cuda.CU_CTX_SCHED_AUTO = 0

cudart.CUDA_SUCCESS = 0
cudart.cudaMemcpyHostToDevice = 1
cudart.cudaMemcpyDeviceToHost = 2
cudart.cudaDevAttrComputeCapabilityMajor = 75
cudart.cudaDevAttrComputeCapabilityMinor = 76
