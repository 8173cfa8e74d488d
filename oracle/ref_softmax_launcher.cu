// Test infrastructure (oracle/): host launcher around the REFERENCE's own softmax kernel.
//
// The one source file of the reference's hot path that compiles on its own is tinyfusers/native/cuda/softmax.cu
// (softmax_kernel, :24-112; warpReduceMax / warpReduceSum from utils.h:13-26). oracle/build_ref.py compiles it FROM WHERE IT
// LIES under /root/reference (REF_SOFTMAX_CU is that path, given on the nvcc command line; nothing is copied into this
// repository) together with this launcher into oracle/_ref/libref_softmax.so, with the options the reference gives its
// RawModule (attention/sdpa.py:13: --use_fast_math -D__CUDA_NO_HALF_CONVERSIONS__ -I<native/cuda>).
// The launch geometry is the reference's (attention/sdpa.py:59-61,72-73): one 256-thread block per row,
// 2 * 256 / 32 floats of dynamic shared memory. Only tests/ load the library; the product never does.
#include REF_SOFTMAX_CU

extern "C" int ref_softmax_forward(float* out, const float* inp, int N, int C, void* stream) {
  const int block = 256;
  const size_t smem = 2 * block / 32 * sizeof(float);
  softmax_kernel<<<N, block, smem, (cudaStream_t)stream>>>(out, inp, N, C);
  return (int)cudaGetLastError();
}
