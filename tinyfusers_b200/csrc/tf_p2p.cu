// tinyfusers_b200 — the cond / uncond halves of classifier-free guidance on TWO GPUs (one image, lower latency).
//
// Reference: StableDiffusion.get_model_output (tinyfusers/variants/sd.py:27-46) evaluates the UNet on the batch
// [uncond ; cond] and combines e_t = u + g (c - u); the two halves are independent until that line. Here GPU 0 runs the UNet
// on the unconditional half and GPU 1 on the conditional half (batch 1 each), and ONE kernel per GPU does the exchange and
// the update: it stores its own noise prediction (4 x H x W fp32 = 64 KiB at 512^2) straight into the peer's mailbox over
// NVLink (peer-mapped memory, plain st.global), publishes a per-block sequence flag, waits for the peer's flag, and applies
// CFG + DDIM to its own copy of the latent. Both GPUs evaluate the same expression on the same fp32 operands, so the two
// latents stay bit-identical without any further exchange. No NCCL launch and no host involvement on the step path: the
// kernel is captured in the sampler's CUDA graph like every other launch.
//
// Mailbox: 2 slots (step parity) x C*HW floats; flags: 2 slots x gridDim.x ints, value = sequence number + 1. A rank can run
// at most one step ahead of its peer (its next UNet needs the latent this exchange produces), so two slots never collide.
// Memory comes from cudaMalloc (not torch's caching allocator: CUDA IPC handles name whole allocations) and is mapped into the
// peer process with cudaIpcOpenMemHandle; the 64-byte handles travel through torch.distributed once, at set-up.
#include <string.h>

#include "tf_common.cuh"
#include "tinyfusers_b200.h"

namespace {

constexpr int kP2PThreads = 256;
constexpr int kP2PElemsPerThread = 4;
#ifndef TF_P2P_TIMEOUT_CYCLES
// the peer may be busy with set-up (weight packing, graph capture) the first time: wait long, but not forever
#define TF_P2P_TIMEOUT_CYCLES 120000000000ll   // ~60 s
#endif

__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_cv(const float* p) {   // do not trust a stale L1 line: the peer wrote this over NVLink
  float v;
  asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

// mode bits: 1 = send own half to the peer, 2 = wait for the peer's half (both set in production; the single-GPU test drives
// the two ranks one after the other: rank 1 send-only, then rank 0 complete)
__global__ void __launch_bounds__(kP2PThreads)
cfg_ddim_split_kernel(const float* __restrict__ eps, int eps_stride, const float* __restrict__ latent,
                      float* __restrict__ latent_out, float* __restrict__ e_t_out, const float* __restrict__ a_t_tab,
                      const float* __restrict__ a_prev_tab, const int* __restrict__ idx_dev, float guidance, int C, int HW,
                      int rank, const float* my_mailbox, float* peer_mailbox, const int* my_flags, int* peer_flags,
                      const int* __restrict__ seq_dev, int mode) {
  tf::pdl_prologue();
  const int total = C * HW;
  const int seq = *seq_dev;
  const int slot = seq & 1;
  const int nblk = gridDim.x;
  const int base = (blockIdx.x * kP2PThreads + threadIdx.x) * kP2PElemsPerThread;
  float own[kP2PElemsPerThread];
#pragma unroll
  for (int j = 0; j < kP2PElemsPerThread; ++j) {
    const int i = base + j;
    own[j] = 0.f;
    if (i < total) {
      const int p = i % HW, c = i / HW;
      own[j] = eps[(size_t)p * eps_stride + c];
    }
  }
  if (mode & 1) {
    float* dst = peer_mailbox + (size_t)slot * total;
#pragma unroll
    for (int j = 0; j < kP2PElemsPerThread; ++j)
      if (base + j < total) dst[base + j] = own[j];
    __threadfence_system();          // every thread's stores are ordered before the flag below (after the block barrier)
    __syncthreads();
    if (threadIdx.x == 0) st_release_sys(peer_flags + slot * nblk + blockIdx.x, seq + 1);
  }
  if (mode & 2) {
    if (threadIdx.x == 0) {
      const int* f = my_flags + slot * nblk + blockIdx.x;
      const long long t0 = clock64();
      while (ld_acquire_sys(f) != seq + 1) {
        if (clock64() - t0 > TF_P2P_TIMEOUT_CYCLES) __trap();   // peer gone: surface as a launch failure, do not hang the GPU
      }
    }
    __syncthreads();
    const float* src = my_mailbox + (size_t)slot * total;
    const int idx = idx_dev ? *idx_dev : 0;
    const float a_t = a_t_tab[idx], a_prev = a_prev_tab[idx];
    const float sqrt_one_minus_at = sqrtf(1.f - a_t);
    const float inv_sqrt_at = 1.f / sqrtf(a_t);
    const float sqrt_aprev = sqrtf(a_prev);
    const float dir_coef = sqrtf(1.f - a_prev);
#pragma unroll
    for (int j = 0; j < kP2PElemsPerThread; ++j) {
      const int i = base + j;
      if (i < total) {
        const float other = ld_cv(src + i);
        const float u = rank == 0 ? own[j] : other;      // rank 0 holds the unconditional half ([uncond ; cond], sd.py:32)
        const float cnd = rank == 0 ? other : own[j];
        const float e = u + guidance * (cnd - u);         // the same expression, operand order and rounding on both GPUs
        const float x = latent[i];
        const float pred_x0 = (x - sqrt_one_minus_at * e) * inv_sqrt_at;
        latent_out[i] = sqrt_aprev * pred_x0 + dir_coef * e;
        if (e_t_out) e_t_out[i] = e;
      }
    }
  }
}

}  // namespace

extern "C" int tf_p2p_blocks(int C, int HW) {
  const long total = (long)C * HW;
  return (int)((total + kP2PThreads * kP2PElemsPerThread - 1) / (kP2PThreads * kP2PElemsPerThread));
}

extern "C" int tf_p2p_alloc(size_t bytes, void** dev_ptr, void* handle64) {
  TF_CHECK_ARG(bytes > 0 && dev_ptr && handle64, "tf_p2p_alloc: bad arguments");
  void* p = nullptr;
  TF_CUDA(cudaMalloc(&p, bytes));
  TF_CUDA(cudaMemset(p, 0, bytes));
  TF_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    tf_set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    return (int)e;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  memcpy(handle64, &h, 64);
  *dev_ptr = p;
  return TF_OK;
}

extern "C" int tf_p2p_open(const void* handle64, void** dev_ptr) {
  TF_CHECK_ARG(handle64 && dev_ptr, "tf_p2p_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  TF_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *dev_ptr = p;
  return TF_OK;
}

extern "C" int tf_p2p_close(void* dev_ptr) {
  if (dev_ptr) TF_CUDA(cudaIpcCloseMemHandle(dev_ptr));
  return TF_OK;
}

extern "C" int tf_p2p_free(void* dev_ptr) {
  if (dev_ptr) TF_CUDA(cudaFree(dev_ptr));
  return TF_OK;
}

extern "C" int tf_cfg_ddim_step_split_f32(const float* eps_nhwc, int eps_pixel_stride, const float* latent, float* latent_out,
                                          float* e_t_out, const float* alphas_dev, const float* alphas_prev_dev,
                                          const int* index_dev, float guidance, int C, int HW, int rank,
                                          const void* my_mailbox, void* peer_mailbox, const void* my_flags, void* peer_flags,
                                          const int* seq_dev, int mode, void* stream) {
  TF_CHECK_ARG(eps_nhwc && latent && latent_out && alphas_dev && alphas_prev_dev && seq_dev,
               "tf_cfg_ddim_step_split_f32: null pointer");
  TF_CHECK_ARG(C > 0 && HW > 0 && eps_pixel_stride >= C && (rank == 0 || rank == 1) && mode >= 1 && mode <= 3,
               "tf_cfg_ddim_step_split_f32: bad arguments");
  TF_CHECK_ARG((!(mode & 1) || (peer_mailbox && peer_flags)) && (!(mode & 2) || (my_mailbox && my_flags)),
               "tf_cfg_ddim_step_split_f32: mailbox / flag pointers missing for mode %d", mode);
  const int blocks = tf_p2p_blocks(C, HW);
  TF_LAUNCH(cfg_ddim_split_kernel, blocks, kP2PThreads, 0, (cudaStream_t)stream, eps_nhwc, eps_pixel_stride, latent, latent_out,
            e_t_out, alphas_dev, alphas_prev_dev, index_dev, guidance, C, HW, rank, reinterpret_cast<const float*>(my_mailbox),
            reinterpret_cast<float*>(peer_mailbox), reinterpret_cast<const int*>(my_flags), reinterpret_cast<int*>(peer_flags),
            seq_dev, mode);
  TF_LAUNCH_CHECK();
  tf_launch_count_add(1);
  return TF_OK;
}
