// tinyfusers_b200 — the small HBM/launch-bound pieces of the UNet step:
//   timestep embedding, the M=1 time-MLP / ResBlock emb projections (GEMV), the Cin=4 input conv,
//   nearest x2 upsample, NCHW<->NHWC edge conversions and the fused CFG-combine + DDIM update.
#include "tf_common.cuh"
#include "tinyfusers_b200.h"

namespace {

// reference: tinyfusers/vision/unet.py:92-97. Angles are formed in fp64 like the reference
// (int64 timestep * float promotes to float64 in CuPy), the result is fp32.
__global__ void timestep_embedding_kernel(const float* __restrict__ t_dev, const int* __restrict__ idx_dev,
                                          int dim, float max_period, float* __restrict__ out) {
  tf::pdl_prologue();  // PDL: let the next kernel start launching, then wait for our producers
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = dim / 2;
  if (i >= half) return;
  const double t = (double)t_dev[idx_dev ? *idx_dev : 0];
  const double f = exp(-log((double)max_period) * (double)i / (double)half);
  const double a = t * f;
  out[i] = (float)cos(a);
  out[half + i] = (float)sin(a);
}

// out[n] = sum_k act(x[k]) * W[n,k] + bias[n] (+ bias2[n]); one warp per output row, fp16 weights.
// reference: Linear at M=1 — UNet time_embed (vision/unet.py:11,54) and ResBlock.emb_layers
// (vision/resnet.py:13-16,27). All 22 ResBlock projections are served by ONE launch over the
// row-concatenated weight matrix.
__global__ void gemv_kernel(const float* __restrict__ x, const __half* __restrict__ W, const float* __restrict__ bias,
                            const float* __restrict__ bias2, float* __restrict__ out, int N, int K, int silu_in) {
  tf::pdl_prologue();  // PDL: let the next kernel start launching, then wait for our producers
  extern __shared__ float xs[];
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float v = x[k];
    xs[k] = silu_in ? tf::silu_f(v) : v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + warp;
  if (n >= N) return;
  const __half* w = W + (size_t)n * K;
  float acc = 0.f;
  for (int k = lane * 8; k < K; k += 32 * 8) {
    tf::Pack16 pk;
    pk.v = __ldg(reinterpret_cast<const uint4*>(w + k));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = __half22float2(pk.h2[j]);
      acc += f.x * xs[k + 2 * j] + f.y * xs[k + 2 * j + 1];
    }
  }
  acc = tf::warp_sum(acc);
  if (lane == 0) out[n] = acc + (bias ? bias[n] : 0.f) + (bias2 ? bias2[n] : 0.f);
}

// 3x3 pad-1 conv with a tiny input-channel count, fp32 NCHW in -> fp16 NHWC out.
// reference: UNetModel.input_blocks[0] = Conv2d(4, 320, 3x3, pad 1) (vision/unet.py:13).
// w: fp32 (Cout, Cin, 3, 3) exactly as the reference stores it. Block = 32 pixels x 8 warps: lane = pixel (its
// 36 inputs stay in registers), warp = an eighth of the output-channel groups, so weight reads are
// warp-uniform shared-memory broadcasts.
template <int CIN>
__global__ void __launch_bounds__(256)
conv3x3_smallcin_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                        __half* __restrict__ out, int NI, int H, int W_, int Cout, int out_stride, int x_images,
                        float2* __restrict__ gn_stats, int gn_unit) {
  tf::pdl_trigger();
  extern __shared__ float ws[];  // [Cout][CIN*9] then [Cout] bias, then [Cout] float2 channel sums (statistics)
  float* bs = ws + Cout * CIN * 9;
  float2* csum = reinterpret_cast<float2*>(bs + Cout);
  // weights do not depend on the producer kernels: stage them (128-bit loads) before the dependency wait
  {
    const int n4 = (Cout * CIN * 9) >> 2;   // Cout % 8 == 0 -> a whole number of float4
    const float4* w4 = reinterpret_cast<const float4*>(w);
    float4* ws4 = reinterpret_cast<float4*>(ws);
    for (int i = threadIdx.x; i < n4; i += blockDim.x) ws4[i] = __ldg(w4 + i);
    for (int i = threadIdx.x; i < Cout; i += blockDim.x) bs[i] = bias ? bias[i] : 0.f;
  }
  tf::pdl_wait();
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long npix = (long)NI * H * W_;
  const long pix = (long)blockIdx.x * 32 + lane;
  if (pix >= npix && gn_stats == nullptr) return;   // with statistics H*W % 32 == 0 (host-checked): no partial block
  const int xo = (int)(pix % W_);
  const int yo = (int)((pix / W_) % H);
  const int n = (int)(pix / ((long)W_ * H)) % x_images;
  float in[CIN * 9];
#pragma unroll
  for (int c = 0; c < CIN; ++c)
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int yy = yo + r - 1, xx = xo + s - 1;
        in[c * 9 + r * 3 + s] =
            (yy >= 0 && yy < H && xx >= 0 && xx < W_) ? x[(((size_t)n * CIN + c) * H + yy) * W_ + xx] : 0.f;
      }
  const int groups = Cout / 8;
  for (int g = warp; g < groups; g += 8) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float* wr = ws + (g * 8 + j) * CIN * 9;
      float a = bs[g * 8 + j];
#pragma unroll
      for (int k = 0; k < CIN * 9; ++k) a += in[k] * wr[k];
      acc[j] = a;
    }
    tf::Pack16 pk;
#pragma unroll
    for (int j = 0; j < 4; ++j) pk.h2[j] = __floats2half2_rn(acc[2 * j], acc[2 * j + 1]);
    *reinterpret_cast<uint4*>(out + (size_t)pix * out_stride + g * 8) = pk.v;
    if (gn_stats) {
      // GroupNorm statistics of the OUTPUT, same layout as the GEMM epilogue's: this block's 32 pixels are one slot;
      // per channel {sum, sumsq} of the rounded fp16 values over the 32 lanes (fixed xor-shuffle order)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(pk.h2[j]);
        const float s0 = tf::warp_sum(f.x), q0 = tf::warp_sum(f.x * f.x);
        const float s1 = tf::warp_sum(f.y), q1 = tf::warp_sum(f.y * f.y);
        if (lane == 0) {
          csum[g * 8 + 2 * j] = make_float2(s0, q0);
          csum[g * 8 + 2 * j + 1] = make_float2(s1, q1);
        }
      }
    }
  }
  if (gn_stats) {
    __syncthreads();
    const int units = Cout / gn_unit;
    const long slot_global = blockIdx.x;   // (image, slot) flattened: H*W % 32 == 0
    for (int u = threadIdx.x; u < units; u += blockDim.x) {
      float a = 0.f, b = 0.f;
      for (int k = 0; k < gn_unit; ++k) { a += csum[u * gn_unit + k].x; b += csum[u * gn_unit + k].y; }
      gn_stats[slot_global * units + u] = make_float2(a, b);
    }
  }
}

// nearest-neighbour x2 (reference: Upsample.__call__ broadcast+reshape, vision/unet.py:81-83)
__global__ void upsample2x_kernel(const __half* __restrict__ x, int x_stride, __half* __restrict__ out,
                                  int out_stride, int NI, int H, int W_, int C) {
  tf::pdl_prologue();  // PDL: let the next kernel start launching, then wait for our producers
  const int nvec = C / 8;
  const long total = (long)NI * 2 * H * 2 * W_ * nvec;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int v = (int)(idx % nvec);
    const long pix = idx / nvec;
    const int xo = (int)(pix % (2 * W_));
    const int yo = (int)((pix / (2 * W_)) % (2 * H));
    const int n = (int)(pix / ((long)4 * W_ * H));
    const size_t src = (((size_t)n * H + (yo >> 1)) * W_ + (xo >> 1)) * x_stride + v * 8;
    *reinterpret_cast<uint4*>(out + (size_t)pix * out_stride + v * 8) = *reinterpret_cast<const uint4*>(x + src);
  }
}

// fp32/fp16 NCHW -> fp16 NHWC (API edge: per-op wrappers keep the reference's NCHW signatures)
template <typename TIn>
__global__ void nchw_to_nhwc_kernel(const TIn* __restrict__ x, __half* __restrict__ out, int C, int HW,
                                    int out_stride) {
  tf::pdl_prologue();  // PDL: let the next kernel start launching, then wait for our producers
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? (float)x[((size_t)n * C + c) * HW + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    if (p < HW && c < C) out[((size_t)n * HW + p) * out_stride + c] = __float2half_rn(tile[threadIdx.x][i]);
  }
}

template <typename TOut>
__global__ void nhwc_to_nchw_kernel(const __half* __restrict__ x, int x_stride, TOut* __restrict__ out, int C,
                                    int HW) {
  tf::pdl_prologue();  // PDL: let the next kernel start launching, then wait for our producers
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (p < HW && c < C) ? __half2float(x[((size_t)n * HW + p) * x_stride + c]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    if (c < C && p < HW) out[((size_t)n * C + c) * HW + p] = (TOut)tile[threadIdx.x][i];
  }
}

// fp32 (rows, C) -> fp16 (rows_pad, C) with zero rows appended per batch (prompt context 77 -> 80 tokens)
__global__ void pad_tokens_kernel(const float* __restrict__ x, __half* __restrict__ out, int B, int T, int Tpad,
                                  int C) {
  tf::pdl_prologue();  // PDL: let the next kernel start launching, then wait for our producers
  const long total = (long)B * Tpad * C;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const int t = (int)((idx / C) % Tpad);
    const int b = (int)(idx / ((long)C * Tpad));
    out[idx] = t < T ? __float2half_rn(x[((size_t)b * T + t) * C + c]) : __float2half_rn(0.f);
  }
}

// CFG combine + DDIM (eta = 0) update, fused.
// reference: variants/sd.py:44-45 (e_t = u + g (c - u)) and sd.py:14-25 (x_prev).
// eps: fp32 NHWC, pixel stride eps_stride, images [0,B) = unconditional, [B,2B) = conditional
// (order [uncond ; cond], sd.py:32). latent in/out: fp32 NCHW (B, C, H, W).
__global__ void cfg_ddim_kernel(const float* __restrict__ eps, int eps_stride, const float* __restrict__ latent,
                                float* __restrict__ latent_out, float* __restrict__ e_t_out,
                                const float* __restrict__ a_t_tab, const float* __restrict__ a_prev_tab,
                                const int* __restrict__ idx_dev, float guidance, int B, int C, int HW) {
  tf::pdl_prologue();  // PDL: let the next kernel start launching, then wait for our producers
  const int idx = idx_dev ? *idx_dev : 0;
  const float a_t = a_t_tab[idx], a_prev = a_prev_tab[idx];
  const float sqrt_one_minus_at = sqrtf(1.f - a_t);
  const float inv_sqrt_at = 1.f / sqrtf(a_t);
  const float sqrt_aprev = sqrtf(a_prev);
  const float dir_coef = sqrtf(1.f - a_prev);
  const long total = (long)B * C * HW;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    const int c = (int)((i / HW) % C);
    const int b = (int)(i / ((long)HW * C));
    const float u = eps[((size_t)b * HW + p) * eps_stride + c];
    const float cnd = eps[((size_t)(B + b) * HW + p) * eps_stride + c];
    const float e = u + guidance * (cnd - u);
    const float x = latent[i];
    const float pred_x0 = (x - sqrt_one_minus_at * e) * inv_sqrt_at;
    latent_out[i] = sqrt_aprev * pred_x0 + dir_coef * e;
    if (e_t_out) e_t_out[i] = e;
  }
}

// latent (B,C,H,W) fp32 -> (2B,C,H,W) fp32 duplicated (reference: broadcast_to in sd.py:31); trivial copy
__global__ void add_int_kernel(int* p, int delta) {
  tf::pdl_prologue();  // PDL: let the next kernel start launching, then wait for our producers
  if (threadIdx.x == 0 && blockIdx.x == 0) *p += delta;
}

static int ew_blocks(long total, int threads) {
  long b = (total + threads - 1) / threads;
  const long cap = (long)tf_num_sms() * 8;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace

extern "C" int tf_timestep_embedding_f32(const float* timesteps_dev, const int* index_dev, int dim,
                                         float max_period, float* out, void* stream) {
  TF_CHECK_ARG(timesteps_dev && out && dim > 0 && dim % 2 == 0, "tf_timestep_embedding_f32: bad arguments");
  const int half = dim / 2;
  TF_LAUNCH(timestep_embedding_kernel, ceil_div_i(half, 128), 128, 0, (cudaStream_t)stream, timesteps_dev, index_dev, dim,
                                                                                    max_period, out);
  TF_LAUNCH_CHECK();
  tf_launch_count_add(1);
  return TF_OK;
}

extern "C" int tf_gemv_f16w(const float* x, const void* W, const float* bias, const float* bias2, float* out, int N,
                            int K, int silu_input, void* stream) {
  TF_CHECK_ARG(x && W && out && N > 0 && K > 0 && K % 8 == 0, "tf_gemv_f16w: bad arguments (N=%d K=%d)", N, K);
  TF_CHECK_ARG(K * sizeof(float) <= 48 * 1024, "tf_gemv_f16w: K too large (%d)", K);
  const int threads = 256;
  TF_LAUNCH(gemv_kernel, ceil_div_i(N, threads / 32), threads, K * sizeof(float), (cudaStream_t)stream, 
      x, (const __half*)W, bias, bias2, out, N, K, silu_input);
  TF_LAUNCH_CHECK();
  tf_launch_count_add(1);
  return TF_OK;
}

static int smallcin_impl(const float* x, int x_images, const float* w, const float* bias, void* out, int NI, int Cin, int H,
                         int W, int Cout, int out_pixel_stride, void* gn_stats, int gn_unit, void* stream) {
  TF_CHECK_ARG(x && w && out && x_images > 0, "tf_conv3x3_smallcin_f32nchw: null pointer");
  TF_CHECK_ARG(Cin == 4, "tf_conv3x3_smallcin_f32nchw: only Cin == 4 is built (got %d)", Cin);
  const size_t smem = (size_t)(Cout * Cin * 9 + Cout) * sizeof(float) + (gn_stats ? (size_t)Cout * sizeof(float2) : 0);
  TF_CHECK_ARG(Cout % 8 == 0 && out_pixel_stride % 8 == 0 && smem <= 100 * 1024,
               "tf_conv3x3_smallcin_f32nchw: bad Cout %d", Cout);
  if (gn_stats)
    TF_CHECK_ARG(gn_unit > 0 && Cout % gn_unit == 0 && (H * W) % 32 == 0,
                 "tf_conv3x3_smallcin_gn_f32nchw: statistics need Cout %% unit == 0 and H*W %% 32 == 0");
  const long npix = (long)NI * H * W;
  static bool attr = false;
  if (!attr) {
    TF_CUDA(cudaFuncSetAttribute(conv3x3_smallcin_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr = true;
  }
  TF_CHECK_ARG(((uintptr_t)w & 15) == 0, "tf_conv3x3_smallcin_f32nchw: weights must be 16-byte aligned");
  TF_LAUNCH((conv3x3_smallcin_kernel<4>), (unsigned)((npix + 31) / 32), 256, smem, (cudaStream_t)stream,
            x, w, bias, (__half*)out, NI, H, W, Cout, out_pixel_stride, x_images, (float2*)gn_stats, gn_unit);
  TF_LAUNCH_CHECK();
  tf_launch_count_add(1);
  return TF_OK;
}

extern "C" int tf_conv3x3_smallcin_f32nchw(const float* x, int x_images, const float* w, const float* bias, void* out,
                                           int NI, int Cin, int H, int W, int Cout, int out_pixel_stride,
                                           void* stream) {
  return smallcin_impl(x, x_images, w, bias, out, NI, Cin, H, W, Cout, out_pixel_stride, nullptr, 0, stream);
}

extern "C" int tf_conv3x3_smallcin_gn_f32nchw(const float* x, int x_images, const float* w, const float* bias, void* out,
                                              int NI, int Cin, int H, int W, int Cout, int out_pixel_stride,
                                              void* gn_stats, int gn_unit, void* stream) {
  TF_CHECK_ARG(gn_stats != nullptr, "tf_conv3x3_smallcin_gn_f32nchw: null statistics buffer");
  return smallcin_impl(x, x_images, w, bias, out, NI, Cin, H, W, Cout, out_pixel_stride, gn_stats, gn_unit, stream);
}

extern "C" int tf_upsample_nearest2x_nhwc_f16(const void* x, int x_pixel_stride, void* out, int out_pixel_stride,
                                              int NI, int H, int W, int C, void* stream) {
  TF_CHECK_ARG(x && out && C % 8 == 0 && x_pixel_stride % 8 == 0 && out_pixel_stride % 8 == 0,
               "tf_upsample_nearest2x_nhwc_f16: bad arguments");
  const long total = (long)NI * 4 * H * W * (C / 8);
  TF_LAUNCH(upsample2x_kernel, ew_blocks(total, 256), 256, 0, (cudaStream_t)stream, (const __half*)x, x_pixel_stride,
                                                                            (__half*)out, out_pixel_stride, NI, H, W, C);
  TF_LAUNCH_CHECK();
  tf_launch_count_add(1);
  return TF_OK;
}

extern "C" int tf_nchw_to_nhwc_f16(const void* x, int x_is_f32, void* out, int NI, int C, int HW,
                                   int out_pixel_stride, void* stream) {
  TF_CHECK_ARG(x && out && NI > 0 && C > 0 && HW > 0 && out_pixel_stride >= C, "tf_nchw_to_nhwc_f16: bad arguments");
  dim3 grid(ceil_div_i(HW, 32), ceil_div_i(C, 32), NI), block(32, 8);
  if (x_is_f32)
    TF_LAUNCH((nchw_to_nhwc_kernel<float>), grid, block, 0, (cudaStream_t)stream, (const float*)x, (__half*)out, C, HW,
                                                                         out_pixel_stride);
  else
    TF_LAUNCH((nchw_to_nhwc_kernel<__half>), grid, block, 0, (cudaStream_t)stream, (const __half*)x, (__half*)out, C, HW,
                                                                          out_pixel_stride);
  TF_LAUNCH_CHECK();
  tf_launch_count_add(1);
  return TF_OK;
}

extern "C" int tf_nhwc_to_nchw(const void* x, int x_pixel_stride, void* out, int out_is_f32, int NI, int C, int HW,
                               void* stream) {
  TF_CHECK_ARG(x && out && NI > 0 && C > 0 && HW > 0 && x_pixel_stride >= C, "tf_nhwc_to_nchw: bad arguments");
  dim3 grid(ceil_div_i(HW, 32), ceil_div_i(C, 32), NI), block(32, 8);
  if (out_is_f32)
    TF_LAUNCH((nhwc_to_nchw_kernel<float>), grid, block, 0, (cudaStream_t)stream, (const __half*)x, x_pixel_stride, (float*)out,
                                                                         C, HW);
  else
    TF_LAUNCH((nhwc_to_nchw_kernel<__half>), grid, block, 0, (cudaStream_t)stream, (const __half*)x, x_pixel_stride,
                                                                          (__half*)out, C, HW);
  TF_LAUNCH_CHECK();
  tf_launch_count_add(1);
  return TF_OK;
}

extern "C" int tf_pad_tokens_f32_to_f16(const float* x, void* out, int B, int T, int Tpad, int C, void* stream) {
  TF_CHECK_ARG(x && out && B > 0 && T > 0 && Tpad >= T && C > 0, "tf_pad_tokens_f32_to_f16: bad arguments");
  const long total = (long)B * Tpad * C;
  TF_LAUNCH(pad_tokens_kernel, ew_blocks(total, 256), 256, 0, (cudaStream_t)stream, x, (__half*)out, B, T, Tpad, C);
  TF_LAUNCH_CHECK();
  tf_launch_count_add(1);
  return TF_OK;
}

extern "C" int tf_cfg_ddim_step_f32(const float* eps_nhwc, int eps_pixel_stride, const float* latent,
                                    float* latent_out, float* e_t_out, const float* alphas_dev,
                                    const float* alphas_prev_dev, const int* index_dev, float guidance, int B, int C,
                                    int HW, void* stream) {
  TF_CHECK_ARG(eps_nhwc && latent && latent_out && alphas_dev && alphas_prev_dev,
               "tf_cfg_ddim_step_f32: null pointer");
  TF_CHECK_ARG(B > 0 && C > 0 && HW > 0 && eps_pixel_stride >= C, "tf_cfg_ddim_step_f32: bad dims");
  const long total = (long)B * C * HW;
  TF_LAUNCH(cfg_ddim_kernel, ew_blocks(total, 256), 256, 0, (cudaStream_t)stream, 
      eps_nhwc, eps_pixel_stride, latent, latent_out, e_t_out, alphas_dev, alphas_prev_dev, index_dev, guidance, B, C,
      HW);
  TF_LAUNCH_CHECK();
  tf_launch_count_add(1);
  return TF_OK;
}

extern "C" int tf_add_int(int* p_dev, int delta, void* stream) {
  TF_CHECK_ARG(p_dev, "tf_add_int: null pointer");
  TF_LAUNCH(add_int_kernel, 1, 32, 0, (cudaStream_t)stream, p_dev, delta);
  TF_LAUNCH_CHECK();
  tf_launch_count_add(1);
  return TF_OK;
}

// ------------------------------------------------------------------------------------------------
// stand-alone activations (reference: Tensor.sigmoid/silu/gelu/quick_gelu, storage/tensor.py:64-86).
// On the UNet fast path these are fused into GroupNorm / GEMM epilogues; this entry point keeps the
// drop-in static methods working on their own.
// ------------------------------------------------------------------------------------------------
namespace {
template <typename T>
__global__ void unary_kernel(const T* __restrict__ x, T* __restrict__ out, long n, int op) {
  tf::pdl_prologue();  // PDL: let the next kernel start launching, then wait for our producers
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float v = (float)x[i];
    float r;
    switch (op) {
      case 0: r = 1.0f / (1.0f + __expf(-v)); break;
      case 1: r = tf::silu_f(v); break;
      case 2: r = tf::gelu_tanh_f(v); break;
      default: r = v / (1.0f + __expf(-1.702f * v)); break;
    }
    out[i] = (T)r;
  }
}
}  // namespace

extern "C" int tf_unary(const void* x, void* out, long long n, int op, int is_f32, void* stream) {
  TF_CHECK_ARG(x && out && n >= 0 && op >= 0 && op <= 3, "tf_unary: bad arguments");
  if (n == 0) return TF_OK;
  const int threads = 256;
  long b = (n + threads - 1) / threads;
  const long cap = (long)tf_num_sms() * 8;
  const int blocks = (int)(b < cap ? b : cap);
  if (is_f32)
    TF_LAUNCH((unary_kernel<float>), blocks, threads, 0, (cudaStream_t)stream, (const float*)x, (float*)out, n, op);
  else
    TF_LAUNCH((unary_kernel<__half>), blocks, threads, 0, (cudaStream_t)stream, (const __half*)x, (__half*)out, n, op);
  TF_LAUNCH_CHECK();
  tf_launch_count_add(1);
  return TF_OK;
}

namespace {
__global__ void nhwc_f32_to_nchw_f32_kernel(const float* __restrict__ x, int x_stride, float* __restrict__ out, int NI,
                                            int C, int HW) {
  tf::pdl_prologue();  // PDL: let the next kernel start launching, then wait for our producers
  const long total = (long)NI * C * HW;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    const int c = (int)((i / HW) % C);
    const int n = (int)(i / ((long)HW * C));
    out[i] = x[((size_t)n * HW + p) * x_stride + c];
  }
}
// DDIM (eta = 0) update on its own (reference: variants/sd.py:14-25)
__global__ void ddim_kernel(const float* __restrict__ x, const float* __restrict__ e_t, const float* __restrict__ a_t_p,
                            const float* __restrict__ a_prev_p, float* __restrict__ x_prev, float* __restrict__ pred_x0,
                            long n) {
  tf::pdl_prologue();  // PDL: let the next kernel start launching, then wait for our producers
  const float a_t = *a_t_p, a_prev = *a_prev_p;
  const float s1 = sqrtf(1.f - a_t), is = 1.f / sqrtf(a_t), sp = sqrtf(a_prev), dp = sqrtf(1.f - a_prev);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float e = e_t[i];
    const float p0 = (x[i] - s1 * e) * is;
    x_prev[i] = sp * p0 + dp * e;
    if (pred_x0) pred_x0[i] = p0;
  }
}
}  // namespace

extern "C" int tf_nhwc_f32_to_nchw_f32(const float* x, int x_pixel_stride, float* out, int NI, int C, int HW,
                                       void* stream) {
  TF_CHECK_ARG(x && out && NI > 0 && C > 0 && HW > 0 && x_pixel_stride >= C, "tf_nhwc_f32_to_nchw_f32: bad arguments");
  const long total = (long)NI * C * HW;
  TF_LAUNCH(nhwc_f32_to_nchw_f32_kernel, ew_blocks(total, 256), 256, 0, (cudaStream_t)stream, x, x_pixel_stride, out, NI, C, HW);
  TF_LAUNCH_CHECK();
  tf_launch_count_add(1);
  return TF_OK;
}

extern "C" int tf_ddim_step_f32(const float* x, const float* e_t, const float* a_t_dev, const float* a_prev_dev,
                                float* x_prev, float* pred_x0, long long n, void* stream) {
  TF_CHECK_ARG(x && e_t && a_t_dev && a_prev_dev && x_prev && n > 0, "tf_ddim_step_f32: bad arguments");
  TF_LAUNCH(ddim_kernel, ew_blocks(n, 256), 256, 0, (cudaStream_t)stream, x, e_t, a_t_dev, a_prev_dev, x_prev, pred_x0, n);
  TF_LAUNCH_CHECK();
  tf_launch_count_add(1);
  return TF_OK;
}
