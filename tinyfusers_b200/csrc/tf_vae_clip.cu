// tinyfusers_b200 — small kernels of the rows next to the UNet hot path: the VAE decoder's
// AttnBlock attention, post_quant_conv and image post-processing, and the CLIP text encoder's embedding lookup.
// The heavy parts of both models (3x3 / 1x1 convs, Linears, GroupNorm, LayerNorm, 12-head causal attention) run on
// the tcgen05 GEMM / conv / attention kernels of tf_gemm.cu / tf_attention.cu / tf_norm.cu.
#include "tf_common.cuh"
#include "tinyfusers_b200.h"

namespace {

// ------------------------------------------------------------------------------------------------
// Plane attention — the reference's AttnBlock (tinyfusers/attention/attention.py:19-24) hands 4-D (B,C,H,W) q/k/v to
// scaled_dot_product_attention (attention/sdpa.py:53-77), which reads them as (B, NH = C, T = H, HS = W): for every
// channel plane, S = scale * Q K^T (H x H), row softmax, O = P V (H x W), scale = 1/sqrt(W). One CTA per plane; the
// three planes and the score matrix live in shared memory (rows padded to an odd word count), fp32 arithmetic.
// 0.5 GFLOP for the whole 512-channel 64x64 block: latency, not throughput, is what matters here.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
plane_attention_kernel(const __half* __restrict__ q, const __half* __restrict__ k, const __half* __restrict__ v,
                       __half* __restrict__ out, int H, int W, float scale_log2) {
  tf::pdl_prologue();
  extern __shared__ __align__(16) uint8_t psm[];
  const int WP = W + 2;                       // halfs per padded row: (W + 2) / 2 words is odd for W % 4 == 0
  __half* sq = reinterpret_cast<__half*>(psm);
  __half* sk = sq + H * WP;
  __half* sv = sk + H * WP;
  float* sS = reinterpret_cast<float*>(sv + H * WP);
  const int HP = H + 1;
  const size_t plane = (size_t)blockIdx.x * H * W;
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int i = tid; i < H * (W >> 1); i += nthr) {
    const int r = i / (W >> 1), c2 = i % (W >> 1);
    const size_t g = plane + (size_t)r * W + 2 * c2;
    *reinterpret_cast<__half2*>(sq + r * WP + 2 * c2) = *reinterpret_cast<const __half2*>(q + g);
    *reinterpret_cast<__half2*>(sk + r * WP + 2 * c2) = *reinterpret_cast<const __half2*>(k + g);
    *reinterpret_cast<__half2*>(sv + r * WP + 2 * c2) = *reinterpret_cast<const __half2*>(v + g);
  }
  __syncthreads();
  // S[i][j] = scale * sum_w q[i][w] k[j][w]   (consecutive threads -> consecutive j: q row is a broadcast)
  for (int idx = tid; idx < H * H; idx += nthr) {
    const int i = idx / H, j = idx % H;
    const __half2* qi = reinterpret_cast<const __half2*>(sq + i * WP);
    const __half2* kj = reinterpret_cast<const __half2*>(sk + j * WP);
    float a0 = 0.f, a1 = 0.f;
    for (int w2 = 0; w2 < (W >> 1); ++w2) {
      const float2 a = __half22float2(qi[w2]), b = __half22float2(kj[w2]);
      a0 = fmaf(a.x, b.x, a0);
      a1 = fmaf(a.y, b.y, a1);
    }
    sS[i * HP + j] = (a0 + a1) * scale_log2;
  }
  __syncthreads();
  // row softmax (softmax.cu:24-112: max, exp(x - max), sum, divide), one warp per row
  const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
  for (int i = warp; i < H; i += nwarp) {
    float m = -INFINITY;
    for (int j = lane; j < H; j += 32) m = fmaxf(m, sS[i * HP + j]);
    m = tf::warp_max(m);
    float s = 0.f;
    for (int j = lane; j < H; j += 32) {
      const float e = exp2f(sS[i * HP + j] - m);
      sS[i * HP + j] = e;
      s += e;
    }
    s = tf::warp_sum(s);
    const float inv = 1.0f / s;
    for (int j = lane; j < H; j += 32) sS[i * HP + j] *= inv;
  }
  __syncthreads();
  // O[i][w] = sum_j P[i][j] v[j][w]; thread -> (row, pair of columns)
  for (int idx = tid; idx < H * (W >> 1); idx += nthr) {
    const int i = idx / (W >> 1), c2 = idx % (W >> 1);
    float a0 = 0.f, a1 = 0.f;
    for (int j = 0; j < H; ++j) {
      const float pj = sS[i * HP + j];
      const float2 vv = __half22float2(*reinterpret_cast<const __half2*>(sv + j * WP + 2 * c2));
      a0 = fmaf(pj, vv.x, a0);
      a1 = fmaf(pj, vv.y, a1);
    }
    *reinterpret_cast<__half2*>(out + plane + (size_t)i * W + 2 * c2) = __floats2half2_rn(a0, a1);
  }
}

// out[n, co, p] = bias[co] + sum_ci w[co, ci] * (scale * x[n, ci, p]), fp32 NCHW, Cin, Cout <= 8
__global__ void conv1x1_small_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                     float* __restrict__ out, int Cin, int Cout, long HW, float scale, long total) {
  tf::pdl_prologue();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long n = i / HW, p = i % HW;
    float xin[8];
    for (int c = 0; c < Cin; ++c) xin[c] = scale * x[(n * Cin + c) * HW + p];
    for (int co = 0; co < Cout; ++co) {
      float a = bias ? bias[co] : 0.f;
      for (int c = 0; c < Cin; ++c) a = fmaf(w[co * Cin + c], xin[c], a);
      out[(n * Cout + co) * HW + p] = a;
    }
  }
}

// clip((x + 1) / 2, 0, 1) * 255 -> uint8 (truncation), fp32 NHWC (pixel stride ld) -> uint8 HWC
__global__ void image_u8_kernel(const float* __restrict__ x, int ld, unsigned char* __restrict__ out, int C, long total) {
  tf::pdl_prologue();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long p = i / C;
    const int c = (int)(i % C);
    const float v = fminf(fmaxf((x[p * ld + c] + 1.0f) / 2.0f, 0.f), 1.f) * 255.f;
    out[i] = (unsigned char)v;
  }
}

// out[t, :] = tok[ids[t], :] + pos[t % T, :]   (fp32 tables -> fp16 rows)
__global__ void embedding_kernel(const int* __restrict__ ids, const float* __restrict__ tok, const float* __restrict__ pos,
                                 __half* __restrict__ out, int T, int E, int vocab) {
  tf::pdl_prologue();
  const int t = blockIdx.x;
  int id = ids[t];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const float* tr = tok + (size_t)id * E;
  const float* pr = pos ? pos + (size_t)(t % T) * E : nullptr;
  for (int e = threadIdx.x; e < E; e += blockDim.x) out[(size_t)t * E + e] = __float2half_rn(tr[e] + (pr ? pr[e] : 0.f));
}

__global__ void cast_f16_f32_kernel(const __half* __restrict__ x, float* __restrict__ out, long n) {
  tf::pdl_prologue();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) out[i] = __half2float(x[i]);
}

int ew_grid(long n, int threads) {
  long b = (n + threads - 1) / threads;
  const long cap = (long)tf_num_sms() * 16;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace

namespace {
// Row softmax of the canonical (single-head, head dim = channels) AttnBlock: P[r, :] = softmax(scale * S[r, :]),
// fp32 scores from the QK^T GEMM -> fp16 probabilities for the PV GEMM. One block per row, three passes over a row
// that stays in L1/L2 (reference: native/cuda/softmax.cu:24-112 does the same three passes in fp32).
__global__ void __launch_bounds__(256)
softmax_f32_to_f16_kernel(const float* __restrict__ S, long lds, __half* __restrict__ P, long ldp, int cols, float scale_log2) {
  tf::pdl_prologue();
  __shared__ float red[8];
  const float* __restrict__ s = S + (long)blockIdx.x * lds;
  __half* __restrict__ p = P + (long)blockIdx.x * ldp;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float m = -INFINITY;
  for (int c = threadIdx.x; c < cols; c += 256) m = fmaxf(m, s[c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[w] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int c = threadIdx.x; c < cols; c += 256) sum += exp2f((s[c] - m) * scale_log2);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) red[w] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) sum += red[i];
  const float inv = 1.f / sum;
  for (int c = threadIdx.x; c < cols; c += 256) p[c] = __float2half_rn(exp2f((s[c] - m) * scale_log2) * inv);
}
}  // namespace

extern "C" int tf_softmax_rows_f32_to_f16(const float* scores, long long lds, void* probs, long long ldp, long long rows,
                                          int cols, float scale, void* stream) {
  TF_CHECK_ARG(scores && probs && rows > 0 && rows < (1ll << 31) && cols > 0 && lds >= cols && ldp >= cols && scale > 0.f,
               "tf_softmax_rows_f32_to_f16: bad arguments");
  TF_LAUNCH(softmax_f32_to_f16_kernel, (unsigned)rows, 256, 0, (cudaStream_t)stream, scores, (long)lds, (__half*)probs,
            (long)ldp, cols, scale * 1.4426950408889634f);
  TF_LAUNCH_CHECK();
  tf_launch_count_add(1);
  return TF_OK;
}

extern "C" int tf_plane_attention_f16(const void* q, const void* k, const void* v, void* out, int planes, int H, int W,
                                      float scale, void* stream) {
  TF_CHECK_ARG(q && k && v && out && planes > 0, "tf_plane_attention_f16: null pointer");
  TF_CHECK_ARG(H > 0 && W > 0 && W % 2 == 0, "tf_plane_attention_f16: W must be even (H=%d W=%d)", H, W);
  const size_t smem = (size_t)3 * H * (W + 2) * sizeof(__half) + (size_t)H * (H + 1) * sizeof(float);
  TF_CHECK_ARG(smem <= 200 * 1024, "tf_plane_attention_f16: a %d x %d plane does not fit shared memory", H, W);
  static size_t attr_bytes = 0;
  if (smem > 48 * 1024 && smem > attr_bytes) {
    TF_CUDA(cudaFuncSetAttribute(plane_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_bytes = 200 * 1024;
  }
  TF_LAUNCH(plane_attention_kernel, planes, 256, smem, (cudaStream_t)stream, (const __half*)q, (const __half*)k,
            (const __half*)v, (__half*)out, H, W, scale * 1.4426950408889634f);
  TF_LAUNCH_CHECK();
  tf_launch_count_add(1);
  return TF_OK;
}

extern "C" int tf_conv1x1_small_f32nchw(const float* x, const float* w, const float* bias, float* out, int NI, int Cin,
                                        int Cout, int HW, float scale, void* stream) {
  TF_CHECK_ARG(x && w && out && NI > 0 && HW > 0, "tf_conv1x1_small_f32nchw: bad arguments");
  TF_CHECK_ARG(Cin > 0 && Cin <= 8 && Cout > 0 && Cout <= 8, "tf_conv1x1_small_f32nchw: Cin, Cout <= 8 (got %d, %d)", Cin, Cout);
  const long total = (long)NI * HW;
  TF_LAUNCH(conv1x1_small_kernel, ew_grid(total, 256), 256, 0, (cudaStream_t)stream, x, w, bias, out, Cin, Cout, (long)HW,
            scale, total);
  TF_LAUNCH_CHECK();
  tf_launch_count_add(1);
  return TF_OK;
}

extern "C" int tf_image_to_u8(const float* x_nhwc, int x_pixel_stride, void* out_u8, long long pixels, int C, void* stream) {
  TF_CHECK_ARG(x_nhwc && out_u8 && pixels > 0 && C > 0 && x_pixel_stride >= C, "tf_image_to_u8: bad arguments");
  const long total = (long)pixels * C;
  TF_LAUNCH(image_u8_kernel, ew_grid(total, 256), 256, 0, (cudaStream_t)stream, x_nhwc, x_pixel_stride,
            (unsigned char*)out_u8, C, total);
  TF_LAUNCH_CHECK();
  tf_launch_count_add(1);
  return TF_OK;
}

extern "C" int tf_embedding_f16(const int* ids, const float* table, const float* pos_table, void* out, int rows, int T,
                                int E, int vocab, void* stream) {
  TF_CHECK_ARG(ids && table && out && rows > 0 && T > 0 && E > 0 && vocab > 0, "tf_embedding_f16: bad arguments");
  TF_LAUNCH(embedding_kernel, rows, 256, 0, (cudaStream_t)stream, ids, table, pos_table, (__half*)out, T, E, vocab);
  TF_LAUNCH_CHECK();
  tf_launch_count_add(1);
  return TF_OK;
}

extern "C" int tf_cast_f16_to_f32(const void* x, float* out, long long n, void* stream) {
  TF_CHECK_ARG(x && out && n >= 0, "tf_cast_f16_to_f32: bad arguments");
  if (n == 0) return TF_OK;
  TF_LAUNCH(cast_f16_f32_kernel, ew_grid((long)n, 256), 256, 0, (cudaStream_t)stream, (const __half*)x, out, (long)n);
  TF_LAUNCH_CHECK();
  tf_launch_count_add(1);
  return TF_OK;
}
