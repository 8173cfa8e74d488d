// tf_fp32.cu — the fp32 PARITY mode of the hot path (BASELINE.json configs[0]: down-block-0 in fp32, per-op error <= 1e-5).
//
// The reference computes everything in fp32 (cuBLAS SGEMM, cuDNN fp32 graphs, CuPy elementwise; SURVEY.md §8 banner).
// The fast path of this library is fp16-operand / fp32-accumulate on tcgen05 and is held to 1e-2; this file is the same
// operator set in plain fp32 so that the 1e-5 bar of the north star can be checked on the GPU: CUDA-core FMA kernels in the
// reference's own layouts (NCHW images, (B,T,C) tokens), no tensor cores (kind::tf32 keeps 10 mantissa bits - not 1e-5).
// It is a correctness mode, not the measured path: bench.py never runs it. Compiled WITHOUT --use_fast_math
// (csrc/build.py) so expf / tanhf / division / sqrt are the IEEE-accurate versions.
//
//   tf_gemm_f32            out = alpha * A.W^T (+bias) (+residual), batched over (outer, inner) with separate strides
//                          (heads), W as [N][K] or [K][N], row-major or NCHW output
//                          (Linear ff/linear.py:119-120, cp.matmul of attention/sdpa.py:66,76, 1x1 proj_out)
//   tf_conv2d_nchw_f32     implicit-GEMM cross-correlation straight from NCHW / OIHW (vision/conv2d.py:9-28,55-59)
//   tf_groupnorm_nchw_f32  literal two-pass GroupNorm (+affine, +SiLU) (ff/group_norm.py:3-21)
//   tf_layernorm_f32       LayerNorm over the last dimension (ff/layer_norm.py:8-32, the semantics cuDNN executes)
//   tf_softmax_rows_f32    max-subtracted row softmax, in place (native/cuda/softmax.cu:24-112)
//   tf_unary_f32 / tf_geglu_f32   activations and the GEGLU gate (storage/tensor.py:64-86, ff/nn.py:5-12)
//   tf_cfg_combine_f32     e_t = u + g (c - u) (variants/sd.py:44-45)
//
// Sums over K are blocked (16 products into a fresh partial, partials into the accumulator) so the rounding error grows
// with K/16 + 16 instead of K.
#include <cuda_runtime.h>

#include "tf_common.cuh"
#include "tinyfusers_b200.h"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, PITCH = BM + 4;

struct GemmF32 {
  const float* A;
  const float* W;
  const float* bias;      // [N] or null
  const float* bias_img;  // [images][N] or null: added after bias (ResBlock emb projection, resnet.py:27)
  const float* residual;  // addressed like out, or null
  float* out;
  int M, N, K;
  int a_conv;  // 0: A[m * lda + k]; 1: NCHW gather
  long lda;
  int C, H, Wd, R, S, stride, pad, Ho, Wo;  // a_conv geometry: K = C*R*S, M = images*Ho*Wo
  int w_kn;                                  // 0: W[n * ldw + k]; 1: W[k * ldw + n]
  long ldw;
  int out_nchw;  // 0: out[m * ldc + n]; 1: out[(img * N + n) * rows_per_img + pix]
  long ldc, ldr;
  int rows_per_img;  // rows of one image (bias_img / NCHW output)
  float alpha;
  int inner;  // batch z -> (z / inner, z % inner)
  long sAo, sAi, sWo, sWi, sCo, sCi;
};

__global__ void __launch_bounds__(256) gemm_f32_kernel(const GemmF32 p) {
  tf::pdl_prologue();
  __shared__ float As[BK][PITCH];
  __shared__ float Bs[BK][PITCH];
  const int t = threadIdx.x;
  const int zo = blockIdx.z / p.inner, zi = blockIdx.z % p.inner;
  const float* __restrict__ A = p.A + zo * p.sAo + zi * p.sAi;
  const float* __restrict__ W = p.W + zo * p.sWo + zi * p.sWi;
  const long c_off = zo * p.sCo + zi * p.sCi;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  // thread -> 4 x 4 outputs, strided by 16 so that the fastest-varying thread index runs along the contiguous output dim
  const int fast = t & 15, slow = t >> 4;
  const int rsel = p.out_nchw ? fast : slow, csel = p.out_nchw ? slow : fast;

  // conv gather: the pixel of this thread's A-tile row is fixed for the whole K loop
  const int am = t & 63, ak0 = t >> 6;
  int g_img = 0, g_ih0 = 0, g_iw0 = 0;
  bool g_valid = false;
  if (p.a_conv) {
    const int m = m0 + am;
    g_valid = m < p.M;
    if (g_valid) {
      const int hw = p.Ho * p.Wo;
      g_img = m / hw;
      const int pix = m - g_img * hw;
      const int oh = pix / p.Wo, ow = pix - oh * p.Wo;
      g_ih0 = oh * p.stride - p.pad;
      g_iw0 = ow * p.stride - p.pad;
    }
  }
  const int RS = p.R * p.S;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < p.K; k0 += BK) {
    // ---- A tile -> As[k][m]
    if (p.a_conv) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kl = ak0 + 4 * i, k = k0 + kl;
        float v = 0.f;
        if (g_valid && k < p.K) {
          const int c = k / RS, rs = k - c * RS;
          const int r = rs / p.S, s = rs - r * p.S;
          const int ih = g_ih0 + r, iw = g_iw0 + s;
          if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.Wd) v = A[(((long)g_img * p.C + c) * p.H + ih) * p.Wd + iw];
        }
        As[kl][am] = v;
      }
    } else {
      const int row = t >> 2, kq = (t & 3) * 4;
      const int m = m0 + row;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = k0 + kq + i;
        As[kq + i][row] = (m < p.M && k < p.K) ? A[(long)m * p.lda + k] : 0.f;
      }
    }
    // ---- W tile -> Bs[k][n]
    if (p.w_kn) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kl = ak0 + 4 * i, k = k0 + kl, n = n0 + am;
        Bs[kl][am] = (n < p.N && k < p.K) ? W[(long)k * p.ldw + n] : 0.f;
      }
    } else {
      const int row = t >> 2, kq = (t & 3) * 4;
      const int n = n0 + row;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = k0 + kq + i;
        Bs[kq + i][row] = (n < p.N && k < p.K) ? W[(long)n * p.ldw + k] : 0.f;
      }
    }
    __syncthreads();
    float part[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) part[i][j] = 0.f;
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][rsel + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][csel + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) part[i][j] = fmaf(a[i], b[j], part[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += part[i][j];
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + rsel + 16 * i;
    if (m >= p.M) continue;
    const int img = p.rows_per_img > 0 ? m / p.rows_per_img : 0;
    const int pix = p.rows_per_img > 0 ? m - img * p.rows_per_img : 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + csel + 16 * j;
      if (n >= p.N) continue;
      float v = p.alpha * acc[i][j];
      if (p.bias) v += p.bias[n];
      if (p.bias_img) v += p.bias_img[(long)img * p.N + n];
      const long o = p.out_nchw ? ((long)img * p.N + n) * p.rows_per_img + pix : (long)m * p.ldc + n;
      const long ro = p.out_nchw ? o : (long)m * p.ldr + n;
      if (p.residual) v += p.residual[c_off + ro];
      p.out[c_off + o] = v;
    }
  }
}

int launch_gemm(const GemmF32& p, int batch, cudaStream_t stream) {
  dim3 grid((p.M + BM - 1) / BM, (p.N + BN - 1) / BN, batch);
  TF_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "fp32 gemm: grid too large (N=%d batch=%d)", p.N, batch);
  TF_LAUNCH(gemm_f32_kernel, grid, 256, 0, stream, p);
  TF_LAUNCH_CHECK();
  tf_launch_count_add(1);
  return TF_OK;
}

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  float s = 0.f;
  for (int i = 0; i < nw; ++i) s += red[i];
  return s;
}
__device__ __forceinline__ float block_max(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  float s = red[0];
  for (int i = 1; i < nw; ++i) s = fmaxf(s, red[i]);
  return s;
}

__device__ __forceinline__ float sigmoid_exact(float x) { return 1.f / (1.f + expf(-x)); }

// one block per (image, group); the group's channels are contiguous in NCHW
__global__ void __launch_bounds__(512) groupnorm_nchw_f32_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, float* __restrict__ out,
                                                                 int C, int HW, int groups, float eps, int silu) {
  tf::pdl_prologue();
  __shared__ float red[16];
  const int cpg = C / groups;
  const int n = blockIdx.x / groups, g = blockIdx.x % groups;
  const long base = ((long)n * C + (long)g * cpg) * HW;
  const long len = (long)cpg * HW;
  const float* __restrict__ xp = x + base;
  float s = 0.f;
  for (long i = threadIdx.x; i < len; i += blockDim.x) s += xp[i];
  const float mean = block_sum(s, red) / (float)len;
  float q = 0.f;
  for (long i = threadIdx.x; i < len; i += blockDim.x) {
    const float d = xp[i] - mean;
    q = fmaf(d, d, q);
  }
  const float var = block_sum(q, red) / (float)len;
  const float inv = 1.f / sqrtf(var + eps);  // group_norm.py:9-10: yn * (1 / sqrt(mean(yn^2) + eps))
  for (long i = threadIdx.x; i < len; i += blockDim.x) {
    const int c = g * cpg + (int)(i / HW);
    float v = (xp[i] - mean) * inv;
    if (gamma) v = v * gamma[c] + (beta ? beta[c] : 0.f);
    if (silu) v = v * sigmoid_exact(v);
    out[base + i] = v;
  }
}

// one warp per row
__global__ void __launch_bounds__(128) layernorm_f32_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float* __restrict__ out, long rows,
                                                            int C, float eps) {
  tf::pdl_prologue();
  const long row = (long)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* __restrict__ xp = x + row * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += xp[c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)C;
  float q = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float d = xp[c] - mean;
    q = fmaf(d, d, q);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float inv = 1.f / sqrtf(q / (float)C + eps);
  for (int c = lane; c < C; c += 32) {
    float v = (xp[c] - mean) * inv;
    if (gamma) v *= gamma[c];
    if (beta) v += beta[c];
    out[row * C + c] = v;
  }
}

// one block per row, three passes like the reference kernel (softmax.cu:24-112): max, exp + sum, normalise
// causal_tq > 0: row r is query (r % causal_tq) and sees keys 0..query only - the additive triu(-inf, k=1) mask of CLIP
// (vae/encoder.py:79); masked entries come out as exact zeros, as exp(-inf) does in the reference
__global__ void __launch_bounds__(256) softmax_rows_f32_kernel(float* __restrict__ x, int cols, int causal_tq) {
  tf::pdl_prologue();
  __shared__ float red[8];
  float* __restrict__ xp = x + (long)blockIdx.x * cols;
  const int valid = causal_tq > 0 ? min(cols, (int)(blockIdx.x % causal_tq) + 1) : cols;
  float m = -INFINITY;
  for (int c = threadIdx.x; c < valid; c += blockDim.x) m = fmaxf(m, xp[c]);
  m = block_max(m, red);
  float s = 0.f;
  for (int c = threadIdx.x; c < valid; c += blockDim.x) {
    const float e = expf(xp[c] - m);
    xp[c] = e;
    s += e;
  }
  s = block_sum(s, red);
  for (int c = threadIdx.x; c < cols; c += blockDim.x) xp[c] = c < valid ? xp[c] / s : 0.f;
}

// out[r, :] = table[ids[r], :] + pos[r % T, :]   (ff/embedding.py:15-23 as a row lookup, vae/encoder.py:66-70)
__global__ void embedding_f32_kernel(const int* __restrict__ ids, const float* __restrict__ table,
                                     const float* __restrict__ pos, float* __restrict__ out, int rows, int T, int E, int vocab) {
  tf::pdl_prologue();
  const long total = (long)rows * E;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int r = (int)(i / E), e = (int)(i - (long)r * E);
    const int id = min(max(ids[r], 0), vocab - 1);
    float v = table[(long)id * E + e];
    if (pos) v += pos[(long)(r % T) * E + e];
    out[i] = v;
  }
}

__global__ void unary_f32_kernel(const float* __restrict__ x, float* __restrict__ out, long n, int op) {
  tf::pdl_prologue();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float v = x[i];
    float r;
    switch (op) {
      case 0: r = sigmoid_exact(v); break;                      // tensor.py:65-66
      case 1: r = v * sigmoid_exact(v); break;                  // tensor.py:68-70, 84-86
      case 2: r = 0.5f * v * (1.f + tanhf(v * 0.7978845608f * (1.f + 0.044715f * v * v))); break;  // tensor.py:81-82
      default: r = v * sigmoid_exact(v * 1.702f); break;        // tensor.py:76-78
    }
    out[i] = r;
  }
}

// out[m][j] = y[m][j] * gelu_tanh(y[m][H + j])   (ff/nn.py:10-12: first half = value, second half = gate)
__global__ void geglu_f32_kernel(const float* __restrict__ y, long ldy, float* __restrict__ out, long M, int H) {
  tf::pdl_prologue();
  const long total = M * H;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long m = i / H;
    const int j = (int)(i - m * H);
    const float a = y[m * ldy + j], g = y[m * ldy + H + j];
    out[i] = a * (0.5f * g * (1.f + tanhf(g * 0.7978845608f * (1.f + 0.044715f * g * g))));
  }
}

// e_t = u + g * (c - u)   (variants/sd.py:44-45)
__global__ void cfg_combine_f32_kernel(const float* __restrict__ u, const float* __restrict__ c, float g,
                                       float* __restrict__ out, long n) {
  tf::pdl_prologue();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    out[i] = u[i] + g * (c[i] - u[i]);
}

int grid_for(long n, int threads) {
  long b = (n + threads - 1) / threads;
  const long cap = (long)tf_num_sms() * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}
}  // namespace

extern "C" int tf_gemm_f32(const float* A, long long lda, const float* W, long long ldw, int w_kn, const float* bias,
                           const float* residual, long long ldr, float* out, long long ldc, int M, int N, int K, float alpha,
                           int out_nchw_hw, int batch_outer, int batch_inner, const long long* batch_strides, void* stream) {
  TF_CHECK_ARG(A && W && out && M > 0 && N > 0 && K > 0, "tf_gemm_f32: null pointer or empty shape");
  TF_CHECK_ARG(batch_outer >= 1 && batch_inner >= 1 && (batch_outer * batch_inner == 1 || batch_strides),
               "tf_gemm_f32: batched call needs batch_strides");
  TF_CHECK_ARG(out_nchw_hw == 0 || (M % out_nchw_hw == 0 && batch_outer * batch_inner == 1),
               "tf_gemm_f32: NCHW output needs M %% HW == 0 and no batching");
  GemmF32 p{};
  p.A = A; p.W = W; p.bias = bias; p.residual = residual; p.out = out;
  p.M = M; p.N = N; p.K = K;
  p.lda = lda; p.ldw = ldw; p.w_kn = w_kn; p.ldc = ldc; p.ldr = ldr;
  p.out_nchw = out_nchw_hw > 0;
  p.rows_per_img = out_nchw_hw;
  p.alpha = alpha;
  p.inner = batch_inner;
  if (batch_strides) {
    p.sAo = batch_strides[0]; p.sAi = batch_strides[1]; p.sWo = batch_strides[2];
    p.sWi = batch_strides[3]; p.sCo = batch_strides[4]; p.sCi = batch_strides[5];
  }
  return launch_gemm(p, batch_outer * batch_inner, (cudaStream_t)stream);
}

extern "C" int tf_conv2d_nchw_f32(const float* x, const float* w, const float* bias, const float* bias_img,
                                  const float* residual, float* out, int NI, int C, int H, int Wd, int O, int R, int S,
                                  int stride, int pad, int out_tokens, void* stream) {
  TF_CHECK_ARG(x && w && out && NI > 0 && C > 0 && H > 0 && Wd > 0 && O > 0 && R > 0 && S > 0 && stride > 0 && pad >= 0,
               "tf_conv2d_nchw_f32: bad arguments");
  const int Ho = (H + 2 * pad - R) / stride + 1, Wo = (Wd + 2 * pad - S) / stride + 1;
  TF_CHECK_ARG(Ho > 0 && Wo > 0, "tf_conv2d_nchw_f32: empty output");
  GemmF32 p{};
  p.A = x; p.W = w; p.bias = bias; p.bias_img = bias_img; p.residual = residual; p.out = out;
  p.M = NI * Ho * Wo; p.N = O; p.K = C * R * S;
  p.a_conv = 1;
  p.C = C; p.H = H; p.Wd = Wd; p.R = R; p.S = S; p.stride = stride; p.pad = pad; p.Ho = Ho; p.Wo = Wo;
  p.w_kn = 0; p.ldw = p.K;  // OIHW flattened is [O][C*R*S]
  p.out_nchw = out_tokens ? 0 : 1;
  p.ldc = O; p.ldr = O;
  p.rows_per_img = Ho * Wo;
  p.alpha = 1.f;
  p.inner = 1;
  return launch_gemm(p, 1, (cudaStream_t)stream);
}

extern "C" int tf_groupnorm_nchw_f32(const float* x, const float* gamma, const float* beta, float* out, int NI, int C, int HW,
                                     int groups, float eps, int silu, void* stream) {
  TF_CHECK_ARG(x && out && NI > 0 && C > 0 && HW > 0 && groups > 0 && C % groups == 0, "tf_groupnorm_nchw_f32: bad arguments");
  TF_LAUNCH(groupnorm_nchw_f32_kernel, NI * groups, 512, 0, (cudaStream_t)stream, x, gamma, beta, out, C, HW, groups, eps, silu);
  TF_LAUNCH_CHECK();
  tf_launch_count_add(1);
  return TF_OK;
}

extern "C" int tf_layernorm_f32(const float* x, const float* gamma, const float* beta, float* out, long long rows, int C,
                                float eps, void* stream) {
  TF_CHECK_ARG(x && out && rows > 0 && C > 0, "tf_layernorm_f32: bad arguments");
  TF_LAUNCH(layernorm_f32_kernel, (unsigned)((rows + 3) / 4), 128, 0, (cudaStream_t)stream, x, gamma, beta, out, (long)rows, C, eps);
  TF_LAUNCH_CHECK();
  tf_launch_count_add(1);
  return TF_OK;
}

extern "C" int tf_embedding_f32(const int* ids, const float* table, const float* pos_table, float* out, int rows, int T,
                                int E, int vocab, void* stream) {
  TF_CHECK_ARG(ids && table && out && rows > 0 && T > 0 && E > 0 && vocab > 0, "tf_embedding_f32: bad arguments");
  TF_LAUNCH(embedding_f32_kernel, grid_for((long)rows * E, 256), 256, 0, (cudaStream_t)stream, ids, table, pos_table, out, rows,
            T, E, vocab);
  TF_LAUNCH_CHECK();
  tf_launch_count_add(1);
  return TF_OK;
}

extern "C" int tf_softmax_rows_f32(float* x, long long rows, int cols, int causal_tq, void* stream) {
  TF_CHECK_ARG(x && rows > 0 && rows < (1ll << 31) && cols > 0 && causal_tq >= 0, "tf_softmax_rows_f32: bad arguments");
  TF_LAUNCH(softmax_rows_f32_kernel, (unsigned)rows, 256, 0, (cudaStream_t)stream, x, cols, causal_tq);
  TF_LAUNCH_CHECK();
  tf_launch_count_add(1);
  return TF_OK;
}

extern "C" int tf_unary_f32(const float* x, float* out, long long n, int op, void* stream) {
  TF_CHECK_ARG(x && out && n >= 0 && op >= 0 && op <= 3, "tf_unary_f32: bad arguments");
  if (n == 0) return TF_OK;
  TF_LAUNCH(unary_f32_kernel, grid_for(n, 256), 256, 0, (cudaStream_t)stream, x, out, (long)n, op);
  TF_LAUNCH_CHECK();
  tf_launch_count_add(1);
  return TF_OK;
}

extern "C" int tf_cfg_combine_f32(const float* uncond, const float* cond, float guidance, float* out, long long n,
                                  void* stream) {
  TF_CHECK_ARG(uncond && cond && out && n > 0, "tf_cfg_combine_f32: bad arguments");
  TF_LAUNCH(cfg_combine_f32_kernel, grid_for(n, 256), 256, 0, (cudaStream_t)stream, uncond, cond, guidance, out, (long)n);
  TF_LAUNCH_CHECK();
  tf_launch_count_add(1);
  return TF_OK;
}

extern "C" int tf_geglu_f32(const float* y, long long ldy, float* out, long long M, int H, void* stream) {
  TF_CHECK_ARG(y && out && M > 0 && H > 0 && ldy >= 2 * H, "tf_geglu_f32: bad arguments");
  TF_LAUNCH(geglu_f32_kernel, grid_for(M * H, 256), 256, 0, (cudaStream_t)stream, y, (long)ldy, out, (long)M, H);
  TF_LAUNCH_CHECK();
  tf_launch_count_add(1);
  return TF_OK;
}
