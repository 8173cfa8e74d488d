// tinyfusers_b200 — GroupNorm(+SiLU) and LayerNorm for NHWC / (B,T,C) fp16 activations (HBM-bound).
//
// GroupNorm (reference: tinyfusers/ff/group_norm.py:3-21, followed by Tensor.silu in
// vision/resnet.py:8-10,17-19 and vision/unet.py:45-47): the reference runs >= 8 CuPy elementwise /
// reduction launches and >= 6 full read+write passes. Here: one statistics pass (128-bit loads, register
// partials per channel, fixed-order block and cross-block reduction - no atomics, bit-reproducible) and one
// apply pass that fuses normalise + affine + SiLU and writes the fp16 NHWC tensor the following
// implicit-GEMM conv reads through TMA. Both passes accept a channel slice of a wider tensor
// (pixel stride != C), which is how the UNet's skip concatenation is consumed without a copy.
//
// LayerNorm (reference: tinyfusers/ff/layer_norm.py:8-49, cuDNN graph rebuilt per call): one warp per
// row, row cached in registers, exact two-pass statistics. `interleave` > 1 reproduces the reference's
// NHWC-stride declaration at batch > 1 (SURVEY.md §8 parity note 1): the buffer is viewed as
// (T, C, B) and normalised over C for each (t, b).
#include "tf_common.cuh"
#include "tinyfusers_b200.h"

namespace {

// ------------------------------------------------------------------------------------------------
// GroupNorm statistics, deterministic (no atomics): each block reduces its pixel chunk to per-group
// {sum, sumsq} in a fixed order and writes partial[(chunk*NI + n)*G + g]; gn_finalize_kernel folds the
// chunks in order and emits {mean, rstd}. block = nvec * k threads (nvec = Cs/8): a thread always sees
// the same 8 channels.
// ------------------------------------------------------------------------------------------------
__global__ void gn_stats_kernel(const __half* __restrict__ x, int HW, int Cs, int x_stride, int c_off,
                                int cpg, int G, int pix_per_block, float2* __restrict__ partial) {
  tf::pdl_prologue();  // PDL: let the next kernel start launching, then wait for our producers
  extern __shared__ float gsm[];  // [npl][Cs] sums, [npl][Cs] squares, then [Cs] x 2 channel totals
  const int n = blockIdx.y;
  const int nvec = Cs >> 3;
  const int v = threadIdx.x % nvec;
  const int pl = threadIdx.x / nvec;
  const int npl = blockDim.x / nvec;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
  const __half* base = x + (size_t)n * HW * x_stride + v * 8;
  auto acc = [&](const uint4& vv) {
    tf::Pack16 pk;
    pk.v = vv;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = __half22float2(pk.h2[j]);
      s[2 * j] += f.x; q[2 * j] += f.x * f.x;
      s[2 * j + 1] += f.y; q[2 * j + 1] += f.y * f.y;
    }
  };
  int p = p0 + pl;
  for (; p + 3 * npl < p1; p += 4 * npl) {   // four independent loads in flight per thread; accumulation order unchanged
    uint4 v4[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v4[k] = *reinterpret_cast<const uint4*>(base + (size_t)(p + k * npl) * x_stride);
#pragma unroll
    for (int k = 0; k < 4; ++k) acc(v4[k]);
  }
  for (; p < p1; p += npl) acc(*reinterpret_cast<const uint4*>(base + (size_t)p * x_stride));
  float* ps = gsm;
  float* pq = gsm + npl * Cs;
  float* cs = gsm + 2 * npl * Cs;
  float* cq = cs + Cs;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ps[pl * Cs + v * 8 + j] = s[j];
    pq[pl * Cs + v * 8 + j] = q[j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < Cs; c += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int l = 0; l < npl; ++l) { a += ps[l * Cs + c]; b += pq[l * Cs + c]; }
    cs[c] = a;
    cq[c] = b;
  }
  __syncthreads();
  // groups covered (fully or partly) by this channel slice
  const int g_lo = c_off / cpg, g_hi = (c_off + Cs - 1) / cpg;
  for (int g = g_lo + threadIdx.x; g <= g_hi; g += blockDim.x) {
    const int c_begin = max(g * cpg, c_off) - c_off, c_end = min((g + 1) * cpg, c_off + Cs) - c_off;
    float a = 0.f, b = 0.f;
    for (int c = c_begin; c < c_end; ++c) { a += cs[c]; b += cq[c]; }
    partial[((size_t)blockIdx.x * gridDim.y + n) * G + g] = make_float2(a, b);
  }
}

// stats[(n*G+g)] = {mean, rstd}; one block per image, one WARP per group: lanes stride over the chunk
// partials, then a fixed-pattern xor-shuffle reduction (deterministic). Up to two slice partial arrays.
__global__ void gn_finalize_kernel(const float2* __restrict__ partial_a, const float2* __restrict__ partial_b, int chunks,
                                   int NI, int G, int split_group_lo, int split_group_hi, float inv_count, float eps,
                                   float2* __restrict__ stats) {
  tf::pdl_prologue();  // PDL: let the next kernel start launching, then wait for our producers
  const int n = blockIdx.x, g = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (g >= G) return;
  float a = 0.f, b = 0.f;
  // slice A covers groups [0, split_group_hi], slice B (optional) covers [split_group_lo, G)
  if (g <= split_group_hi)
    for (int c = lane; c < chunks; c += 32) { float2 t = partial_a[((size_t)c * NI + n) * G + g]; a += t.x; b += t.y; }
  if (partial_b && g >= split_group_lo)
    for (int c = lane; c < chunks; c += 32) { float2 t = partial_b[((size_t)c * NI + n) * G + g]; a += t.x; b += t.y; }
  a = tf::warp_sum(a);
  b = tf::warp_sum(b);
  if (lane == 0) {
    const float mean = a * inv_count;
    const float var = fmaxf(b * inv_count - mean * mean, 0.f);
    stats[(size_t)n * G + g] = make_float2(mean, rsqrtf(var + eps));
  }
}

// ------------------------------------------------------------------------------------------------
// GroupNorm apply: out = act((x - mean) * rstd * gamma + beta), act = SiLU or identity
// ------------------------------------------------------------------------------------------------
__global__ void gn_apply_kernel(const __half* __restrict__ x, int HW, int Cs, int x_stride, int c_off,
                                int cpg, int G, int pix_per_block, const float2* __restrict__ stats,
                                const float* __restrict__ gamma, const float* __restrict__ beta, int silu,
                                __half* __restrict__ out, int out_stride) {
  tf::pdl_prologue();  // PDL: let the next kernel start launching, then wait for our producers
  const int n = blockIdx.y;
  const int nvec = Cs >> 3;
  const int v = threadIdx.x % nvec;
  const int pl = threadIdx.x / nvec;
  const int npl = blockDim.x / nvec;
  const int c0 = c_off + v * 8;
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + j;
    const float2 mr = stats[(size_t)n * G + c / cpg];
    const float ga = gamma ? gamma[c] : 1.f;
    const float be = beta ? beta[c] : 0.f;
    a[j] = mr.y * ga;
    b[j] = be - mr.x * mr.y * ga;
  }
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  const __half* xin = x + (size_t)n * HW * x_stride + v * 8;
  __half* o = out + (size_t)n * HW * out_stride + c0;
  auto apply = [&](const uint4& vv, int p) {
    tf::Pack16 pk, r;
    pk.v = vv;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = __half22float2(pk.h2[j]);
      float y0 = f.x * a[2 * j] + b[2 * j];
      float y1 = f.y * a[2 * j + 1] + b[2 * j + 1];
      if (silu) { y0 = tf::silu_f(y0); y1 = tf::silu_f(y1); }
      r.h2[j] = __floats2half2_rn(y0, y1);
    }
    *reinterpret_cast<uint4*>(o + (size_t)p * out_stride) = r.v;
  };
  int p = p0 + pl;
  for (; p + 3 * npl < p1; p += 4 * npl) {   // long pixel chunks (VAE-size images): four independent loads in flight per thread
    uint4 v4[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v4[k] = *reinterpret_cast<const uint4*>(xin + (size_t)(p + k * npl) * x_stride);
#pragma unroll
    for (int k = 0; k < 4; ++k) apply(v4[k], p + k * npl);
  }
  for (; p < p1; p += npl) apply(*reinterpret_cast<const uint4*>(xin + (size_t)p * x_stride), p);
}

// ------------------------------------------------------------------------------------------------
// Single-launch GroupNorm (+SiLU) for inputs whose producer (tf_gemm_gn_f16 / tf_conv2d_nhwc_gn_f16) already left
// per-slot statistics: stats[image][slot = 32 rows][unit] = {sum, sumsq}. Every block first folds the slots of
// its image (one warp per unit, lanes stride over slots, fixed xor-shuffle order), then units -> groups, then
// normalises its pixel chunk. Up to two sources (channel concatenation), each with its own unit size.
// ------------------------------------------------------------------------------------------------
__global__ void gn_fused_apply_kernel(const __half* __restrict__ x1, int stride1, int C1, const float2* __restrict__ st1,
                                      int unit1, const __half* __restrict__ x2, int stride2, int C2,
                                      const float2* __restrict__ st2, int unit2, int HW, int cpg, int G, int pix_per_block,
                                      const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                      float inv_count, int silu, __half* __restrict__ out, int out_stride) {
  tf::pdl_trigger();
  extern __shared__ float2 fsm[];   // [units] unit totals, [G] {mean, rstd}, [SG][units] partials of the fold
  const int n = blockIdx.y;
  const int slots = HW >> 5;
  const int units1 = C1 / unit1, units2 = x2 ? C2 / unit2 : 0;
  const int units = units1 + units2;
  // ---- this thread's channels / pixels; everything that can be fetched early is: the affine parameters do not
  // depend on the producer kernels (loaded before the dependency wait), the pixels only on the producer's output
  // (loaded before the statistics fold), so the three global round trips of the kernel overlap ----
  const int nvec = (C1 + C2) >> 3;
  const int v = threadIdx.x % nvec;
  const int pl = threadIdx.x / nvec;
  const int npl = blockDim.x / nvec;
  const bool worker = pl < npl;
  const int c0 = v * 8;
  float gg[8], bb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { gg[j] = 1.f; bb[j] = 0.f; }
  if (worker && gamma) {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
    gg[0] = g0.x; gg[1] = g0.y; gg[2] = g0.z; gg[3] = g0.w; gg[4] = g1.x; gg[5] = g1.y; gg[6] = g1.z; gg[7] = g1.w;
  }
  if (worker && beta) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c0)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c0 + 4));
    bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w; bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
  }
  tf::pdl_wait();
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  const __half* xin;
  int xs;
  if (c0 < C1) { xin = x1 + (size_t)n * HW * stride1 + c0; xs = stride1; }
  else { xin = x2 + (size_t)n * HW * stride2 + (c0 - C1); xs = stride2; }
  constexpr int kPre = 3;
  uint4 pre[kPre];
#pragma unroll
  for (int i = 0; i < kPre; ++i) {
    const int p = p0 + pl + i * npl;
    pre[i] = make_uint4(0u, 0u, 0u, 0u);
    if (worker && p < p1) pre[i] = *reinterpret_cast<const uint4*>(xin + (size_t)p * xs);
  }
  // ---- fold: thread (sg, u) sums slots sg, sg + SG, ... of unit u (independent, coalesced loads: no shuffle in the
  // loop, so they pipeline), then thread u adds the SG partials in a fixed order ----
  float2* part = fsm + units + G;
  const int SG = min((int)blockDim.x / units, slots);
  {
    const int u = threadIdx.x % units, sg = threadIdx.x / units;
    if (sg < SG) {
      const float2* src;
      int ld;
      if (u < units1) { src = st1 + (size_t)n * slots * units1 + u; ld = units1; }
      else { src = st2 + (size_t)n * slots * units2 + (u - units1); ld = units2; }
      float a = 0.f, b = 0.f;
#pragma unroll 4
      for (int j = sg; j < slots; j += SG) {
        const float2 t = __ldg(src + (size_t)j * ld);
        a += t.x; b += t.y;
      }
      part[sg * units + u] = make_float2(a, b);
    }
  }
  __syncthreads();
  for (int u = threadIdx.x; u < units; u += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int sg = 0; sg < SG; ++sg) { a += part[sg * units + u].x; b += part[sg * units + u].y; }
    fsm[u] = make_float2(a, b);
  }
  __syncthreads();
  float2* gst = fsm + units;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    float a = 0.f, b = 0.f;
    int c = g * cpg;
    const int c_end = c + cpg;
    while (c < c_end) {           // unit boundaries are aligned with group and source boundaries (host-checked)
      int u, step;
      if (c < C1) { u = c / unit1; step = unit1; }
      else { u = units1 + (c - C1) / unit2; step = unit2; }
      a += fsm[u].x; b += fsm[u].y;
      c += step;
    }
    const float mean = a * inv_count;
    const float var = fmaxf(b * inv_count - mean * mean, 0.f);
    gst[g] = make_float2(mean, rsqrtf(var + eps));
  }
  __syncthreads();
  if (!worker) return;
  float a8[8], b8[8];
  {
    int g = c0 / cpg, left = (g + 1) * cpg - c0;   // channels left in group g (one division per thread)
    float2 mr = gst[g];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (left == 0) { ++g; mr = gst[min(g, G - 1)]; left = cpg; }
      --left;
      a8[j] = mr.y * gg[j];
      b8[j] = bb[j] - mr.x * mr.y * gg[j];
    }
  }
  __half* o = out + (size_t)n * HW * out_stride + c0;
  int i = 0;
  for (int p = p0 + pl; p < p1; p += npl, ++i) {
    tf::Pack16 pk;
    if (i < kPre) {
#pragma unroll
      for (int k = 0; k < kPre; ++k)
        if (k == i) pk.v = pre[k];
    } else {
      pk.v = *reinterpret_cast<const uint4*>(xin + (size_t)p * xs);
    }
    tf::Pack16 r;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = __half22float2(pk.h2[j]);
      float y0 = f.x * a8[2 * j] + b8[2 * j];
      float y1 = f.y * a8[2 * j + 1] + b8[2 * j + 1];
      if (silu) { y0 = tf::silu_f(y0); y1 = tf::silu_f(y1); }
      r.h2[j] = __floats2half2_rn(y0, y1);
    }
    *reinterpret_cast<uint4*>(o + (size_t)p * out_stride) = r.v;
  }
}

static int gn_block_threads(int Cs) {
  const int nvec = Cs / 8;
  int k = 512 / nvec;
  const int k_smem = (40 * 1024 / 4 - 2 * Cs) / (2 * Cs);  // statistics kernel shared-memory budget
  if (k > k_smem) k = k_smem;
  if (k < 1) k = 1;
  return nvec * k;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row; IL = interleave factor (1 canonical; >1 reference stride quirk)
// memory index of logical element (row r = (t, b), channel c):  t*C*IL + c*IL + b
// ------------------------------------------------------------------------------------------------
template <int MAXV>  // max 8-half vectors per lane
__global__ void ln_rows_kernel(const __half* __restrict__ x, __half* __restrict__ out, int rows, int C,
                               const float* __restrict__ gamma, const float* __restrict__ beta, float eps) {
  tf::pdl_prologue();  // PDL: let the next kernel start launching, then wait for our producers
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int nvec = C >> 3;
  const __half* xr = x + (size_t)warp * C;
  tf::Pack16 pk[MAXV];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      pk[i].v = *reinterpret_cast<const uint4*>(xr + v * 8);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = __half22float2(pk[i].h2[j]);
        sum += f.x + f.y;
      }
    }
  }
  const float mean = tf::warp_sum(sum) / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = __half22float2(pk[i].h2[j]);
        sq += (f.x - mean) * (f.x - mean) + (f.y - mean) * (f.y - mean);
      }
    }
  }
  const float rstd = rsqrtf(tf::warp_sum(sq) / (float)C + eps);
  __half* orow = out + (size_t)warp * C;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8 + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + v * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + v * 8 + 4));
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      tf::Pack16 r;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = __half22float2(pk[i].h2[j]);
        r.h2[j] = __floats2half2_rn((f.x - mean) * rstd * g[2 * j] + b[2 * j],
                                    (f.y - mean) * rstd * g[2 * j + 1] + b[2 * j + 1]);
      }
      *reinterpret_cast<uint4*>(orow + v * 8) = r.v;
    }
  }
}

// Block-per-chunk-row LayerNorm for interleave IL in {1,2,4,8}: the chunk row is C*IL contiguous halfs; element
// p belongs to group (p % IL) and channel (p / IL). 128-bit loads, the row stays in registers, exact two-pass
// statistics, fixed-order block reduction. Used when rows are long or few (one warp per row is latency-bound).
template <int IL, int THREADS, int MAXV, int ROWS>
__global__ void __launch_bounds__(THREADS * ROWS)
ln_block_kernel(const __half* __restrict__ x, __half* __restrict__ out, int nrows, int C, const float* __restrict__ gamma,
                const float* __restrict__ beta, float eps) {
  tf::pdl_trigger();
  __shared__ float red[2][ROWS][THREADS / 32][IL];   // ROWS rows per block, THREADS threads each (whole warps)
  const int rib = threadIdx.x / THREADS, tid = threadIdx.x % THREADS;
  const int row = blockIdx.x * ROWS + rib;
  const bool live = row < nrows;
  const int nvec = (C * IL) >> 3;
  // the affine parameters do not depend on the producer kernels: fetch them before the dependency wait (IL == 1)
  float4 gpre[MAXV][2], bpre[MAXV][2];
  if (IL == 1) {
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int v = tid + i * THREADS;
      if (v < nvec) {
        gpre[i][0] = __ldg(reinterpret_cast<const float4*>(gamma + v * 8));
        gpre[i][1] = __ldg(reinterpret_cast<const float4*>(gamma + v * 8 + 4));
        bpre[i][0] = __ldg(reinterpret_cast<const float4*>(beta + v * 8));
        bpre[i][1] = __ldg(reinterpret_cast<const float4*>(beta + v * 8 + 4));
      }
    }
  }
  tf::pdl_wait();
  const __half* xr = x + (size_t)row * C * IL;
  const int warp = tid >> 5, lane = tid & 31;
  tf::Pack16 pk[MAXV];
  float s[IL];
#pragma unroll
  for (int b = 0; b < IL; ++b) s[b] = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = tid + i * THREADS;
    if (live && v < nvec) {
      pk[i].v = *reinterpret_cast<const uint4*>(xr + v * 8);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(pk[i].h2[j]);
        s[(2 * j) % IL] += f.x;
        s[(2 * j + 1) % IL] += f.y;
      }
    }
  }
#pragma unroll
  for (int b = 0; b < IL; ++b) {
    const float w = tf::warp_sum(s[b]);
    if (lane == 0) red[0][rib][warp][b] = w;
  }
  __syncthreads();
  float mean[IL];
#pragma unroll
  for (int b = 0; b < IL; ++b) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) t += red[0][rib][w][b];
    mean[b] = t / (float)C;
    s[b] = 0.f;
  }
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = tid + i * THREADS;
    if (live && v < nvec) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(pk[i].h2[j]);
        const float d0 = f.x - mean[(2 * j) % IL], d1 = f.y - mean[(2 * j + 1) % IL];
        s[(2 * j) % IL] += d0 * d0;
        s[(2 * j + 1) % IL] += d1 * d1;
      }
    }
  }
#pragma unroll
  for (int b = 0; b < IL; ++b) {
    const float w = tf::warp_sum(s[b]);
    if (lane == 0) red[1][rib][warp][b] = w;
  }
  __syncthreads();
  float rstd[IL];
#pragma unroll
  for (int b = 0; b < IL; ++b) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) t += red[1][rib][w][b];
    rstd[b] = rsqrtf(t / (float)C + eps);
  }
  __half* orow = out + (size_t)row * C * IL;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = tid + i * THREADS;
    if (live && v < nvec) {
      tf::Pack16 r;
      const int c0 = (v * 8) / IL;  // first channel covered by this vector (8 / IL channels)
      if (IL == 1) {
        const float4 g0 = gpre[i][0], g1 = gpre[i][1], b0 = bpre[i][0], b1 = bpre[i][1];
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(pk[i].h2[j]);
          const float a0 = rstd[0] * gg[2 * j], a1 = rstd[0] * gg[2 * j + 1];
          r.h2[j] = __floats2half2_rn((f.x - mean[0]) * a0 + bb[2 * j], (f.y - mean[0]) * a1 + bb[2 * j + 1]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = c0 + j / IL;
          const float y = (__half2float(pk[i].h[j]) - mean[j % IL]) * rstd[j % IL] * __ldg(gamma + c) + __ldg(beta + c);
          r.h[j] = __float2half_rn(y);
        }
      }
      *reinterpret_cast<uint4*>(orow + v * 8) = r.v;
    }
  }
}

template <int IL>
static bool launch_ln_block(const __half* x, __half* out, int chunk_rows, int C, const float* gamma, const float* beta,
                            float eps, cudaStream_t stream) {
  const int nvec = C * IL / 8;
  if ((C * IL) % 8 != 0) return false;
  // 256-thread blocks: several short rows per block keep enough warps resident to hide the load latency
  if (nvec <= 64) TF_LAUNCH((ln_block_kernel<IL, 32, 2, 8>), ceil_div_i(chunk_rows, 8), 256, 0, stream, x, out, chunk_rows, C, gamma, beta, eps);
  else if (nvec <= 128) TF_LAUNCH((ln_block_kernel<IL, 64, 2, 4>), ceil_div_i(chunk_rows, 4), 256, 0, stream, x, out, chunk_rows, C, gamma, beta, eps);
  else if (nvec <= 256) TF_LAUNCH((ln_block_kernel<IL, 128, 2, 2>), ceil_div_i(chunk_rows, 2), 256, 0, stream, x, out, chunk_rows, C, gamma, beta, eps);
  else if (nvec <= 512) TF_LAUNCH((ln_block_kernel<IL, 128, 4, 1>), chunk_rows, 128, 0, stream, x, out, chunk_rows, C, gamma, beta, eps);
  else if (nvec <= 1024) TF_LAUNCH((ln_block_kernel<IL, 256, 4, 1>), chunk_rows, 256, 0, stream, x, out, chunk_rows, C, gamma, beta, eps);
  else return false;
  return true;
}

// generic interleave: one warp per (t, b); strided scalar access (correct for any IL, not tuned)
__global__ void ln_il_generic_kernel(const __half* __restrict__ x, __half* __restrict__ out, int trow, int C,
                                     int IL, const float* __restrict__ gamma, const float* __restrict__ beta,
                                     float eps) {
  tf::pdl_prologue();  // PDL: let the next kernel start launching, then wait for our producers
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= trow * IL) return;
  const int t = warp / IL, b = warp % IL;
  const __half* xr = x + (size_t)t * C * IL + b;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += __half2float(xr[(size_t)c * IL]);
  const float mean = tf::warp_sum(s) / (float)C;
  float q = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float d = __half2float(xr[(size_t)c * IL]) - mean;
    q += d * d;
  }
  const float rstd = rsqrtf(tf::warp_sum(q) / (float)C + eps);
  __half* orow = out + (size_t)t * C * IL + b;
  for (int c = lane; c < C; c += 32)
    orow[(size_t)c * IL] = __float2half_rn((__half2float(xr[(size_t)c * IL]) - mean) * rstd * gamma[c] + beta[c]);
}

}  // namespace

extern "C" size_t tf_groupnorm_workspace_bytes(int NI, int groups) {
  // 2 slices x chunks x NI x G float2 partials + NI x G float2 {mean, rstd}; chunks <= 2*SMs + 1
  const size_t chunks = (size_t)2 * tf_num_sms() + 1;
  return sizeof(float2) * ((size_t)2 * chunks * NI * groups + (size_t)NI * groups);
}

extern "C" int tf_groupnorm_nhwc_f16(const void* x, int x_pixel_stride, int Cx, const void* x2,
                                     int x2_pixel_stride, int Cx2, void* out, int out_pixel_stride, int NI,
                                     int HW, int groups, const float* gamma, const float* beta, float eps,
                                     int apply_silu, float* stats_ws, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TF_CHECK_ARG(x && out && stats_ws, "tf_groupnorm_nhwc_f16: null pointer");
  const int C = Cx + (x2 ? Cx2 : 0);
  TF_CHECK_ARG(NI > 0 && HW > 0 && groups > 0 && groups <= 32 && C % groups == 0,
               "tf_groupnorm_nhwc_f16: bad dims (C=%d groups=%d)", C, groups);
  TF_CHECK_ARG(Cx % 8 == 0 && (!x2 || Cx2 % 8 == 0) && x_pixel_stride % 8 == 0 && out_pixel_stride % 8 == 0 &&
                   (!x2 || x2_pixel_stride % 8 == 0),
               "tf_groupnorm_nhwc_f16: channels and strides must be multiples of 8");
  TF_CHECK_ARG(Cx / 8 <= 1024 && (!x2 || Cx2 / 8 <= 1024), "tf_groupnorm_nhwc_f16: too many channels");
  const int cpg = C / groups;
  const int sms = tf_num_sms();
  int chunks = (2 * sms + NI - 1) / NI;
  if (chunks > HW) chunks = HW;
  const int ppb = ceil_div_i(HW, chunks);
  chunks = ceil_div_i(HW, ppb);
  const dim3 grid(chunks, NI);
  const float inv_count = 1.0f / ((float)cpg * (float)HW);
  const __half* xs[2] = {reinterpret_cast<const __half*>(x), reinterpret_cast<const __half*>(x2)};
  const int cs[2] = {Cx, Cx2}, st[2] = {x_pixel_stride, x2_pixel_stride}, off[2] = {0, Cx};
  const int nsrc = x2 ? 2 : 1;
  float2* partial[2];
  partial[0] = reinterpret_cast<float2*>(stats_ws);
  partial[1] = partial[0] + (size_t)chunks * NI * groups;
  float2* stats = partial[1] + (size_t)chunks * NI * groups;
  for (int i = 0; i < nsrc; ++i) {
    const int threads = gn_block_threads(cs[i]);
    const int npl = threads / (cs[i] / 8);
    const size_t smem = sizeof(float) * ((size_t)2 * npl * cs[i] + 2 * cs[i]);
    TF_CHECK_ARG(smem <= 48 * 1024, "tf_groupnorm_nhwc_f16: channel slice of %d too wide for the statistics kernel", cs[i]);
    TF_LAUNCH(gn_stats_kernel, grid, threads, smem, stream, xs[i], HW, cs[i], st[i], off[i], cpg, groups, ppb, partial[i]);
    TF_LAUNCH_CHECK();
  }
  TF_LAUNCH(gn_finalize_kernel, NI, 32 * groups, 0, stream, partial[0], nsrc == 2 ? partial[1] : nullptr, chunks, NI, groups,
                                            nsrc == 2 ? Cx / cpg : 0, (Cx - 1) / cpg, inv_count, eps, stats);
  TF_LAUNCH_CHECK();
  for (int i = 0; i < nsrc; ++i) {
    TF_LAUNCH(gn_apply_kernel, grid, gn_block_threads(cs[i]), 0, stream, 
        xs[i], HW, cs[i], st[i], off[i], cpg, groups, ppb, stats, gamma, beta, apply_silu,
        reinterpret_cast<__half*>(out), out_pixel_stride);
    TF_LAUNCH_CHECK();
  }
  tf_launch_count_add(2 * nsrc + 1);
  return TF_OK;
}

extern "C" int tf_groupnorm_fused_nhwc_f16(const void* x, int x_pixel_stride, int Cx, const void* x_stats, int x_unit,
                                           const void* x2, int x2_pixel_stride, int Cx2, const void* x2_stats, int x2_unit,
                                           void* out, int out_pixel_stride, int NI, int HW, int groups, const float* gamma,
                                           const float* beta, float eps, int apply_silu, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TF_CHECK_ARG(x && out && x_stats && (!x2 || x2_stats), "tf_groupnorm_fused_nhwc_f16: null pointer");
  if (!x2) { Cx2 = 0; x2_unit = 1; }
  const int C = Cx + Cx2;
  TF_CHECK_ARG(NI > 0 && HW > 0 && HW % 32 == 0 && groups > 0 && groups <= 64 && C % groups == 0,
               "tf_groupnorm_fused_nhwc_f16: bad dims (C=%d groups=%d HW=%d)", C, groups, HW);
  TF_CHECK_ARG(Cx % 8 == 0 && Cx2 % 8 == 0 && x_pixel_stride % 8 == 0 && out_pixel_stride % 8 == 0 &&
                   (!x2 || x2_pixel_stride % 8 == 0) && C / 8 <= 1024,
               "tf_groupnorm_fused_nhwc_f16: channels and strides must be multiples of 8");
  const int cpg = C / groups;
  TF_CHECK_ARG(x_unit > 0 && Cx % x_unit == 0 && cpg % x_unit == 0 &&
                   (!x2 || (x2_unit > 0 && Cx2 % x2_unit == 0 && cpg % x2_unit == 0 && Cx % x2_unit == 0)),
               "tf_groupnorm_fused_nhwc_f16: statistics units (%d, %d) do not tile groups of %d channels", x_unit, x2_unit, cpg);
  const int sms = tf_num_sms();
  int chunks = (2 * sms + NI - 1) / NI;
  if (chunks > HW) chunks = HW;
  const int ppb = ceil_div_i(HW, chunks);
  chunks = ceil_div_i(HW, ppb);
  int threads = gn_block_threads(C);
  threads = (threads + 31) / 32 * 32;   // whole warps for the fold; surplus threads skip the apply loop
  const int units = Cx / x_unit + (x2 ? Cx2 / x2_unit : 0);
  TF_CHECK_ARG(units <= threads, "tf_groupnorm_fused_nhwc_f16: %d statistics units exceed the block size", units);
  const size_t smem = sizeof(float2) * ((size_t)units + groups + (size_t)(threads / units) * units);
  TF_LAUNCH(gn_fused_apply_kernel, dim3(chunks, NI), threads, smem, stream, reinterpret_cast<const __half*>(x), x_pixel_stride,
            Cx, reinterpret_cast<const float2*>(x_stats), x_unit, reinterpret_cast<const __half*>(x2), x2_pixel_stride, Cx2,
            reinterpret_cast<const float2*>(x2_stats), x2_unit, HW, cpg, groups, ppb, gamma, beta, eps,
            1.0f / ((float)cpg * (float)HW), apply_silu, reinterpret_cast<__half*>(out), out_pixel_stride);
  TF_LAUNCH_CHECK();
  tf_launch_count_add(1);
  return TF_OK;
}

extern "C" int tf_layernorm_f16(const void* x, void* out, int rows, int C, const float* gamma, const float* beta,
                                float eps, int interleave, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TF_CHECK_ARG(x && out && gamma && beta, "tf_layernorm_f16: null pointer");
  TF_CHECK_ARG(rows > 0 && C > 0 && interleave >= 1 && rows % interleave == 0,
               "tf_layernorm_f16: bad dims rows=%d C=%d interleave=%d", rows, C, interleave);
  const int threads = 256, wpb = threads / 32;
  const __half* xh = reinterpret_cast<const __half*>(x);
  __half* oh = reinterpret_cast<__half*>(out);
  const int chunk_rows = rows / interleave;
  bool done = false;
  if (interleave == 1) done = launch_ln_block<1>(xh, oh, chunk_rows, C, gamma, beta, eps, stream);
  else if (interleave == 2) done = launch_ln_block<2>(xh, oh, chunk_rows, C, gamma, beta, eps, stream);
  else if (interleave == 4) done = launch_ln_block<4>(xh, oh, chunk_rows, C, gamma, beta, eps, stream);
  else if (interleave == 8) done = launch_ln_block<8>(xh, oh, chunk_rows, C, gamma, beta, eps, stream);
  if (!done) {
    if (interleave == 1) {
      TF_CHECK_ARG(C % 8 == 0 && C <= 8 * 32 * 8, "tf_layernorm_f16: C must be a multiple of 8 and <= 2048");
      TF_LAUNCH((ln_rows_kernel<8>), ceil_div_i(rows, wpb), threads, 0, stream, xh, oh, rows, C, gamma, beta, eps);
    } else {
      TF_LAUNCH(ln_il_generic_kernel, ceil_div_i(rows, wpb), threads, 0, stream, xh, oh, chunk_rows, C, interleave, gamma,
                                                                         beta, eps);
    }
  }
  TF_LAUNCH_CHECK();
  tf_launch_count_add(1);
  return TF_OK;
}
