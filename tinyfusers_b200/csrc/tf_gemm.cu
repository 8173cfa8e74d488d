// tinyfusers_b200 — tcgen05 GEMM and NHWC implicit-GEMM convolution for sm_100a.
//
// One persistent, warp-specialised kernel serves both:
//   * Linear / 1x1 conv :  D[M,N] = A[M,K] · W[N,K]^T          (reference: tinyfusers/ff/linear.py:116-121,
//                                                               tinyfusers/vision/conv2d.py:9-28 with R=S=1)
//   * 3x3 conv (s1/s2)  :  same contraction with K = 9·Cin; the A tile of tap (r,s) is a TMA box of the
//                          NHWC input shifted by (r-1, s-1) — out-of-bounds pixels are zero-filled by the
//                          TMA unit, which *is* the conv padding (reference: tinyfusers/vision/conv2d.py:48-59).
//
// Roles (192 threads): warp 0 = TMA producer, warp 1 = tcgen05.mma issuer, warps 2..5 = epilogue
// (TMEM -> registers -> bias / residual / GEGLU -> global). Accumulators are double-buffered in TMEM
// (2 x 256 columns) so the epilogue of tile i overlaps the main loop of tile i+1.
// Tile = 128 x BN x 64, BN in {16..256 step 16} chosen per shape; optional split-K writes fp32 partials
// that tf_splitk_reduce folds (deep, small-M UNet levels are weight-bandwidth bound at batch 2).
#include "tf_common.cuh"
#include "tinyfusers_b200.h"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;                      // one 128-byte swizzle atom of fp16
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KiB
constexpr int kThreads = 192;
constexpr int kMaxStages = 8;
constexpr int kSmemBudget = 227 * 1024;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kAccStride = 256;  // columns between the two accumulator buffers
// epilogue staging: each epilogue warp owns one 32-row x 32-column block (<= 128 B per row, fp32 worst case)
// written thread-per-row in the TMA swizzle pattern and shipped with one TMA store per chunk, plus a copy of
// the tile's bias slice.
constexpr int kEpiStageBytes = 32 * 128;                    // per warp, 1024-aligned (swizzle atom)
constexpr int kEpiBiasFloats = 256;
constexpr int kEpiBytes = 4 * kEpiStageBytes + 4 * kEpiBiasFloats * 4;  // 20480

struct ConvGeom {
  int H, W, NI;          // OUTPUT height/width, images
  int TW, TH, TN;        // tile extents in output pixels / images; TW*TH*TN == 128
  int tiles_x, tiles_y;  // tiles per image along x / y
  int cblocks;           // Cin / 64
  int cscale;            // input coord = output coord * cscale + tap offset (1: stride 1, 2: stride 2)
  int pad;               // 1 for 3x3
  int ksize;             // 3
  int sbw, sbh;          // store box of one epilogue warp (32 tile rows): sbw x sbh x (32/(sbw*sbh)) pixels
};

struct GemmParams {
  int M, N, K;
  int bn, m_tiles, n_tiles, splits, k_blocks, kb_per_split, stages;
  int is_conv;
  ConvGeom g;
  void* out;
  int ldc;
  const float* bias;
  const __half* residual;
  int ldr;
  float* partial;
  int flags;
  long long* timeline;  // optional debug: per-CTA clock stamps [grid][8] (tf_gemm_set_timeline)
};

// tile-local row (0..127) -> global output row (pixel index for conv), or -1 if padding
__device__ __forceinline__ int tile_row_to_m(const GemmParams& p, int mt, int r) {
  if (!p.is_conv) {
    int m = mt * BM + r;
    return m < p.M ? m : -1;
  }
  const ConvGeom& g = p.g;
  int tx = mt % g.tiles_x;
  int t2 = mt / g.tiles_x;
  int ty = t2 % g.tiles_y;
  int tn = t2 / g.tiles_y;
  int xi = r % g.TW;
  int r2 = r / g.TW;
  int yi = r2 % g.TH;
  int ni = r2 / g.TH;
  int x = tx * g.TW + xi, y = ty * g.TH + yi, n = tn * g.TN + ni;
  if (x >= g.W || y >= g.H || n >= g.NI) return -1;
  return (n * g.H + y) * g.W + x;
}

__global__ void __launch_bounds__(kThreads, 1)
tf_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  tf::pdl_trigger();
  const long long t_entry = clock64();
  const uint32_t raw_u32 = tf::smem_u32(smem_raw);
  const uint32_t smem_base = (raw_u32 + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const uint32_t b_stage_bytes = (uint32_t)p.bn * 128u;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + p.stages * A_STAGE_BYTES;
  const uint32_t epi_base = smem_b + p.stages * b_stage_bytes;
  const uint32_t bar_base = epi_base + kEpiBytes;
  const int S = p.stages;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * S + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * S + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw_u32));

  if (warp == 0 && lane == 0) {
    tf::tma_prefetch_desc(&tmA);
    tf::tma_prefetch_desc(&tmB);
    tf::tma_prefetch_desc(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      tf::mbar_init(full_bar(s), 1);
      tf::mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      tf::mbar_init(tfull_bar(a), 1);
      tf::mbar_init(tempty_bar(a), 128);
    }
    tf::fence_mbar_init();
  }
  if (warp == 2) {
    tf::tmem_alloc(tmem_slot, kTmemCols);
    tf::tmem_relinquish();
  }
  tf::tcgen05_fence_before();
  __syncthreads();
  tf::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  tf::pdl_wait();   // barriers / TMEM / descriptors were set up while the producer kernels drained

  const int total_tiles = p.m_tiles * p.n_tiles * p.splits;
  long long* tl = p.timeline ? p.timeline + (size_t)blockIdx.x * 16 : nullptr;
#define TF_STAMP(i) do { if (tl) tl[i] = clock64(); } while (0)
  if (threadIdx.x == 0) { TF_STAMP(0); if (tl) tl[7] = t_entry; }   // setup done (barriers, TMEM alloc)

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = A_STAGE_BYTES + b_stage_bytes;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int split = t % p.splits;
        const int t1 = t / p.splits;
        const int nt = t1 % p.n_tiles;
        const int mt = t1 / p.n_tiles;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
        int x0 = 0, y0 = 0, n0 = 0;
        if (p.is_conv) {
          const ConvGeom& g = p.g;
          int tx = mt % g.tiles_x;
          int t2 = mt / g.tiles_x;
          x0 = tx * g.TW * g.cscale - g.pad;
          y0 = (t2 % g.tiles_y) * g.TH * g.cscale - g.pad;
          n0 = (t2 / g.tiles_y) * g.TN;
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          tf::mbar_wait(empty_bar(stage), phase ^ 1u);
          tf::mbar_expect_tx(full_bar(stage), tx_bytes);
          const uint32_t a_dst = smem_a + stage * A_STAGE_BYTES;
          const uint32_t b_dst = smem_b + stage * b_stage_bytes;
          if (p.is_conv) {
            const int tap = kb / p.g.cblocks;
            const int cb = kb - tap * p.g.cblocks;
            const int r = tap / p.g.ksize;
            const int s = tap - r * p.g.ksize;
            tf::tma_load_4d(a_dst, &tmA, full_bar(stage), cb * BK, x0 + s, y0 + r, n0);
          } else {
            tf::tma_load_2d(a_dst, &tmA, full_bar(stage), kb * BK, mt * BM);
          }
          tf::tma_load_2d(b_dst, &tmB, full_bar(stage), kb * BK, nt * p.bn);
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = tf::umma_idesc_f16(BM, p.bn);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int split = t % p.splits;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
        tf::mbar_wait(tempty_bar(as), aphase ^ 1u);
        if (t == blockIdx.x) TF_STAMP(1);
        tf::tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + as * kAccStride;
        for (int kb = kb0; kb < kb1; ++kb) {
          tf::mbar_wait(full_bar(stage), phase);
          if (t == blockIdx.x && kb == kb0) TF_STAMP(2);   // first operands landed
          tf::tcgen05_fence_after();
          const uint64_t adesc = tf::umma_desc_sw128_kmajor(smem_a + stage * A_STAGE_BYTES);
          const uint64_t bdesc = tf::umma_desc_sw128_kmajor(smem_b + stage * b_stage_bytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // +32 bytes per UMMA_K step inside the 128-byte swizzle atom (start-address field is >>4)
            tf::umma_f16_ss(tmem_d, adesc + 2u * k, bdesc + 2u * k, idesc,
                            (kb > kb0 || k > 0) ? 1u : 0u);
          }
          tf::umma_commit(empty_bar(stage));  // frees this smem stage once the MMAs have read it
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
        if (t == blockIdx.x) TF_STAMP(3);   // all MMAs of the first tile issued
        tf::umma_commit(tfull_bar(as));  // accumulator complete -> epilogue
        as ^= 1;
        if (as == 0) aphase ^= 1u;
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    // Thread = accumulator row (the 32x32b TMEM load hands every lane its own row). Everything that does not
    // depend on the accumulator is fetched while the main loop runs: the tile's bias slice (-> shared
    // memory, read back as broadcasts) and the first residual chunk (64 contiguous bytes of the thread's
    // row). The accumulator is drained 32 columns at a time: + bias + residual in registers, convert, write
    // the row chunk into this warp's staging block in the TMA swizzle pattern (conflict-free 16-byte
    // stores), then ONE TMA store ships the 32x32 block; rows / columns outside the tensor are clipped by
    // the TMA unit, so there is no per-element address or bounds arithmetic at all.
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const uint32_t stg = epi_base + q * kEpiStageBytes;
    float* bsm = reinterpret_cast<float*>(smem_raw + (epi_base + 4 * kEpiStageBytes - raw_u32)) + q * kEpiBiasFloats;
    const bool partial = p.partial != nullptr;
    const bool geglu = (p.flags & TF_EPI_GEGLU) != 0 && !partial;
    const bool out_f32 = (p.flags & TF_EPI_OUT_F32) != 0 || partial;
    const bool use_bias = p.bias != nullptr && !partial;
    const bool use_res = p.residual != nullptr && !partial;
    // swizzle of 16-byte chunk j in row r (row = lane): fp32 rows are 128 B (SW128), fp16 64 B (SW64), GEGLU 32 B (SW32)
    const uint32_t row_bytes = out_f32 ? 128u : (geglu ? 32u : 64u);
    const uint32_t swz = out_f32 ? (uint32_t)(lane & 7) : (geglu ? (uint32_t)((lane >> 2) & 1) : (uint32_t)((lane >> 1) & 3));
    int as = 0;
    uint32_t aphase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int split = t % p.splits;
      const int t1 = t / p.splits;
      const int nt = t1 % p.n_tiles;
      const int mt = t1 / p.n_tiles;
      const int n_tile = nt * p.bn;
      const int m_own = tile_row_to_m(p, mt, row);
      // store coordinates of this warp's 32-row block
      int sc1, sc2 = 0, sc3 = 0;
      if (p.is_conv) {
        const ConvGeom& g = p.g;
        const int tx = mt % g.tiles_x, t2 = mt / g.tiles_x;
        const int r0 = q * 32;
        sc1 = tx * g.TW + r0 % g.TW;
        sc2 = (t2 % g.tiles_y) * g.TH + (r0 / g.TW) % g.TH;
        sc3 = (t2 / g.tiles_y) * g.TN + r0 / (g.TW * g.TH);
      } else {
        sc1 = mt * BM + q * 32;
      }
      // bias slice of this tile -> shared memory (zeros where absent / beyond N)
#pragma unroll
      for (int j = 0; j < kEpiBiasFloats / 32; ++j) {
        const int nn = n_tile + j * 32 + lane;
        bsm[j * 32 + lane] = (use_bias && j * 32 < p.bn && nn < p.N) ? __ldg(p.bias + nn) : 0.f;
      }
      // residual chunk: 32 halfs of this thread's row (zeros outside the tensor), prefetched one chunk ahead
      const __half* res_row = (use_res && m_own >= 0) ? p.residual + (size_t)m_own * p.ldr + n_tile : nullptr;
      auto load_res = [&](int c, uint4* rr) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          rr[j] = make_uint4(0u, 0u, 0u, 0u);
          if (res_row != nullptr && c < p.bn && n_tile + c + j * 8 < p.N)
            rr[j] = *reinterpret_cast<const uint4*>(res_row + c + j * 8);
        }
      };
      uint4 rr[4];
      load_res(0, rr);
      __syncwarp();
      tf::mbar_wait(tfull_bar(as), aphase);
      if (t == blockIdx.x && threadIdx.x == 64) TF_STAMP(4);   // accumulator of the first tile complete
      tf::tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * kAccStride;
      for (int c = 0; c < p.bn; c += 32) {
        uint32_t v[32];
        tf::tmem_ld_x16(taddr + c, v);
        tf::tmem_ld_x16(taddr + c + 16, v + 16);   // BN is a multiple of 32 (TMA store granularity)
        uint4 rn[4];
        load_res(c + 32, rn);            // next chunk's residual goes in flight now
        tf::tmem_ld_wait();
        if (c + 32 >= p.bn) {            // accumulator fully read: hand the TMEM buffer back to the MMA warp
          tf::tcgen05_fence_before();
          tf::mbar_arrive(tempty_bar(as));
        }
        // fp16 / GEGLU blocks are <= 2 KB: two staging halves alternate, so only the store issued two chunks
        // ago must have finished reading shared memory; fp32 blocks use the whole 4 KB
        const uint32_t stg_c = stg + (out_f32 ? 0u : (uint32_t)((c >> 5) & 1) * 2048u);
        const uint32_t srow = stg_c + lane * row_bytes;
        if (lane == 0) {
          if (out_f32) tf::tma_store_wait_read<0>();
          else tf::tma_store_wait_read<1>();
        }
        __syncwarp();
        if (geglu) {
          // packed columns: [c, c+16) = value, [c+16, c+32) = gate  (tinyfusers_b200/packing.py: geglu_pack)
          uint32_t h[8];
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            float f[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const float val = __uint_as_float(v[j + i]) + bsm[c + j + i];
              const float gate = __uint_as_float(v[16 + j + i]) + bsm[c + 16 + j + i];
              f[i] = val * tf::gelu_tanh_f(gate);
            }
            __half2 hh = __floats2half2_rn(f[0], f[1]);
            h[j >> 1] = *reinterpret_cast<uint32_t*>(&hh);
          }
#pragma unroll
          for (int j = 0; j < 2; ++j)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + ((j ^ swz) << 4)), "r"(h[4 * j]),
                         "r"(h[4 * j + 1]), "r"(h[4 * j + 2]), "r"(h[4 * j + 3])
                         : "memory");
        } else {
          float f[32];
          const __half2* r2 = reinterpret_cast<const __half2*>(rr);
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float2 rv = __half22float2(r2[j >> 1]);
            f[j] = __uint_as_float(v[j]) + bsm[c + j] + rv.x;
            f[j + 1] = __uint_as_float(v[j + 1]) + bsm[c + j + 1] + rv.y;
          }
          if (out_f32) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + ((j ^ swz) << 4)),
                           "r"(__float_as_uint(f[4 * j])), "r"(__float_as_uint(f[4 * j + 1])),
                           "r"(__float_as_uint(f[4 * j + 2])), "r"(__float_as_uint(f[4 * j + 3]))
                           : "memory");
          } else {
            uint32_t h[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              __half2 hh = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
              h[j] = *reinterpret_cast<uint32_t*>(&hh);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + ((j ^ swz) << 4)), "r"(h[4 * j]),
                           "r"(h[4 * j + 1]), "r"(h[4 * j + 2]), "r"(h[4 * j + 3])
                           : "memory");
          }
        }
        tf::fence_proxy_async_smem();   // generic-proxy writes -> visible to the TMA (async proxy)
        __syncwarp();
        if (lane == 0) {
          const int col = geglu ? ((n_tile + c) >> 1) : (n_tile + c);
          // split-K partials live in a tensor with one extra (split) dimension, so a tile's overhang is clipped
          // per split instead of spilling into the next split's slab
          if (p.is_conv) {
            if (partial) tf::tma_store_5d(&tmC, stg_c, col, sc1, sc2, sc3, split);
            else tf::tma_store_4d(&tmC, stg_c, col, sc1, sc2, sc3);
          } else {
            if (partial) tf::tma_store_3d(&tmC, stg_c, col, sc1, split);
            else tf::tma_store_2d(&tmC, stg_c, col, sc1);
          }
          tf::tma_store_commit();
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) rr[j] = rn[j];
      }
      if (t == blockIdx.x && threadIdx.x == 64) TF_STAMP(5);   // first tile stored (issued)
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
    if (lane == 0) tf::tma_store_wait<0>();   // all bulk stores complete before the CTA retires
  }

  tf::tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tf::tcgen05_fence_after();
    tf::tmem_dealloc(tmem_base, kTmemCols);
  }
  if (threadIdx.x == 0) TF_STAMP(6);
#undef TF_STAMP
}

// ------------------------------------------------------------------------------------------------
// split-K fold: out[m,n] = sum_s partial[s][m][n] (+bias) (+residual)
// ------------------------------------------------------------------------------------------------
__global__ void tf_splitk_reduce_kernel(const float* __restrict__ partial, int splits, int M, int N,
                                        const float* __restrict__ bias,
                                        const __half* __restrict__ residual, int ldr, void* out,
                                        int ldc, int out_f32) {
  tf::pdl_prologue();  // PDL: let the next kernel start launching, then wait for our producers
  const size_t idx = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const size_t total = (size_t)M * N;
  if (idx >= total) return;
  const int m = (int)(idx / N);
  const int n = (int)(idx % N);
  float4 acc = *reinterpret_cast<const float4*>(partial + idx);
  for (int s = 1; s < splits; ++s) {
    float4 t = *reinterpret_cast<const float4*>(partial + (size_t)s * total + idx);
    acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
  }
  if (bias) {
    float4 b = __ldg(reinterpret_cast<const float4*>(bias + n));
    acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
  }
  if (residual) {
    const __half2* r = reinterpret_cast<const __half2*>(residual + (size_t)m * ldr + n);
    float2 r0 = __half22float2(r[0]), r1 = __half22float2(r[1]);
    acc.x += r0.x; acc.y += r0.y; acc.z += r1.x; acc.w += r1.y;
  }
  if (out_f32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + (size_t)m * ldc + n) = acc;
  } else {
    __half2* o = reinterpret_cast<__half2*>(reinterpret_cast<__half*>(out) + (size_t)m * ldc + n);
    o[0] = __floats2half2_rn(acc.x, acc.y);
    o[1] = __floats2half2_rn(acc.z, acc.w);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct TileChoice {
  int bn, splits;
};

// crude cycle model: per 64-deep k-block a CTA needs max(tensor, smem-feed) cycles; pick the
// (BN, split-K) pair with the lowest wave-quantised estimate.
static TileChoice choose_tiles(int m_tiles, int N, int k_blocks, int flags, bool allow_split,
                               size_t ws_bytes, int M, int force_bn, int force_splits) {
  const int sms = tf_num_sms();
  TileChoice best{128, 1};
  double best_cost = 1e30;
  (void)flags;
  const int step = 32;  // epilogue ships 32-column blocks
  for (int bn = step; bn <= 256; bn += step) {
    if (force_bn > 0 && bn != force_bn) continue;
    const int n_tiles = ceil_div_i(N, bn);
    // avoid heavily padded N tiles
    const double n_eff = (double)N / (n_tiles * bn);
    if (n_eff < 0.8 && force_bn <= 0 && bn > step) continue;
    const int max_split = allow_split ? 16 : 1;
    for (int sp = 1; sp <= max_split; ++sp) {
      if (force_splits > 0 && sp != force_splits) continue;
      if (sp > 1) {
        if (k_blocks / sp < 4 && force_splits <= 0) break;
        if ((size_t)sp * M * N * sizeof(float) > ws_bytes) break;
      }
      const int kbs = ceil_div_i(k_blocks, sp);
      if ((sp - 1) * kbs >= k_blocks) continue;  // an empty split
      const long tiles = (long)m_tiles * n_tiles * sp;
      const long waves = (tiles + sms - 1) / sms;
      const double per_kb = (2.0 * bn > 128.0 + bn) ? 2.0 * bn : 128.0 + bn;
      double cost = waves * (kbs * per_kb + 1500.0 + 4.0 * bn);
      if (sp > 1) cost += 8000.0 + (double)sp * M * N * 8.0 / (sms * 64.0);
      if (cost < best_cost) {
        best_cost = cost;
        best = {bn, sp};
      }
    }
  }
  return best;
}

static int g_force_bn = 0, g_force_splits = 0;
static long long* g_timeline = nullptr;

static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, GemmParams& p,
                       cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    TF_CUDA(cudaFuncSetAttribute(tf_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 kSmemBudget));
    attr_set = true;
  }
  const int stage_bytes = A_STAGE_BYTES + p.bn * 128;
  int stages = (kSmemBudget - 2048 - kEpiBytes) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) {
    tf_set_error("gemm: tile too large for shared memory");
    return TF_ERR_ARG;
  }
  p.stages = stages;
  // always carve > half of the SM's shared memory: one CTA per SM, so the 512-column TMEM
  // allocation can never contend with a co-resident CTA.
  size_t smem = (size_t)stages * stage_bytes + kEpiBytes + 2048;
  if (smem < 120 * 1024) smem = 120 * 1024;
  const int total_tiles = p.m_tiles * p.n_tiles * p.splits;
  int grid = total_tiles < tf_num_sms() ? total_tiles : tf_num_sms();
  p.timeline = g_timeline;
  TF_LAUNCH(tf_gemm_kernel, grid, kThreads, smem, stream, tmA, tmB, tmC, p);
  TF_LAUNCH_CHECK();
  tf_launch_count_add(1);
  if (p.splits > 1) {
    const size_t total = (size_t)p.M * p.N;
    const int threads = 256;
    const int blocks = (int)((total / 4 + threads - 1) / threads);
    TF_LAUNCH(tf_splitk_reduce_kernel, blocks, threads, 0, stream, 
        p.partial, p.splits, p.M, p.N, p.bias, p.residual, p.ldr, p.out, p.ldc,
        (p.flags & TF_EPI_OUT_F32) ? 1 : 0);
    TF_LAUNCH_CHECK();
    tf_launch_count_add(1);
  }
  return TF_OK;
}

}  // namespace

extern "C" int tf_gemm_set_timeline(long long* dev_buf) {
  g_timeline = dev_buf;  // >= 148*8 int64; debug only
  return TF_OK;
}

extern "C" int tf_gemm_set_tuning(int force_bn, int force_splits) {
  g_force_bn = force_bn;
  g_force_splits = force_splits;
  return TF_OK;
}

extern "C" int tf_gemm_f16(const void* A, int lda, const void* W, int ldw, void* out, int ldc, int M,
                           int N, int K, const float* bias, const void* residual, int ldr, int flags,
                           void* workspace, size_t ws_bytes, void* stream) {
  TF_CHECK_ARG(A && W && out, "tf_gemm_f16: null pointer");
  TF_CHECK_ARG(M > 0 && N > 0 && K > 0, "tf_gemm_f16: bad dims M=%d N=%d K=%d", M, N, K);
  TF_CHECK_ARG(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "tf_gemm_f16: K, lda, ldw must be multiples of 8");
  TF_CHECK_ARG(N % 8 == 0 && ldc % 8 == 0, "tf_gemm_f16: N and ldc must be multiples of 8 (N=%d ldc=%d)", N, ldc);
  TF_CHECK_ARG(((uintptr_t)A & 15) == 0 && ((uintptr_t)W & 15) == 0 && ((uintptr_t)out & 15) == 0,
               "tf_gemm_f16: pointers must be 16-byte aligned");
  if (residual) TF_CHECK_ARG(ldr % 8 == 0 && ((uintptr_t)residual & 15) == 0, "tf_gemm_f16: residual alignment");
  if (flags & TF_EPI_GEGLU) TF_CHECK_ARG(N % 32 == 0 && !residual, "tf_gemm_f16: GEGLU needs N %% 32 == 0, no residual");

  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.is_conv = 0;
  p.m_tiles = ceil_div_i(M, BM);
  p.k_blocks = ceil_div_i(K, BK);
  const bool allow_split = !(flags & TF_EPI_GEGLU) && workspace != nullptr;
  TileChoice tc = choose_tiles(p.m_tiles, N, p.k_blocks, flags, allow_split, ws_bytes, M, g_force_bn,
                               g_force_splits);
  p.bn = tc.bn;
  p.splits = tc.splits;
  p.n_tiles = ceil_div_i(N, p.bn);
  p.kb_per_split = ceil_div_i(p.k_blocks, p.splits);
  p.out = out; p.ldc = ldc; p.bias = bias;
  p.residual = reinterpret_cast<const __half*>(residual); p.ldr = ldr;
  p.flags = flags;

  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t strides[1] = {(uint64_t)lda * 2};
    uint32_t box[2] = {BK, BM};
    uint32_t es[2] = {1, 1};
    int rc = tf_encode_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, A, dims, strides, box, es,
                            CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t strides[1] = {(uint64_t)ldw * 2};
    uint32_t box[2] = {BK, (uint32_t)p.bn};
    uint32_t es[2] = {1, 1};
    int rc = tf_encode_tmap(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, W, dims, strides, box, es,
                            CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  CUtensorMap tmC;
  {
    const bool geglu = (flags & TF_EPI_GEGLU) != 0;
    const bool f32 = (flags & TF_EPI_OUT_F32) != 0 || p.splits > 1;
    p.partial = p.splits > 1 ? reinterpret_cast<float*>(workspace) : nullptr;
    const void* base = p.splits > 1 ? workspace : out;
    const uint64_t cols = geglu ? (uint64_t)N / 2 : (uint64_t)N;
    const bool part = p.splits > 1;
    const uint64_t ld = part ? (uint64_t)N : (uint64_t)ldc;
    uint64_t dims[3] = {cols, (uint64_t)M, (uint64_t)p.splits};
    uint64_t strides[2] = {ld * (f32 ? 4 : 2), (uint64_t)M * ld * (f32 ? 4 : 2)};
    uint32_t box[3] = {geglu ? 16u : 32u, 32u, 1u};
    uint32_t es[3] = {1, 1, 1};
    int rc = tf_encode_tmap(&tmC, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, part ? 3 : 2, base,
                            dims, strides, box, es,
                            f32 ? CU_TENSOR_MAP_SWIZZLE_128B : (geglu ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B));
    if (rc) return rc;
  }
  return launch_gemm(tmA, tmB, tmC, p, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int tf_conv2d_nhwc_f16(const void* x, int NI, int H, int W, int Cin, int x_pixel_stride,
                                  const void* w, int Cout, int ksize, int stride, void* out, int ldc,
                                  const float* bias, const void* residual, int ldr, int flags,
                                  void* workspace, size_t ws_bytes, void* stream) {
  TF_CHECK_ARG(x && w && out, "tf_conv2d_nhwc_f16: null pointer");
  TF_CHECK_ARG(ksize == 1 || ksize == 3, "tf_conv2d_nhwc_f16: kernel size %d unsupported (1 or 3)", ksize);
  TF_CHECK_ARG(stride == 1 || stride == 2, "tf_conv2d_nhwc_f16: stride %d unsupported (1 or 2)", stride);
  TF_CHECK_ARG(NI > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "tf_conv2d_nhwc_f16: bad dims");
  TF_CHECK_ARG(x_pixel_stride >= Cin && x_pixel_stride % 8 == 0, "tf_conv2d_nhwc_f16: bad pixel stride");
  if (ksize == 1 && stride == 1) {
    return tf_gemm_f16(x, x_pixel_stride, w, Cin, out, ldc, NI * H * W, Cout, Cin, bias, residual, ldr,
                       flags, workspace, ws_bytes, stream);
  }
  TF_CHECK_ARG(ksize == 3, "tf_conv2d_nhwc_f16: strided 1x1 unsupported");
  TF_CHECK_ARG(Cin % BK == 0, "tf_conv2d_nhwc_f16: Cin must be a multiple of 64 (got %d)", Cin);
  TF_CHECK_ARG(Cout % 8 == 0 && ldc % 8 == 0, "tf_conv2d_nhwc_f16: Cout and ldc must be multiples of 8");
  TF_CHECK_ARG(!(flags & TF_EPI_GEGLU), "tf_conv2d_nhwc_f16: GEGLU epilogue not valid for conv");
  TF_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0 && ((uintptr_t)out & 15) == 0,
               "tf_conv2d_nhwc_f16: pointers must be 16-byte aligned");
  const int pad = 1;
  const int Ho = (H + 2 * pad - 3) / stride + 1;
  const int Wo = (W + 2 * pad - 3) / stride + 1;

  GemmParams p{};
  p.is_conv = 1;
  p.M = NI * Ho * Wo; p.N = Cout; p.K = 9 * Cin;
  ConvGeom& g = p.g;
  g.H = Ho; g.W = Wo; g.NI = NI; g.cblocks = Cin / BK; g.cscale = stride; g.pad = pad; g.ksize = 3;
  // choose the 128-row tile footprint (TW x TH x TN) with the least padding
  long best_tiles = -1;
  for (int tw = 128; tw >= 1; tw >>= 1) {
    for (int th = 128 / tw; th >= 1; th >>= 1) {
      int tn = 128 / (tw * th);
      if (tw * stride > 256 || th * stride > 256) continue;
      long tiles = (long)ceil_div_i(Wo, tw) * ceil_div_i(Ho, th) * ceil_div_i(NI, tn);
      if (best_tiles < 0 || tiles < best_tiles) {
        best_tiles = tiles;
        g.TW = tw; g.TH = th; g.TN = tn;
      }
    }
  }
  g.tiles_x = ceil_div_i(Wo, g.TW);
  g.tiles_y = ceil_div_i(Ho, g.TH);
  p.m_tiles = g.tiles_x * g.tiles_y * ceil_div_i(NI, g.TN);
  p.k_blocks = 9 * g.cblocks;
  // store box of one epilogue warp = 32 consecutive tile rows (x fastest, then y, then image)
  g.sbw = g.TW >= 32 ? 32 : g.TW;
  g.sbh = g.TW >= 32 ? 1 : (g.TW * g.TH >= 32 ? 32 / g.TW : g.TH);
  TileChoice tc = choose_tiles(p.m_tiles, Cout, p.k_blocks, flags, workspace != nullptr, ws_bytes, p.M, g_force_bn,
                               g_force_splits);
  p.bn = tc.bn;
  p.splits = tc.splits;
  p.n_tiles = ceil_div_i(Cout, p.bn);
  p.kb_per_split = ceil_div_i(p.k_blocks, p.splits);
  p.out = out; p.ldc = ldc; p.bias = bias;
  p.residual = reinterpret_cast<const __half*>(residual); p.ldr = ldr;
  p.flags = flags;

  CUtensorMap tmA, tmB;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)NI};
    uint64_t strides[3] = {(uint64_t)x_pixel_stride * 2, (uint64_t)W * x_pixel_stride * 2,
                           (uint64_t)H * W * x_pixel_stride * 2};
    uint32_t box[4] = {BK, (uint32_t)(g.TW * stride), (uint32_t)(g.TH * stride), (uint32_t)g.TN};
    uint32_t es[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
    int rc = tf_encode_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, x, dims, strides, box, es,
                            CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)p.K, (uint64_t)Cout};
    uint64_t strides[1] = {(uint64_t)p.K * 2};
    uint32_t box[2] = {BK, (uint32_t)p.bn};
    uint32_t es[2] = {1, 1};
    int rc = tf_encode_tmap(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, w, dims, strides, box, es,
                            CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  CUtensorMap tmC;
  {
    const bool part = p.splits > 1;
    const bool f32 = (flags & TF_EPI_OUT_F32) != 0 || part;
    p.partial = part ? reinterpret_cast<float*>(workspace) : nullptr;
    const void* base = part ? workspace : out;
    const uint64_t ld = part ? (uint64_t)Cout : (uint64_t)ldc;
    const uint64_t es_b = f32 ? 4 : 2;
    uint64_t dims[5] = {(uint64_t)Cout, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)NI, (uint64_t)p.splits};
    uint64_t strides[4] = {ld * es_b, (uint64_t)Wo * ld * es_b, (uint64_t)Ho * Wo * ld * es_b,
                           (uint64_t)NI * Ho * Wo * ld * es_b};
    uint32_t box[5] = {32u, (uint32_t)g.sbw, (uint32_t)g.sbh, (uint32_t)(32 / (g.sbw * g.sbh)), 1u};
    uint32_t es[5] = {1, 1, 1, 1, 1};
    int rc = tf_encode_tmap(&tmC, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, part ? 5 : 4, base,
                            dims, strides, box, es, f32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
  }
  return launch_gemm(tmA, tmB, tmC, p, reinterpret_cast<cudaStream_t>(stream));
}
