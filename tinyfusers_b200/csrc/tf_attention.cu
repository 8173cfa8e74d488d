// tinyfusers_b200 — fused flash-style attention on tcgen05 / TMEM for sm_100a.
//
// Replaces the reference's scaled_dot_product_attention (tinyfusers/attention/sdpa.py:53-77):
//   cuBLAS batched SGEMM (scores materialised in HBM, 1.07 GB at 4096 tokens) -> softmax_kernel
//   (native/cuda/softmax.cu:24-112, three global passes per row) -> cuBLAS batched SGEMM.
// Here scores never leave the SM: S = Q·K^T goes to TMEM, four softmax warps (one thread per query
// row — the 32x32b TMEM load hands each thread its own row, so row max / sum need no shuffles) run the
// online softmax and write P (fp16) to shared memory in the 128B-swizzled K-major layout, and
// O += P·V accumulates in TMEM. S lives in ONE TMEM buffer: the softmax warps copy it to registers, so QK^T of
// block j+1 is issued as soon as block j has been drained and overlaps its exponentials. Key blocks are 128 wide
// where shared memory allows two CTAs per SM (head dim <= 48), 64 otherwise: the per-block fixed cost (barrier round
// trips, TMEM load latency, fences; ~700 of ~1800 cycles per 64-key block, measured with tools/dev_attn_timeline.py)
// is paid half as often.
//
// Operand layouts (produced by the projection GEMMs, see tinyfusers_b200/attention/attention.py):
//   Q  : (B*Tq,      ldq)  fp16, head h at columns [h*dp, (h+1)*dp), dp = head dim padded to 16
//                          (pad columns are exact zeros: the projection weight rows are zero-padded)
//   K  : (B*Tk_pad,  ldk)  fp16, same head layout
//   V  : (B*Tk_pad,  ldv)  fp16, NATURAL layout (tf_attention_v_f16, what the models use): head h at columns
//                          [h*dvp, (h+1)*dvp), dvp = head dim zero-padded to 64 (128-byte swizzle atoms) or 16 (32-byte
//                          atoms); tiles are the MN-major B operand of O += P V, so the fused [Q | K | V] projection GEMM
//                          feeds the kernel directly. With dvp - 16 >= round16(d) the L accumulator sits in O's pad columns.
//   Vt : (NH*dp, B*Tk_pad) fp16, V transposed (tf_attention_f16 / tf_attention_causal_f16: the first layout of the round,
//                          kept as an entry point) - K-major B operand
//   O  : fp16, element (b,h,t,j) at b*osb + h*osh + t*ost + j, j < d. The reference's CrossAttention
//        reshapes (B,NH,T,HS) straight to (B,T,NH*HS) (attention.py:39) — that is osb=NH*T*d, osh=T*d,
//        ost=d; the canonical head merge is osb=T*NH*d, osh=d, ost=NH*d.
// Q/K tiles use 32-byte swizzle slabs of 16 head-dim elements (any dp % 16 == 0 without padding to 64);
// P and V^T use 128-byte swizzle atoms of 64 keys; natural-layout V uses (keys x 64 | 16 columns) boxes.
#include "tf_common.cuh"
#include "tinyfusers_b200.h"

namespace {

// Debug aid: a host-mapped (pinned) buffer that survives a device trap. When set (tf_attention_set_debug), a barrier wait of
// the attention kernels that times out records {code, block x/y/z, warp, key block} there before it traps, so the host can
// tell WHICH wait hung after the context is gone. Null in production: one predictable branch on the (never taken) timeout path.
__device__ unsigned long long* g_att_dbg = nullptr;

__device__ __forceinline__ void att_wait(uint32_t bar, uint32_t parity, int code, int j) {
  if (tf::mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!tf::mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > TF_WAIT_TIMEOUT_CYCLES) {
      unsigned long long* d = g_att_dbg;
      if (d != nullptr && atomicCAS(d, 0ull, 1ull) == 0ull) {
        d[1] = (unsigned long long)code; d[2] = blockIdx.x; d[3] = blockIdx.y; d[4] = blockIdx.z;
        d[5] = threadIdx.x >> 5; d[6] = (unsigned long long)(long long)j; d[7] = parity;
        __threadfence_system();
      }
      __trap();
    }
  }
}

constexpr int kAttThreads = 192;
constexpr int BQ = 128;

struct AttnParams {
  int B, NH, Tq, Tk, Tk_pad, d, dp;
  int dov;      // O columns in TMEM: dp (V^T operand), or the V head dim padded to 64 (natural-layout V operand)
  int v_atom;   // natural-layout V: columns per swizzle atom - 64 (128-byte swizzle) or 16 (32-byte swizzle, dov % 64 != 0)
  int l_off;    // first of the 16 L (= P . 1) columns, relative to O: dov, or dov - 16 when O's zero pad columns can host them
  int ol_cols;  // columns spanned by O and L together
  int nkv;      // key blocks
  int stages;   // K/V ring depth
  int causal;   // 1: query t only sees keys <= t (CLIP's additive triu(-inf, k=1) mask, vae/encoder.py:79)
  uint32_t tmem_cols;
  float scale_log2;  // (1/sqrt(d)) * log2(e)
  __half* out;
  long long osb, osh, ost;
  long long* timeline;   // debug (TF_ATT_TRACE build): clock stamps of CTA (0,0,0), softmax warp 0: [block][8]
};

#ifndef TF_ATT_TRACE
#define TF_ATT_TRACE 0
#endif
#if TF_ATT_TRACE
#define ATT_STAMP(i) do { if (p.timeline && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 64 && j < 64) p.timeline[j * 8 + (i)] = clock64(); } while (0)
#else
#define ATT_STAMP(i) do { } while (0)
#endif

// K-major operand, 32-byte swizzle: rows are 32 B (16 fp16), 8-row groups 256 B apart.
__device__ __forceinline__ uint64_t umma_desc_sw32_kmajor(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(256u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;  // SWIZZLE_32B
  return d;
}

// OCC = CTAs per SM the register allocation is bounded for (3: 64-key blocks, head dim <= 48, long sequences: three
// resident CTAs keep the MUFU pipe ~85 % busy instead of ~57 %, and 444 slots swallow the 512-CTA grid of a
// 4096-token SD self-attention in one wave plus a short tail)
// MN-major operand (rows = K index, 128 contiguous bytes = 64 elements of the M/N index), 128-byte swizzle:
// canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units (cute/atom/mma_traits_sm100.hpp) - 8-row groups are
// SBO = 1024 B apart, 64-element column blocks LBO apart. This is V in its natural (keys, head dim) layout as the
// B operand of O += P V: no transposed copy of V has to be produced by anyone.
__device__ __forceinline__ uint64_t umma_desc_sw128_mnmajor(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}

// Same operand with 32-byte swizzle atoms (16 columns): ((2,n),(8,k)):((1,LBO),(2,SBO)) - 8-row groups 256 B apart. Used
// when the head dim is not worth padding to 64 columns (d = 40 -> 48, d = 80 -> 80).
__device__ __forceinline__ uint64_t umma_desc_sw32_mnmajor(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(256u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;  // SWIZZLE_32B
  return d;
}

// VNAT: V is read in its natural layout (B*Tk_pad rows, head h at columns [h*dov, (h+1)*dov), dov = head dim padded to
// 64 with zero columns) instead of transposed.
template <int BN, int OCC, bool VNAT>
__global__ void __launch_bounds__(kAttThreads, OCC)
tf_attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = tf::smem_u32(smem_raw);
  const uint32_t smem_base = (raw_u32 + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int ST = p.stages;

  const uint32_t q_bytes = BQ * p.dp * 2;
  const uint32_t k_bytes = BN * p.dp * 2;
  const uint32_t v_bytes = BN * p.dov * 2;  // V^T tile: dp rows x BN keys; natural V tile: BN keys x dov columns
  const uint32_t p_bytes = BQ * BN * 2;
  const uint32_t smem_q = smem_base;
  const uint32_t smem_k = smem_q + q_bytes;
  const uint32_t smem_v = smem_k + ST * k_bytes;
  const uint32_t smem_p = smem_v + ST * v_bytes;
  const uint32_t smem_ones = smem_p + p_bytes;   // 16 x 64 fp16 ones (one 128B-swizzle atom): the row sums come from the MMA
  const uint32_t bar_base = smem_ones + 2048;
  // barriers
  const uint32_t q_full = bar_base;
  auto kv_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar_base + 8u * (1 + ST + s); };
  const uint32_t s_full = bar_base + 8u * (1 + 2 * ST);
  const uint32_t s_empty = bar_base + 8u * (3 + 2 * ST);
  const uint32_t p_full = bar_base + 8u * (5 + 2 * ST);
  const uint32_t pv_done = bar_base + 8u * (6 + 2 * ST);
  const uint32_t tmem_slot = bar_base + 8u * (7 + 2 * ST);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw_u32));

  if (warp == 0 && lane == 0) {
    tf::tma_prefetch_desc(&tmQ);
    tf::tma_prefetch_desc(&tmK);
    tf::tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    tf::mbar_init(q_full, 1);
    for (int s = 0; s < ST; ++s) {
      tf::mbar_init(kv_full(s), 1);
      tf::mbar_init(kv_empty(s), 1);
    }
    tf::mbar_init(s_full, 1);
    tf::mbar_init(s_empty, 128);
    tf::mbar_init(p_full, 128);
    tf::mbar_init(pv_done, 1);
    tf::fence_mbar_init();
  }
  if (warp == 2) {
    tf::tmem_alloc(tmem_slot, p.tmem_cols);
    tf::tmem_relinquish();
  }
  if (warp >= 2) {   // 2 KB of fp16 1.0 (swizzling a constant tile is the identity)
    const uint32_t i = threadIdx.x - 64;
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(smem_ones + i * 16u), "r"(0x3C003C00u) : "memory");
    tf::fence_proxy_async_smem();
  }
  tf::tcgen05_fence_before();
  __syncthreads();
  tf::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  tf::pdl_trigger();   // only once this CTA holds its tensor memory (see tf_gemm_kernel: trigger-before-alloc can deadlock)
  tf::pdl_wait();
  const uint32_t tmem_s0 = tmem_base;            // S: columns [0,BN)
  const uint32_t tmem_o = tmem_base + BN;        // O: dp columns, then 16 columns of L = P . 1 (the softmax denominator)
  const uint32_t tmem_l = tmem_o + p.l_off;

  const int nkv = p.nkv;
  const int slabs = p.dp / 16;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      tf::mbar_expect_tx(q_full, q_bytes);
      tf::tma_load_3d(smem_q, &tmQ, q_full, 0, b * p.Tq + qt * BQ, h * slabs);
      for (int j = 0; j < nkv; ++j) {
        const int s = j % ST;
        const uint32_t u = j / ST;
        tf::mbar_wait(kv_empty(s), (u & 1u) ^ 1u);
        tf::mbar_expect_tx(kv_full(s), k_bytes + v_bytes);
        const int key0 = b * p.Tk_pad + j * BN;
        tf::tma_load_3d(smem_k + s * k_bytes, &tmK, kv_full(s), 0, key0, h * slabs);
        if (VNAT) {   // one (BN keys x v_atom columns) box per column block of the head
          const int nblk = p.dov / p.v_atom;
          for (int nb = 0; nb < nblk; ++nb)
            tf::tma_load_2d(smem_v + s * v_bytes + nb * (BN * p.v_atom * 2), &tmV, kv_full(s), h * p.dov + nb * p.v_atom, key0);
        } else {
#pragma unroll
          for (int i = 0; i < BN / 64; ++i)
            tf::tma_load_2d(smem_v + s * v_bytes + i * (p.dp * 128), &tmV, kv_full(s), key0 + 64 * i, h * p.dp);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc_s = tf::umma_idesc_f16(BQ, BN);
      const uint32_t idesc_o = tf::umma_idesc_f16(BQ, p.dov) | (VNAT ? (1u << 16) : 0u);   // bit 16: B is MN-major
      const uint32_t idesc_l = tf::umma_idesc_f16(BQ, 16);
      auto issue_s = [&](int j) {
        const int s = j % ST;
        tf::mbar_wait(kv_full(s), (uint32_t)(j / ST) & 1u);
        // S_{j-1} must have been drained to registers by all four softmax warps
        tf::mbar_wait(s_empty, ((uint32_t)j & 1u) ^ 1u);
        tf::tcgen05_fence_after();
        const uint32_t kbase = smem_k + s * k_bytes;
        for (int k = 0; k < slabs; ++k) {
          tf::umma_f16_ss(tmem_s0, umma_desc_sw32_kmajor(smem_q + k * (BQ * 32)),
                          umma_desc_sw32_kmajor(kbase + k * (BN * 32)), idesc_s, k > 0 ? 1u : 0u);
        }
        tf::umma_commit(s_full);
      };
      tf::mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < nkv; ++j) {
        if (j + 1 < nkv) issue_s(j + 1);
        tf::mbar_wait(p_full, (uint32_t)j & 1u);
        tf::tcgen05_fence_after();
        const int s = j % ST;
        const uint32_t vbase = smem_v + s * v_bytes;
#pragma unroll
        for (int k = 0; k < BN / 16; ++k) {
          const uint32_t atom = k >> 2, sub = k & 3;
          const uint64_t vdesc = !VNAT ? tf::umma_desc_sw128_kmajor(vbase + atom * (p.dp * 128)) + 2u * sub
                                 : (p.v_atom == 64 ? umma_desc_sw128_mnmajor(vbase + k * 2048u, BN * 128u)   // keys 16k .. 16k+15
                                                   : umma_desc_sw32_mnmajor(vbase + k * 512u, BN * 32u));
          tf::umma_f16_ss(tmem_o, tf::umma_desc_sw128_kmajor(smem_p + atom * (BQ * 128)) + 2u * sub, vdesc, idesc_o,
                          (j > 0 || k > 0) ? 1u : 0u);
          // L += P . ones: the denominator accumulates from the SAME fp16-rounded P the numerator uses
          tf::umma_f16_ss(tmem_l, tf::umma_desc_sw128_kmajor(smem_p + atom * (BQ * 128)) + 2u * sub,
                          tf::umma_desc_sw128_kmajor(smem_ones) + 2u * sub, idesc_l, (j > 0 || k > 0) ? 1u : 0u);
        }
        tf::umma_commit(pv_done);
        tf::umma_commit(kv_empty(s));
      }
    }
  } else {
    // ===================== softmax / correction / epilogue (warps 2..5) =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_field = (uint32_t)(q * 32) << 16;
    // Online softmax, one thread per query row. Per element the loop is FMNMX (max) + FFMA (scale and shift folded)
    // + MUFU.EX2 + half a pack: the row sum is NOT accumulated here (the L columns of the PV MMA do it), and the
    // running max is lazy: O / L are only rescaled when a row's max grows by more than 2^8 (P stays <= 256 in
    // fp16, accumulation is fp32, and O / L is exact whatever reference max is used).
    float m_run = -INFINITY;  // reference max of this row, in the scaled log2 domain
    for (int j = 0; j < nkv; ++j) {
      ATT_STAMP(0);
      tf::mbar_wait(s_full, (uint32_t)j & 1u);
      tf::tcgen05_fence_after();
      ATT_STAMP(1);
      constexpr int NH64 = BN / 64;              // 64-column halves of the block: registers hold one half at a time
      int valid = min(BN, p.Tk - j * BN);
      if (p.causal) valid = min(valid, qt * BQ + row + 1 - j * BN);
      // keys >= valid (padding / another batch / above the causal diagonal, per row) are masked
      uint32_t v[64];
      auto load_half = [&](int hh) {
#pragma unroll
        for (int c = 0; c < 64; c += 32) tf::tmem_ld_x32(tmem_s0 + lane_field + hh * 64 + c, v + c);
        tf::tmem_ld_wait();
        const int vh = valid - hh * 64;
        if (vh < 64) {
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (i >= vh) v[i] = 0xff800000u;   // -inf
        }
      };
      // ---- pass 1: block max (BN = 128 re-reads the halves from TMEM in pass 2: a 64-column load is ~30 cycles) ----
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int hh = 0; hh < NH64; ++hh) {
        load_half(hh);
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
#pragma unroll
          for (int u = 0; u < 4; ++u) mx4[u] = fmaxf(mx4[u], __uint_as_float(v[i + u]));
        }
      }
      ATT_STAMP(2);
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * p.scale_log2;   // scale > 0
      float alpha = 1.0f;
      if (mx > m_run + 8.0f) {          // also taken on the first block (m_run = -inf)
        alpha = exp2f(m_run - mx);      // 0 on the first block
        m_run = mx;
      }
      ATT_STAMP(3);
      const float neg_m = -m_run;
      // ---- pass 2: exponentials, pack, store P ----
#pragma unroll
      for (int hh = 0; hh < NH64; ++hh) {
        if (NH64 > 1) load_half(hh);   // BN = 64 still holds its only half
        if (hh == NH64 - 1) {          // S is in registers: the tensor core may overwrite it with S_{j+1}
          tf::tcgen05_fence_before();
          tf::mbar_arrive(s_empty);
        }
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < 64; i += 2) {
          const float p0 = exp2f(fmaf(__uint_as_float(v[i]), p.scale_log2, neg_m));
          const float p1 = exp2f(fmaf(__uint_as_float(v[i + 1]), p.scale_log2, neg_m));
          __half2 h2v = __floats2half2_rn(p0, p1);
          pk[i >> 1] = *reinterpret_cast<uint32_t*>(&h2v);
        }
        if (hh == 0) {
          ATT_STAMP(4);
          // P smem and O are owned by the tensor core until PV_{j-1} has completed
          if (j > 0) {
            tf::mbar_wait(pv_done, (uint32_t)(j - 1) & 1u);
            tf::tcgen05_fence_after();
          }
          ATT_STAMP(5);
        }
        // P -> shared memory, K-major 128B swizzle: 16-byte chunk index XOR (row & 7)
        const uint32_t rbase = smem_p + hh * (BQ * 128) + row * 128;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          const uint32_t addr = rbase + ((uint32_t)(ch ^ (row & 7)) << 4);
          const int w = ch * 4;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[w]), "r"(pk[w + 1]),
                       "r"(pk[w + 2]), "r"(pk[w + 3])
                       : "memory");
        }
      }
      // rescale the running output if any row of this warp moved its max
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
        for (int c = 0; c < p.ol_cols; c += 16) {   // O and the L columns
          uint32_t o[16];
          tf::tmem_ld_x16(tmem_o + lane_field + c, o);
          tf::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tf::tmem_st_x16(tmem_o + lane_field + c, o);
        }
        tf::tmem_st_wait();
      }
      ATT_STAMP(6);
      tf::fence_proxy_async_smem();
      tf::tcgen05_fence_before();
      tf::mbar_arrive(p_full);
      ATT_STAMP(7);
    }
    // ---- epilogue: O / l -> fp16 -> global ----
    tf::mbar_wait(pv_done, (uint32_t)(nkv - 1) & 1u);
    tf::tcgen05_fence_after();
    float inv_l;
    {
      uint32_t l16[16];
      tf::tmem_ld_x16(tmem_l + lane_field, l16);
      tf::tmem_ld_wait();
      inv_l = 1.0f / __uint_as_float(l16[0]);
    }
    const int t = qt * BQ + row;
    __half* orow = p.out + (long long)b * p.osb + (long long)h * p.osh + (long long)t * p.ost;
    for (int c = 0; c < p.d; c += 16) {
      uint32_t o[16];
      tf::tmem_ld_x16(tmem_o + lane_field + c, o);
      tf::tmem_ld_wait();
      if (t < p.Tq) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (c + g * 8 < p.d) {
            tf::Pack16 pk8;
#pragma unroll
            for (int i = 0; i < 4; ++i)
              pk8.h2[i] = __floats2half2_rn(__uint_as_float(o[g * 8 + 2 * i]) * inv_l,
                                            __uint_as_float(o[g * 8 + 2 * i + 1]) * inv_l);
            *reinterpret_cast<uint4*>(orow + c + g * 8) = pk8.v;
          }
        }
      }
    }
  }

  tf::tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tf::tcgen05_fence_after();
    tf::tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// =====================================================================================================================
// tf_attention2_kernel: two query tiles per CTA in ping-pong (round 2).
//
// What limited tf_attention_kernel on the 4096-token self-attention (profiles/attention_timeline_r1.md): a softmax warp's
// key block is a ~1800-cycle serial chain of which only ~700 are exponentials - the rest is waiting for S, the shared-memory
// round trip of P (st.shared + fence.proxy.async), and barrier latency - so two co-resident CTAs keep the XU pipe ~55 % busy.
// Here one CTA (one per SM, 384 threads) owns TWO 128-row query tiles of the same (batch, head) that share every K/V tile:
//   warp 0      TMA producer (Q0, Q1, K/V ring)
//   warp 1      tcgen05.mma issuer
//   warps 4-7   softmax of tile 0,   warps 8-11  softmax of tile 1   (one thread per query row)
// * S is double-buffered per tile in TMEM (64-key blocks): S_{j+1} = Q K_{j+1}^T is issued before the P V_j of either tile,
//   so a softmax warpgroup never waits for the tensor core in steady state - its chain is load, max, exp, store.
// * P never visits shared memory: the softmax writes fp16 P over the S columns it has just consumed (tcgen05.st) and
//   O += P V is a TS MMA (A operand from tensor memory). No st.shared, no fence.proxy.async.
// * EMU of every 8 exponentials run on the FMA pipe (Cody-Waite split + degree-3 minimax polynomial, packed FFMA2 /
//   FADD2), the rest on MUFU.EX2: at head dim 40 the kernel is bound by 16 MUFU lanes per SM, not by the tensor core
//   (per 128x128 block: ~512 tensor cycles vs 1024 MUFU cycles).
// TMEM per tile: [S buf 0 (64) | S buf 1 (64) | O, L (ol_cols)]; P_j aliases the first 32 columns of S buffer j % 2.
// =====================================================================================================================
constexpr int kA2Threads = 384;
constexpr int BN2 = 64;

struct Attn2Params {
  int B, NH, Tq, Tk, Tk_pad, d, dp, dov, v_atom, l_off, ol_cols, nkv, stages, causal;
  int tile_cols;   // TMEM columns per query tile: 128 + ol_cols
  int l_sel;       // position of the row sum inside the 16-column group at l_off (0 or 8)
  int l_mma;       // 1: row sums by an extra N = 16 MMA against a ones tile (at l_off); 0: V carries a ones column at index d
  long long* timeline;   // debug (TF_ATT_TRACE build): clock stamps of CTA (0,0,0): [who: softmax 0, softmax 1, MMA][block < 64][8]
  uint32_t tmem_cols;
  float scale_log2;
  __half* out;
  long long osb, osh, ost;
};

// 2^x for x <= ~8 on the FMA pipe, two values per instruction: x = n + f, n = round(x) (magic-number add), f in [-0.5, 0.5],
// 2^f by a degree-3 minimax polynomial (max relative error 7.6e-5, below half an fp16 ulp of the P it produces), 2^n by
// adding n into the exponent field.
__device__ __forceinline__ void exp2_fma_pair(float& x0, float& x1) {
  const float kMagic = 12582912.f;   // 1.5 * 2^23
  x0 = fmaxf(x0, -126.f);
  x1 = fmaxf(x1, -126.f);
  const uint64_t x = tf::pack_f32x2(x0, x1);
  const uint64_t t = tf::add_f32x2(x, tf::pack_f32x2(kMagic, kMagic));
  const uint64_t n = tf::add_f32x2(t, tf::pack_f32x2(-kMagic, -kMagic));
  const uint64_t f = tf::fma_f32x2(n, tf::pack_f32x2(-1.f, -1.f), x);
  uint64_t pl = tf::fma_f32x2(f, tf::pack_f32x2(0.05520550534129143f, 0.05520550534129143f),
                              tf::pack_f32x2(0.24261397123336792f, 0.24261397123336792f));
  pl = tf::fma_f32x2(pl, f, tf::pack_f32x2(0.6932547688484192f, 0.6932547688484192f));
  pl = tf::fma_f32x2(pl, f, tf::pack_f32x2(0.9999276995658875f, 0.9999276995658875f));
  float p0, p1, t0, t1;
  tf::unpack_f32x2(pl, p0, p1);
  tf::unpack_f32x2(t, t0, t1);
  x0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  x1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

#if TF_ATT_TRACE
#define ATT2_STAMP(who, i) do { if (p.timeline && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && j < 64) p.timeline[((who) * 64 + j) * 8 + (i)] = clock64(); } while (0)
#else
#define ATT2_STAMP(who, i) do { } while (0)
#endif
template <int EMU>
__global__ void __launch_bounds__(kA2Threads, 1)
tf_attention2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const Attn2Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = tf::smem_u32(smem_raw);
  const uint32_t smem_base = (raw_u32 + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int ST = p.stages;

  const uint32_t q_bytes = BQ * p.dp * 2;
  const uint32_t k_bytes = BN2 * p.dp * 2;
  const uint32_t v_bytes = BN2 * p.dov * 2;
  const uint32_t smem_q = smem_base;                       // two query tiles
  const uint32_t smem_k = smem_q + 2 * q_bytes;
  const uint32_t smem_v = smem_k + ST * k_bytes;
  const uint32_t smem_ones = smem_v + ST * v_bytes;        // 16 x 64 fp16 ones: B operand of L = P . 1
  const uint32_t bar_base = smem_ones + 2048;
  auto q_full = [&](int t) { return bar_base + 8u * t; };
  auto kv_full = [&](int s) { return bar_base + 8u * (2 + s); };
  auto kv_empty = [&](int s) { return bar_base + 8u * (2 + ST + s); };
  auto s_full = [&](int t, int buf) { return bar_base + 8u * (2 + 2 * ST + 2 * t + buf); };
  // P_j complete (128 arrivals), one barrier per S buffer. A softmax warp may run a whole block ahead of a slower warp of its
  // tile (S_{j+1} is issued before P_j is awaited), so with ONE barrier its arrival for block j+1 would be counted into phase j
  // and the tensor core could read rows of P_j that are not written yet. Block j+2 cannot be reached before phase j is complete
  // (S_{j+2} is issued after it), so two barriers alternating with the buffer keep the phases apart.
  auto p_full = [&](int t, int buf) { return bar_base + 8u * (6 + 2 * ST + 2 * t + buf); };
  auto pv_done = [&](int t) { return bar_base + 8u * (10 + 2 * ST + t); };
  auto pv_last = [&](int t) { return bar_base + 8u * (12 + 2 * ST + t); };   // single use: P V of the LAST key block complete
  const uint32_t tmem_slot = bar_base + 8u * (14 + 2 * ST);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw_u32));

  if (warp == 0 && lane == 0) {
    tf::tma_prefetch_desc(&tmQ);
    tf::tma_prefetch_desc(&tmK);
    tf::tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    for (int t = 0; t < 2; ++t) {
      tf::mbar_init(q_full(t), 1);
      tf::mbar_init(s_full(t, 0), 1);
      tf::mbar_init(s_full(t, 1), 1);
      tf::mbar_init(p_full(t, 0), 128);
      tf::mbar_init(p_full(t, 1), 128);
      tf::mbar_init(pv_done(t), 1);
      tf::mbar_init(pv_last(t), 1);
    }
    for (int s = 0; s < ST; ++s) {
      tf::mbar_init(kv_full(s), 1);
      tf::mbar_init(kv_empty(s), 1);
    }
    tf::fence_mbar_init();
  }
  if (warp == 2) {
    tf::tmem_alloc(tmem_slot, p.tmem_cols);
    tf::tmem_relinquish();
  }
  if (warp == 3) {   // 2 KB of fp16 1.0
    for (int i = lane; i < 128; i += 32)
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(smem_ones + i * 16u), "r"(0x3C003C00u) : "memory");
    tf::fence_proxy_async_smem();
  }
  tf::tcgen05_fence_before();
  __syncthreads();
  tf::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  tf::pdl_trigger();   // only once this CTA holds its tensor memory (see tf_gemm_kernel: trigger-before-alloc can deadlock)
  tf::pdl_wait();

  const int nkv = p.nkv;
  const int slabs = p.dp / 16;
  auto tmem_s = [&](int t, int buf) { return tmem_base + (uint32_t)(t * p.tile_cols + buf * BN2); };
  auto tmem_o = [&](int t) { return tmem_base + (uint32_t)(t * p.tile_cols + 2 * BN2); };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int t = 0; t < 2; ++t) {
        tf::mbar_expect_tx(q_full(t), q_bytes);
        tf::tma_load_3d(smem_q + t * q_bytes, &tmQ, q_full(t), 0, b * p.Tq + (2 * qt + t) * BQ, h * slabs);
      }
      const int nblk = p.dov / p.v_atom;
      for (int j = 0; j < nkv; ++j) {
        const int s = j % ST;
        const uint32_t u = j / ST;
        att_wait(kv_empty(s), (u & 1u) ^ 1u, 1, j);
        tf::mbar_expect_tx(kv_full(s), k_bytes + v_bytes);
        const int key0 = b * p.Tk_pad + j * BN2;
        tf::tma_load_3d(smem_k + s * k_bytes, &tmK, kv_full(s), 0, key0, h * slabs);
        for (int nb = 0; nb < nblk; ++nb)   // one (64 keys x v_atom columns) box per column block of the head
          tf::tma_load_2d(smem_v + s * v_bytes + nb * (BN2 * p.v_atom * 2), &tmV, kv_full(s), h * p.dov + nb * p.v_atom, key0);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp walks the loop converged and one elected lane issues: descriptors, TMEM addresses and barrier
    // addresses are then warp-uniform (uniform registers feed UTCHMMA directly instead of per-instruction R2UR round
    // trips), they advance by adds instead of being rebuilt, and stage / phase are counters instead of divisions. A single
    // lane rebuilding two 64-bit descriptors per instruction needed ~100 cycles per MMA and 22 MMAs per key block: the issue
    // loop, not the softmax, set the block period of the first version (2650 cycles; tools/dev_attn2_timeline.py).
    const bool elected = tf::elect_one();
    const uint32_t idesc_s = tf::umma_idesc_f16(BQ, BN2);
    const uint32_t idesc_o = tf::umma_idesc_f16(BQ, p.dov) | (1u << 16);   // bit 16: B (V, natural layout) is MN-major
    const uint32_t idesc_l = tf::umma_idesc_f16(BQ, 16);
    // descriptor of slab k of query tile t: qd0 + t * q_step + k * 256 (slabs are BQ * 32 bytes apart; the start-address
    // field counts 16-byte units and never carries out of its 14 bits); of slab k of K stage s: kd0 + s * k_step + k * 128
    const uint64_t qd0 = umma_desc_sw32_kmajor(smem_q), kd0 = umma_desc_sw32_kmajor(smem_k);
    const uint64_t q_step = q_bytes >> 4, k_step = k_bytes >> 4, v_step = v_bytes >> 4;
    const bool v128 = p.v_atom == 64;
    const uint64_t vd0 = v128 ? umma_desc_sw128_mnmajor(smem_v, BN2 * 128u) : umma_desc_sw32_mnmajor(smem_v, BN2 * 32u);
    const uint64_t v_kstep = v128 ? (2048u >> 4) : (512u >> 4);            // 16 keys further down the V tile
    const uint64_t ones_d = tf::umma_desc_sw128_kmajor(smem_ones);
    const bool l_mma = p.l_mma != 0;
    const uint32_t ts0 = tmem_base, to0 = tmem_base + 2 * BN2, tcols = (uint32_t)p.tile_cols;
    int qs = 0;            // K/V stage and phase of the block whose S is issued next
    uint32_t qphase = 0;
    uint64_t kd = kd0;
    auto issue_qk = [&](int jj) {   // S_jj of both tiles; buffer jj % 2 held P_{jj-2}, whose P V MMAs were issued earlier (in order)
      att_wait(kv_full(qs), qphase, 2, jj);
      tf::tcgen05_fence_after();
      const uint32_t sb = (uint32_t)(jj & 1) * BN2;
      if (elected) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          uint64_t qd = qd0 + (uint64_t)t * q_step, kk = kd;
          for (int k = 0; k < slabs; ++k, qd += 256, kk += 128)
            tf::umma_f16_ss(ts0 + t * tcols + sb, qd, kk, idesc_s, k > 0 ? 1u : 0u);
          tf::umma_commit(s_full(t, jj & 1));
        }
      }
      __syncwarp();
      kd += k_step;
      if (++qs == ST) { qs = 0; qphase ^= 1u; kd = kd0; }
    };
    att_wait(q_full(0), 0, 3, 0);
    att_wait(q_full(1), 0, 4, 0);
    issue_qk(0);
    int vs = 0;
    uint64_t vd = vd0;
    for (int j = 0; j < nkv; ++j) {
      if (elected) ATT2_STAMP(2, 0);
      if (j + 1 < nkv) issue_qk(j + 1);
      if (elected) ATT2_STAMP(2, 1);
      const uint32_t sb = (uint32_t)(j & 1) * BN2;
      const uint32_t acc0 = j > 0 ? 1u : 0u;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        att_wait(p_full(t, j & 1), (uint32_t)(j >> 1) & 1u, 5 + t, j);
        tf::tcgen05_fence_after();
        if (elected) {
          ATT2_STAMP(2, 2 + 2 * t);
          const uint32_t pa = ts0 + t * tcols + sb;   // P_j: 64 fp16 per row in 32 columns
          const uint32_t od = to0 + t * tcols;
#pragma unroll
          for (int k = 0; k < BN2 / 16; ++k) {
            tf::umma_f16_ts(od, pa + 8u * k, vd + (uint64_t)k * v_kstep, idesc_o, k > 0 ? 1u : acc0);
            // L += P . ones, from the same fp16 P the numerator uses; issued after P V so a non-accumulating P V cannot
            // wipe it. Skipped when V itself carries a column of ones (the row sums then ARE column d of O).
            if (l_mma) tf::umma_f16_ts(od + p.l_off, pa + 8u * k, ones_d + 2u * k, idesc_l, k > 0 ? 1u : acc0);
          }
          tf::umma_commit(pv_done(t));
          if (j == nkv - 1) tf::umma_commit(pv_last(t));
          ATT2_STAMP(2, 3 + 2 * t);
        }
        __syncwarp();
      }
      if (elected) tf::umma_commit(kv_empty(vs));
      __syncwarp();
      vd += v_step;
      if (++vs == ST) { vs = 0; vd = vd0; }
    }
  } else if (warp >= 4) {
    // ===================== softmax / rescale / epilogue: warps 4-7 tile 0, warps 8-11 tile 1 =====================
    const int t = (warp - 4) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_field = (uint32_t)(q * 32) << 16;
    const int qrow0 = (2 * qt + t) * BQ;
    const float sc = p.scale_log2;
    float m_run = -INFINITY;   // reference max of this row (scaled log2 domain); lazily updated (see tf_attention_kernel)
    const uint32_t o_addr = tmem_o(t) + lane_field;
    const bool stamp = q == 0 && lane == 0;
    bool s_ready = false;   // S_j already seen complete by the probe issued in the middle of block j-1
    for (int j = 0; j < nkv; ++j) {
      if (stamp) ATT2_STAMP(t, 0);
      if (!s_ready) att_wait(s_full(t, j & 1), (uint32_t)(j >> 1) & 1u, 7 + t, j);
      tf::tcgen05_fence_after();
      if (stamp) ATT2_STAMP(t, 1);
      const uint32_t s_addr = tmem_s(t, j & 1) + lane_field;
      uint32_t v[64];
      tf::tmem_ld_x32(s_addr, v);
      tf::tmem_ld_x32(s_addr + 32, v + 32);
      tf::tmem_ld_wait();
      if (stamp) ATT2_STAMP(t, 2);
      int valid = min(BN2, p.Tk - j * BN2);
      if (p.causal) valid = min(valid, qrow0 + row + 1 - j * BN2);
      if (valid < BN2) {
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= valid) v[i] = 0xff800000u;   // -inf
      }
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int i = 0; i < 64; i += 8) {
#pragma unroll
        for (int u = 0; u < 4; ++u) mx4[u] = tf::fmax3(mx4[u], __uint_as_float(v[i + 2 * u]), __uint_as_float(v[i + 2 * u + 1]));
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * sc;   // scale > 0
      float alpha = 1.0f;
      if (mx > m_run + 8.0f) {          // also taken on the first block (m_run = -inf)
        alpha = exp2f(m_run - mx);      // 0 on the first block
        m_run = mx;
      }
      const float neg_m = -m_run;
      uint32_t pk[32];
      const uint64_t sc2 = tf::pack_f32x2(sc, sc), nm2 = tf::pack_f32x2(neg_m, neg_m);
      // (Forcing the two warpgroups to alternate on the exponential phase through named barriers was measured and dropped:
      // the barrier latency cost more than the MUFU collisions it avoided, profiles/attention2_variants_r2.log.)
      if (stamp) ATT2_STAMP(t, 3);
      if (stamp) ATT2_STAMP(t, 4);
#pragma unroll
      for (int i = 0; i < 64; i += 8) {
        float x[8];
#pragma unroll
        for (int u = 0; u < 8; u += 2) {
          const uint64_t xx = tf::fma_f32x2(tf::pack_f32x2(__uint_as_float(v[i + u]), __uint_as_float(v[i + u + 1])), sc2, nm2);
          tf::unpack_f32x2(xx, x[u], x[u + 1]);
        }
#pragma unroll
        for (int u = 0; u < 8; u += 2) {
          if (u >= 8 - EMU) {
            exp2_fma_pair(x[u], x[u + 1]);
          } else {
            x[u] = exp2f(x[u]);
            x[u + 1] = exp2f(x[u + 1]);
          }
          __half2 h2v = __floats2half2_rn(x[u], x[u + 1]);
          pk[(i + u) >> 1] = *reinterpret_cast<uint32_t*>(&h2v);
        }
      }
      if (stamp) ATT2_STAMP(t, 5);
      // Probe S_{j+1} now: the barrier unit answers in ~150 cycles even when the phase is long complete (measured: the wait at
      // the loop top was 10 % of a softmax warp's time, profiles/attention2_r2.md); here that latency hides behind the P store.
      s_ready = (j + 1 < nkv) && tf::mbar_test_wait(s_full(t, (j + 1) & 1), (uint32_t)((j + 1) >> 1) & 1u);
      // rescale the running output if any row of this warp moved its max (rare after the first blocks: lazy threshold 2^8)
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
        // O belongs to the tensor core until P V_{j-1} has completed. The parity wait below cannot alias an older phase:
        // S_j (which this thread has just consumed) was issued after P V_{j-2} of this tile and tcgen05.mma complete in issue
        // order, so every phase of pv_done up to j-2 is complete and the barrier is in phase j-1 or beyond.
        att_wait(pv_done(t), (uint32_t)(j - 1) & 1u, 9 + t, j);
        tf::tcgen05_fence_after();
        for (int c = 0; c < p.ol_cols; c += 16) {   // O and the L columns
          uint32_t o[16];
          tf::tmem_ld_x16(o_addr + c, o);
          tf::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tf::tmem_st_x16(o_addr + c, o);
        }
      }
      if (stamp) ATT2_STAMP(t, 6);
      tf::tmem_st_x32(s_addr, pk);   // P_j over the first 32 columns of the S buffer this thread has just drained
      tf::tmem_st_wait();
      tf::tcgen05_fence_before();
      tf::mbar_arrive(p_full(t, j & 1));
      if (stamp) ATT2_STAMP(t, 7);
    }
    // ---- epilogue: O / l -> fp16 -> global ----
    // The last P V has its own single-use barrier. pv_done cannot serve here: this thread has not followed its phases (it only
    // waits on it when a rescale is due), only phases up to nkv-3 are known complete, and a parity wait cannot tell phase
    // nkv-2, nkv-1 and "all done" apart - waiting for nkv-1 could pass on nkv-3, waiting for nkv-2 first could hang when both
    // have already completed.
    att_wait(pv_last(t), 0, 11 + t, nkv);
    tf::tcgen05_fence_after();
    float inv_l;
    {
      uint32_t l16[16];
      tf::tmem_ld_x16(o_addr + p.l_off, l16);
      tf::tmem_ld_wait();
      inv_l = 1.0f / __uint_as_float(p.l_sel ? l16[8] : l16[0]);
    }
    const int tq = qrow0 + row;
    __half* orow = p.out + (long long)b * p.osb + (long long)h * p.osh + (long long)tq * p.ost;
    for (int c = 0; c < p.d; c += 16) {
      uint32_t o[16];
      tf::tmem_ld_x16(o_addr + c, o);
      tf::tmem_ld_wait();
      if (tq < p.Tq) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (c + g * 8 < p.d) {
            tf::Pack16 pk8;
#pragma unroll
            for (int i = 0; i < 4; ++i)
              pk8.h2[i] = __floats2half2_rn(__uint_as_float(o[g * 8 + 2 * i]) * inv_l,
                                            __uint_as_float(o[g * 8 + 2 * i + 1]) * inv_l);
            *reinterpret_cast<uint4*>(orow + c + g * 8) = pk8.v;
          }
        }
      }
    }
  }

  tf::tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tf::tcgen05_fence_after();
    tf::tmem_dealloc(tmem_base, p.tmem_cols);
  }
}


// =====================================================================================================================
// tf_attention3_kernel: one query tile per CTA, every row split over TWO softmax threads, two CTAs per SM (round 2).
//
// What limited tf_attention2_kernel (ncu, profiles/attention2_r2.md): two softmax warps per scheduler issue on 53 % of the
// cycles and no pipe is above 40 % - MUFU 35 %, ALU 35 %, FMA 39 %, tensor 30 % - the warps wait on their own dependency
// chains and on barrier / tensor-memory latencies with nobody to cover for them. Here 16 softmax warps share an SM (four per
// scheduler) at the same registers per SM:
//   * a 128-row tile is served by 8 warps = 256 threads; thread (half, row) owns keys [32 half, 32 half + 32) of every 64-key
//     block of its row: 32 scores in registers, 16 packed P registers - 96 registers per thread, two 320-thread CTAs per SM;
//   * the two halves of a row NEVER talk inside the loop: each keeps its own running reference max and accumulates into its OWN
//     output accumulator, O_A += P_A V[keys 0..31], O_B += P_B V[keys 32..63] (two K = 16 TS MMAs each; the same tensor work as
//     one K = 64 product). The epilogue combines them: O = (2^(mA-m) O_A + 2^(mB-m) O_B) / (2^(mA-m) lA + 2^(mB-m) lB);
//   * S stays double-buffered and P stays in tensor memory (half A writes P over columns 0..15 of the buffer, half B over
//     columns 32..47: each over scores only it reads).
// TMEM per CTA (256 columns): [S buf 0 (64) | S buf 1 (64) | O_A (64) | O_B (64)]; the row sums live inside the O columns (a
// ones column of V, or the L MMA into V's zero pad columns), which is what restricts this kernel to head dims <= 48 (SD 1.x
// at 64x64 and 96x96 latents: d = 40) - the shapes that dominate the step.
//   warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 softmax (warp w: lane quarter w % 4, half (w - 2) / 4).
// =====================================================================================================================
constexpr int kA3Threads = 320;

struct Attn3Params {
  int B, NH, Tq, Tk, Tk_pad, d, dp, v_atom, l_off, l_sel, l_mma, resc_cols, nkv, stages, causal;
  float scale_log2;
  __half* out;
  long long osb, osh, ost;
};

template <int EMU>
__global__ void __launch_bounds__(kA3Threads, 2)
tf_attention3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const Attn3Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = tf::smem_u32(smem_raw);
  const uint32_t smem_base = (raw_u32 + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int ST = p.stages;
  constexpr int DOV = 64;                                   // O columns per accumulator = padded V head dim

  const uint32_t q_bytes = BQ * p.dp * 2;
  const uint32_t k_bytes = BN2 * p.dp * 2;
  const uint32_t v_bytes = BN2 * DOV * 2;
  const uint32_t smem_q = smem_base;
  const uint32_t smem_k = smem_q + q_bytes;
  const uint32_t smem_v = smem_k + ST * k_bytes;
  const uint32_t smem_ones = smem_v + ST * v_bytes;         // 16 x 64 fp16 ones: B operand of L = P . 1
  const uint32_t bar_base = smem_ones + 2048;
  const uint32_t q_full = bar_base;
  auto kv_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar_base + 8u * (1 + ST + s); };
  auto s_full = [&](int buf) { return bar_base + 8u * (1 + 2 * ST + buf); };
  auto p_full = [&](int buf) { return bar_base + 8u * (3 + 2 * ST + buf); };   // 256 arrivals; one per S buffer (see tf_attention2_kernel)
  const uint32_t pv_done = bar_base + 8u * (5 + 2 * ST);
  const uint32_t pv_last = bar_base + 8u * (6 + 2 * ST);
  const uint32_t tmem_slot = bar_base + 8u * (7 + 2 * ST);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw_u32));

  if (warp == 0 && lane == 0) {
    tf::tma_prefetch_desc(&tmQ);
    tf::tma_prefetch_desc(&tmK);
    tf::tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    tf::mbar_init(q_full, 1);
    for (int s = 0; s < ST; ++s) {
      tf::mbar_init(kv_full(s), 1);
      tf::mbar_init(kv_empty(s), 1);
    }
    for (int i = 0; i < 2; ++i) {
      tf::mbar_init(s_full(i), 1);
      tf::mbar_init(p_full(i), 256);
    }
    tf::mbar_init(pv_done, 1);
    tf::mbar_init(pv_last, 1);
    tf::fence_mbar_init();
  }
  if (warp == 2) {
    tf::tmem_alloc(tmem_slot, 256);
    tf::tmem_relinquish();
  }
  if (warp == 3) {   // 2 KB of fp16 1.0
    for (int i = lane; i < 128; i += 32)
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(smem_ones + i * 16u), "r"(0x3C003C00u) : "memory");
    tf::fence_proxy_async_smem();
  }
  tf::tcgen05_fence_before();
  __syncthreads();
  tf::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  tf::pdl_trigger();   // only once this CTA holds its tensor memory (see tf_gemm_kernel)
  tf::pdl_wait();

  const int nkv = p.nkv;
  const int slabs = p.dp / 16;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      tf::mbar_expect_tx(q_full, q_bytes);
      tf::tma_load_3d(smem_q, &tmQ, q_full, 0, b * p.Tq + qt * BQ, h * slabs);
      const int nblk = DOV / p.v_atom;
      for (int j = 0; j < nkv; ++j) {
        const int s = j % ST;
        const uint32_t u = j / ST;
        att_wait(kv_empty(s), (u & 1u) ^ 1u, 21, j);
        tf::mbar_expect_tx(kv_full(s), k_bytes + v_bytes);
        const int key0 = b * p.Tk_pad + j * BN2;
        tf::tma_load_3d(smem_k + s * k_bytes, &tmK, kv_full(s), 0, key0, h * slabs);
        for (int nb = 0; nb < nblk; ++nb)
          tf::tma_load_2d(smem_v + s * v_bytes + nb * (BN2 * p.v_atom * 2), &tmV, kv_full(s), h * DOV + nb * p.v_atom, key0);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp converged, one elected lane issues; see tf_attention2_kernel) =====================
    const bool elected = tf::elect_one();
    const uint32_t idesc_s = tf::umma_idesc_f16(BQ, BN2);
    const uint32_t idesc_o = tf::umma_idesc_f16(BQ, DOV) | (1u << 16);   // bit 16: B (V, natural layout) is MN-major
    const uint32_t idesc_l = tf::umma_idesc_f16(BQ, 16);
    const uint64_t qd0 = umma_desc_sw32_kmajor(smem_q), kd0 = umma_desc_sw32_kmajor(smem_k);
    const uint64_t k_step = k_bytes >> 4, v_step = v_bytes >> 4;
    const bool v128 = p.v_atom == 64;
    const uint64_t vd0 = v128 ? umma_desc_sw128_mnmajor(smem_v, BN2 * 128u) : umma_desc_sw32_mnmajor(smem_v, BN2 * 32u);
    const uint64_t v_kstep = v128 ? (2048u >> 4) : (512u >> 4);            // 16 keys further down the V tile
    const uint64_t ones_d = tf::umma_desc_sw128_kmajor(smem_ones);
    const bool l_mma = p.l_mma != 0;
    const uint32_t ts0 = tmem_base, to0 = tmem_base + 2 * BN2;
    int qs = 0;
    uint32_t qphase = 0;
    uint64_t kd = kd0;
    auto issue_qk = [&](int jj) {   // S_jj; buffer jj % 2 held P_{jj-2}, whose P V MMAs were issued earlier (in order)
      att_wait(kv_full(qs), qphase, 22, jj);
      tf::tcgen05_fence_after();
      if (elected) {
        uint64_t qd = qd0, kk = kd;
        for (int k = 0; k < slabs; ++k, qd += 256, kk += 128)
          tf::umma_f16_ss(ts0 + (uint32_t)(jj & 1) * BN2, qd, kk, idesc_s, k > 0 ? 1u : 0u);
        tf::umma_commit(s_full(jj & 1));
      }
      __syncwarp();
      kd += k_step;
      if (++qs == ST) { qs = 0; qphase ^= 1u; kd = kd0; }
    };
    att_wait(q_full, 0, 23, 0);
    issue_qk(0);
    int vs = 0;
    uint64_t vd = vd0;
    for (int j = 0; j < nkv; ++j) {
      if (j + 1 < nkv) issue_qk(j + 1);
      const uint32_t acc0 = j > 0 ? 1u : 0u;
      att_wait(p_full(j & 1), (uint32_t)(j >> 1) & 1u, 24, j);
      tf::tcgen05_fence_after();
      if (elected) {
        const uint32_t pa = ts0 + (uint32_t)(j & 1) * BN2;
#pragma unroll
        for (int k = 0; k < BN2 / 16; ++k) {
          const uint32_t hf = (uint32_t)k >> 1;                 // keys 16k .. 16k+15 belong to half hf
          const uint32_t pk_addr = pa + 32u * hf + 8u * ((uint32_t)k & 1u);   // P_A at columns 0..15, P_B at 32..47
          const uint32_t od = to0 + hf * DOV;
          const uint32_t acc = (k & 1) ? 1u : acc0;
          tf::umma_f16_ts(od, pk_addr, vd + (uint64_t)k * v_kstep, idesc_o, acc);
          if (l_mma) tf::umma_f16_ts(od + p.l_off, pk_addr, ones_d + 2u * k, idesc_l, acc);
        }
        tf::umma_commit(pv_done);
        if (j == nkv - 1) tf::umma_commit(pv_last);
        tf::umma_commit(kv_empty(vs));
      }
      __syncwarp();
      vd += v_step;
      if (++vs == ST) { vs = 0; vd = vd0; }
    }
  } else {
    // ===================== softmax / rescale / epilogue: warps 2-5 keys [0,32) of every block, warps 6-9 keys [32,64) =====================
    const int hf = (warp - 2) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_field = (uint32_t)(q * 32) << 16;
    const int qrow0 = qt * BQ;
    const float sc = p.scale_log2;
    float m_run = -INFINITY;   // reference max of this HALF row (scaled log2 domain), lazily updated
    const uint32_t o_addr = tmem_base + 2 * BN2 + (uint32_t)hf * DOV + lane_field;
    bool s_ready = false;
    for (int j = 0; j < nkv; ++j) {
      if (!s_ready) att_wait(s_full(j & 1), (uint32_t)(j >> 1) & 1u, 25, j);
      tf::tcgen05_fence_after();
      const uint32_t s_addr = tmem_base + (uint32_t)(j & 1) * BN2 + 32u * (uint32_t)hf + lane_field;
      uint32_t v[32];
      tf::tmem_ld_x32(s_addr, v);
      tf::tmem_ld_wait();
      int valid = min(BN2, p.Tk - j * BN2);
      if (p.causal) valid = min(valid, qrow0 + row + 1 - j * BN2);
      valid -= 32 * hf;
      if (valid < 32) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i >= valid) v[i] = 0xff800000u;   // -inf
      }
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
#pragma unroll
        for (int u = 0; u < 4; ++u) mx4[u] = tf::fmax3(mx4[u], __uint_as_float(v[i + 2 * u]), __uint_as_float(v[i + 2 * u + 1]));
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * sc;   // scale > 0
      float alpha = 1.0f;
      if (mx > m_run + 8.0f) {          // also taken on the first block with a visible key (m_run = -inf)
        alpha = exp2f(m_run - mx);      // 0 then
        m_run = mx;
      }
      // a half row that has not seen a visible key yet (causal mask / ragged tail): every score is -inf, P must be 0, not NaN
      const float neg_m = m_run == -INFINITY ? 0.f : -m_run;
      uint32_t pk[16];
      const uint64_t sc2 = tf::pack_f32x2(sc, sc), nm2 = tf::pack_f32x2(neg_m, neg_m);
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        float x[8];
#pragma unroll
        for (int u = 0; u < 8; u += 2) {
          const uint64_t xx = tf::fma_f32x2(tf::pack_f32x2(__uint_as_float(v[i + u]), __uint_as_float(v[i + u + 1])), sc2, nm2);
          tf::unpack_f32x2(xx, x[u], x[u + 1]);
        }
#pragma unroll
        for (int u = 0; u < 8; u += 2) {
          if (u >= 8 - EMU) {
            exp2_fma_pair(x[u], x[u + 1]);
          } else {
            x[u] = exp2f(x[u]);
            x[u + 1] = exp2f(x[u + 1]);
          }
          __half2 h2v = __floats2half2_rn(x[u], x[u + 1]);
          pk[(i + u) >> 1] = *reinterpret_cast<uint32_t*>(&h2v);
        }
      }
      s_ready = (j + 1 < nkv) && tf::mbar_test_wait(s_full((j + 1) & 1), (uint32_t)((j + 1) >> 1) & 1u);
      // rescale this half's accumulator if any row of the warp moved its max (rare after the first blocks: threshold 2^8)
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
        // O belongs to the tensor core until P V_{j-1} has completed; the parity wait cannot alias (see tf_attention2_kernel)
        att_wait(pv_done, (uint32_t)(j - 1) & 1u, 26, j);
        tf::tcgen05_fence_after();
        for (int c = 0; c < p.resc_cols; c += 16) {
          uint32_t o[16];
          tf::tmem_ld_x16(o_addr + c, o);
          tf::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tf::tmem_st_x16(o_addr + c, o);
        }
      }
      tf::tmem_st_x16(s_addr, pk);   // P over the first 16 columns of the 32 score columns this thread has just drained
      tf::tmem_st_wait();
      tf::tcgen05_fence_before();
      tf::mbar_arrive(p_full(j & 1));
    }
    // ---- epilogue: combine the two halves of every row, O / l -> fp16 -> global ----
    att_wait(pv_last, 0, 27, nkv);
    tf::tcgen05_fence_after();
    float l_own;
    {
      uint32_t l16[16];
      tf::tmem_ld_x16(o_addr + p.l_off, l16);
      tf::tmem_ld_wait();
      l_own = __uint_as_float(p.l_sel ? l16[8] : l16[0]);
    }
    // exchange through shared memory (the Q tile and the K ring are dead: every MMA that read them has completed): row r of half
    // B leaves {m, l, O[0..d)} at pitch d + 2 floats (d + 2 = 2 mod 8 words: the rows of a warp spread over the banks)
    const int pitch = p.d + 2;
    float* xch = reinterpret_cast<float*>(smem_raw + (smem_q - raw_u32)) + (size_t)row * pitch;
    if (hf == 1) {
      xch[0] = m_run;
      xch[1] = l_own;
      for (int c = 0; c < p.d; c += 16) {
        uint32_t o[16];
        tf::tmem_ld_x16(o_addr + c, o);
        tf::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (c + i < p.d) xch[2 + c + i] = __uint_as_float(o[i]);
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (hf == 0) {
      const float m_b = xch[0], l_b = xch[1];
      const float m = fmaxf(m_run, m_b);               // finite: every row sees at least one key
      const float w_a = exp2f(m_run - m), w_b = exp2f(m_b - m);   // exp2(-inf) = 0: a half that never saw a key drops out
      const float inv_l = 1.0f / (w_a * l_own + w_b * l_b);
      const float f_a = w_a * inv_l, f_b = w_b * inv_l;
      const int tq = qrow0 + row;
      __half* orow = p.out + (long long)b * p.osb + (long long)h * p.osh + (long long)tq * p.ost;
      for (int c = 0; c < p.d; c += 16) {
        uint32_t o[16];
        tf::tmem_ld_x16(o_addr + c, o);
        tf::tmem_ld_wait();
        if (tq < p.Tq) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            if (c + g * 8 < p.d) {
              tf::Pack16 pk8;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int e = c + g * 8 + 2 * i;
                pk8.h2[i] = __floats2half2_rn(fmaf(__uint_as_float(o[g * 8 + 2 * i]), f_a, xch[2 + e] * f_b),
                                              fmaf(__uint_as_float(o[g * 8 + 2 * i + 1]), f_a, xch[3 + e] * f_b));
              }
              *reinterpret_cast<uint4*>(orow + c + g * 8) = pk8.v;
            }
          }
        }
      }
    }
  }

  tf::tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tf::tcgen05_fence_after();
    tf::tmem_dealloc(tmem_base, 256);
  }
}

int g_attn_version = 0;   // 0 auto, 1 tf_attention_kernel only, 2 / 3: tf_attention2_kernel / tf_attention3_kernel wherever it applies
// auto mode: tf_attention3_kernel when its grid is ONE wave of two CTAs per SM (4096 tokens x 8 heads x batch 2 = 512 tiles on
// 592 slots: 99 us vs 107 us for the two-tile kernel, whose 256 CTAs take two waves of 148); on longer grids the two kernels
// run at the same ~770 cycles per 128 x 64 score tile per SM and the two-tile kernel issues half the row-sum MMAs
// (profiles/attention3_r2.log)
long g_attn3_min_tiles = 300;
// Measured on B200 (tools/dev_attn2.py, profiles/attention2_r2.md), 4096 tokens x d = 40, batch 2, us per call: one-tile kernel
// 133.3; this kernel with 0 / 2 / 4 of 8 exponentials on the FMA pipe 121.3 / 111.3 / 110.0 with the warpgroup order, 108.5 /
// 105.0 (2 / 4) without it: forcing the two tiles to alternate costs more in barrier latency than the MUFU collisions it avoids.
int g_attn_emu = -1;      // exponentials per 8 evaluated on the FMA pipe (0, 2 or 4); < 0: each kernel's measured best

int g_force_attn_bn = 0;
int g_force_attn_occ = 0;
long long* g_attn_timeline = nullptr;

}  // namespace

extern "C" int tf_attention_set_debug(void* host_mapped_buf) {
  unsigned long long* p = reinterpret_cast<unsigned long long*>(host_mapped_buf);   // >= 8 x 8 bytes, zeroed, pinned; or NULL
  TF_CUDA(cudaMemcpyToSymbol(g_att_dbg, &p, sizeof(p)));
  return TF_OK;
}

extern "C" int tf_attention_set_timeline(long long* dev_buf) {
  g_attn_timeline = dev_buf;   // >= 64 * 8 int64; debug only (stamps exist in TF_ATT_TRACE builds)
  return TF_OK;
}

extern "C" int tf_attention_set_variant(int version, int emu) {
  TF_CHECK_ARG(version >= 0 && version <= 3 && (emu < 0 || emu == 0 || emu == 2 || emu == 4),
               "tf_attention_set_variant: version in {0 auto, 1, 2, 3}, emu in {0, 2, 4} (< 0: each kernel's default)");
  g_attn_version = version;
  g_attn_emu = emu;   // < 0: back to each kernel's own default
  return TF_OK;
}

extern "C" int tf_attention_set_tuning(int force_bn) {
  g_force_attn_occ = force_bn / 1000;   // thousands digit: CTAs per SM (0 = auto), e.g. 3064 = 64-key blocks, 3 CTAs / SM
  force_bn %= 1000;
  g_force_attn_bn = force_bn;
  return TF_OK;
}

// vnat == 0: `vt` is V transposed, (NH*dp, ldvt >= B*Tk_pad). vnat == 1: `vt` is V in its natural layout,
// (B*Tk_pad, ldvt >= NH*dvp), head h at columns [h*dvp, (h+1)*dvp), dvp % 64 == 0 (zero pad columns).
static int attention_impl(const void* q, int ldq, const void* k, int ldk, const void* vt, int ldvt, void* out,
                          long long out_stride_b, long long out_stride_h, long long out_stride_t, int B,
                          int NH, int Tq, int Tk, int Tk_pad, int d, int dp, float scale, int causal, int vnat, int dvp,
                          void* stream_, int ones_col = 0) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int dov = vnat ? dvp : dp;   // O columns
  // V's pad columns are exact zeros, so the PV MMA adds exact zeros to O's pad columns: when at least 16 of them lie beyond
  // the last 16-column group that holds real head-dim elements, the L = P . 1 accumulator lives there (issued after PV in
  // every k-step, so the first, non-accumulating PV MMA cannot wipe it) and TMEM needs BN + dov columns instead of + 16 more
  const int d16 = (d + 15) / 16 * 16;
  // ones_col: column d of every V head holds 1.0 (the projection GEMM's bias puts it there), so column d of O = P V IS the row
  // sum: tf_attention2_kernel then issues no L MMA at all; tf_attention_kernel keeps its own L columns clear of it
  const int l_off = ones_col ? dov : ((vnat && dov - 16 >= d16) ? dov - 16 : dov);
  const int ol_cols = l_off + 16 > dov ? l_off + 16 : dov;
  if (ones_col) TF_CHECK_ARG(vnat && dvp > d, "tf_attention_v_f16: the ones column needs a padded V head (dvp %d > d %d)", dvp, d);
  TF_CHECK_ARG(q && k && vt && out, "tf_attention_f16: null pointer");
  TF_CHECK_ARG(B > 0 && NH > 0 && Tq > 0 && Tk > 0 && Tk_pad >= Tk, "tf_attention_f16: bad dims");
  TF_CHECK_ARG(dp % 16 == 0 && dp >= 16 && dp <= 256 && d <= dp && d % 8 == 0,
               "tf_attention_f16: head dim d=%d (padded %d) unsupported: need d %% 8 == 0, dp %% 16 == 0, dp <= 256", d, dp);
  // Tk_pad % 8: batch b's V^T columns start at b*Tk_pad and a TMA box must start on a 16-byte boundary
  TF_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldvt % 8 == 0 && (vnat || Tk_pad % 8 == 0),
               "tf_attention_f16: leading dims and Tk_pad must be multiples of 8 (ldq=%d ldk=%d ldvt=%d Tk_pad=%d)", ldq,
               ldk, ldvt, Tk_pad);
  TF_CHECK_ARG(ldq >= NH * dp && ldk >= NH * dp && ldvt >= (vnat ? NH * dvp : B * Tk_pad),
               "tf_attention_f16: leading dims too small");
  if (vnat) TF_CHECK_ARG(dvp % 16 == 0 && dvp >= d && dvp <= 256, "tf_attention_v_f16: V head dim must be padded to a multiple of 16 (<= 256), got %d", dvp);
  const int v_atom = (vnat && dvp % 64 == 0) ? 64 : 16;
  TF_CHECK_ARG(out_stride_t % 8 == 0 && out_stride_h % 8 == 0 && out_stride_b % 8 == 0,
               "tf_attention_f16: output strides must be multiples of 8");
  TF_CHECK_ARG(((uintptr_t)q & 15) == 0 && ((uintptr_t)k & 15) == 0 && ((uintptr_t)vt & 15) == 0 &&
                   ((uintptr_t)out & 15) == 0,
               "tf_attention_f16: pointers must be 16-byte aligned");

  // ---- split-row kernel (tf_attention3_kernel): natural-layout V padded to 64 columns with the row sums inside them (head
  // dims <= 48), whole 128-row query tiles, and enough tiles to give every SM its two CTAs ----
  {
    const long tiles = (long)(Tq / BQ) * NH * B;
    const bool can3 = vnat && dov == 64 && Tq % BQ == 0 && Tk >= 2 * BN2 && (ones_col ? d < 64 : d16 + 16 <= 64);
    const bool want3 = g_attn_version == 3 || (g_attn_version == 0 && tiles >= g_attn3_min_tiles && tiles <= 4L * tf_num_sms());
    if (can3 && want3) {
      Attn3Params p3{};
      p3.B = B; p3.NH = NH; p3.Tq = Tq; p3.Tk = Tk; p3.Tk_pad = Tk_pad; p3.d = d; p3.dp = dp;
      p3.v_atom = v_atom; p3.causal = causal ? 1 : 0;
      p3.l_mma = ones_col ? 0 : 1;
      p3.l_off = ones_col ? (d & ~15) : 48;
      p3.l_sel = ones_col ? (d & 15) : 0;
      p3.resc_cols = ones_col ? (d & ~15) + 16 : 64;
      p3.nkv = ceil_div_i(Tk, BN2);
      p3.scale_log2 = scale * 1.4426950408889634f;
      p3.out = reinterpret_cast<__half*>(out);
      p3.osb = out_stride_b; p3.osh = out_stride_h; p3.ost = out_stride_t;
      const size_t qb = (size_t)BQ * dp * 2, kb = (size_t)BN2 * dp * 2, vb = (size_t)BN2 * 64 * 2;
      const long budget = (long)112 * 1024;               // two CTAs per SM
      int st3 = (int)((budget - 1024 - (long)qb - 2048 - 1024) / (long)(kb + vb));
      if (st3 > 6) st3 = 6;
      if (st3 > p3.nkv) st3 = p3.nkv;
      const size_t xch = (size_t)BQ * (d + 2) * 4;        // epilogue exchange, over the Q tile and the K ring
      if (st3 >= 2 && xch <= qb + (size_t)st3 * kb) {
        p3.stages = st3;
        const size_t smem3 = 1024 + qb + (size_t)st3 * (kb + vb) + 2048 + 1024;
        CUtensorMap tmQ, tmK, tmV;
        {
          uint64_t dims[3] = {16, (uint64_t)B * Tq, (uint64_t)(ldq / 16)};
          uint64_t strides[2] = {(uint64_t)ldq * 2, 32};
          uint32_t box[3] = {16, BQ, (uint32_t)(dp / 16)};
          uint32_t es[3] = {1, 1, 1};
          int rc = tf_encode_tmap(&tmQ, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, q, dims, strides, box, es, CU_TENSOR_MAP_SWIZZLE_32B);
          if (rc) return rc;
        }
        {
          uint64_t dims[3] = {16, (uint64_t)B * Tk_pad, (uint64_t)(ldk / 16)};
          uint64_t strides[2] = {(uint64_t)ldk * 2, 32};
          uint32_t box[3] = {16, (uint32_t)BN2, (uint32_t)(dp / 16)};
          uint32_t es[3] = {1, 1, 1};
          int rc = tf_encode_tmap(&tmK, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, k, dims, strides, box, es, CU_TENSOR_MAP_SWIZZLE_32B);
          if (rc) return rc;
        }
        {
          uint64_t dims[2] = {(uint64_t)ldvt, (uint64_t)B * Tk_pad};
          uint64_t strides[1] = {(uint64_t)ldvt * 2};
          uint32_t box[2] = {(uint32_t)v_atom, (uint32_t)BN2};
          uint32_t es[2] = {1, 1};
          int rc = tf_encode_tmap(&tmV, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, vt, dims, strides, box, es,
                                  v_atom == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B);
          if (rc) return rc;
        }
        static bool attr3_set = false;
        if (!attr3_set) {
          TF_CUDA(cudaFuncSetAttribute(tf_attention3_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
          TF_CUDA(cudaFuncSetAttribute(tf_attention3_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
          TF_CUDA(cudaFuncSetAttribute(tf_attention3_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
          attr3_set = true;
        }
        dim3 grid3(Tq / BQ, NH, B);
        const int emu3 = g_attn_emu < 0 ? 2 : g_attn_emu;     // measured best: 2 of 8 (99.2 us; 4 of 8: 103.3, none: 112.8)
        if (emu3 == 0) TF_LAUNCH((tf_attention3_kernel<0>), grid3, kA3Threads, smem3, stream, tmQ, tmK, tmV, p3);
        else if (emu3 == 4) TF_LAUNCH((tf_attention3_kernel<4>), grid3, kA3Threads, smem3, stream, tmQ, tmK, tmV, p3);
        else TF_LAUNCH((tf_attention3_kernel<2>), grid3, kA3Threads, smem3, stream, tmQ, tmK, tmV, p3);
        TF_LAUNCH_CHECK();
        tf_launch_count_add(1);
        return TF_OK;
      }
    }
  }
  // ---- two-tile ping-pong kernel (tf_attention2_kernel): natural-layout V, whole pairs of 128-row query tiles, and a grid
  // that gives (nearly) every SM a CTA; everything else stays on tf_attention_kernel ----
  {
    const long pairs = (long)(Tq / (2 * BQ)) * NH * B;
    const int ol2 = ones_col ? dov : ol_cols;     // TMEM columns of O (+ L) per tile in the two-tile kernel
    const bool can2 = vnat && Tq % (2 * BQ) == 0 && Tk >= 2 * BN2 && 2 * BN2 + ol2 <= 256;
    const bool want2 = g_attn_version == 2 || (g_attn_version == 0 && pairs >= 120);
    if (can2 && want2) {
      Attn2Params p2{};
      p2.B = B; p2.NH = NH; p2.Tq = Tq; p2.Tk = Tk; p2.Tk_pad = Tk_pad; p2.d = d; p2.dp = dp; p2.dov = dov;
      p2.v_atom = v_atom; p2.causal = causal ? 1 : 0;
      p2.l_mma = ones_col ? 0 : 1;
      p2.l_off = ones_col ? (d & ~15) : l_off;      // 16-column group that holds the row sum ...
      p2.l_sel = ones_col ? (d & 15) : 0;           // ... and its position inside the group (d % 8 == 0: 0 or 8)
      p2.ol_cols = ol2;
      p2.nkv = ceil_div_i(Tk, BN2);
      p2.timeline = g_attn_timeline;
      p2.tile_cols = 2 * BN2 + ol2;
      uint32_t cols2 = 32;
      while (cols2 < 2u * (uint32_t)p2.tile_cols) cols2 <<= 1;
      p2.tmem_cols = cols2;
      p2.scale_log2 = scale * 1.4426950408889634f;
      p2.out = reinterpret_cast<__half*>(out);
      p2.osb = out_stride_b; p2.osh = out_stride_h; p2.ost = out_stride_t;
      const size_t qb = (size_t)BQ * dp * 2, kb = (size_t)BN2 * dp * 2, vb = (size_t)BN2 * dov * 2;
      long room = (long)227 * 1024 - 2048 - (long)(2 * qb) - 2048 - 1024;
      int st2 = (int)(room / (long)(kb + vb));
      if (st2 > 6) st2 = 6;
      if (st2 > p2.nkv) st2 = p2.nkv;
      if (st2 >= 2) {
        p2.stages = st2;
        size_t smem2 = 2 * qb + (size_t)st2 * (kb + vb) + 2048 + 1024 + 1024;
        if (smem2 < (size_t)120 * 1024) smem2 = (size_t)120 * 1024;   // one CTA per SM: the 512-column TMEM allocation is never contended
        CUtensorMap tmQ, tmK, tmV;
        {
          uint64_t dims[3] = {16, (uint64_t)B * Tq, (uint64_t)(ldq / 16)};
          uint64_t strides[2] = {(uint64_t)ldq * 2, 32};
          uint32_t box[3] = {16, BQ, (uint32_t)(dp / 16)};
          uint32_t es[3] = {1, 1, 1};
          int rc = tf_encode_tmap(&tmQ, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, q, dims, strides, box, es, CU_TENSOR_MAP_SWIZZLE_32B);
          if (rc) return rc;
        }
        {
          uint64_t dims[3] = {16, (uint64_t)B * Tk_pad, (uint64_t)(ldk / 16)};
          uint64_t strides[2] = {(uint64_t)ldk * 2, 32};
          uint32_t box[3] = {16, (uint32_t)BN2, (uint32_t)(dp / 16)};
          uint32_t es[3] = {1, 1, 1};
          int rc = tf_encode_tmap(&tmK, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, k, dims, strides, box, es, CU_TENSOR_MAP_SWIZZLE_32B);
          if (rc) return rc;
        }
        {
          uint64_t dims[2] = {(uint64_t)ldvt, (uint64_t)B * Tk_pad};
          uint64_t strides[1] = {(uint64_t)ldvt * 2};
          uint32_t box[2] = {(uint32_t)v_atom, (uint32_t)BN2};
          uint32_t es[2] = {1, 1};
          int rc = tf_encode_tmap(&tmV, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, vt, dims, strides, box, es,
                                  v_atom == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B);
          if (rc) return rc;
        }
        static bool attr2_set = false;
        if (!attr2_set) {
          TF_CUDA(cudaFuncSetAttribute(tf_attention2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
          TF_CUDA(cudaFuncSetAttribute(tf_attention2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
          TF_CUDA(cudaFuncSetAttribute(tf_attention2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
          attr2_set = true;
        }
        dim3 grid2(Tq / (2 * BQ), NH, B);
        const int emu2 = g_attn_emu < 0 ? 4 : g_attn_emu;
        if (emu2 == 0) TF_LAUNCH((tf_attention2_kernel<0>), grid2, kA2Threads, smem2, stream, tmQ, tmK, tmV, p2);
        else if (emu2 == 4) TF_LAUNCH((tf_attention2_kernel<4>), grid2, kA2Threads, smem2, stream, tmQ, tmK, tmV, p2);
        else TF_LAUNCH((tf_attention2_kernel<2>), grid2, kA2Threads, smem2, stream, tmQ, tmK, tmV, p2);
        TF_LAUNCH_CHECK();
        tf_launch_count_add(1);
        return TF_OK;
      }
    }
  }
  // 128-key blocks when two CTAs per SM still fit (TMEM <= 256 columns, >= 2 K/V stages in half an SM's shared
  // memory) and the sequence is long enough to amortise them; 64-key blocks otherwise
  auto fits2 = [&](int bn) {
    const size_t need = (size_t)BQ * dp * 2 + (size_t)BQ * bn * 2 + 4096 + 2 * (size_t)(bn * (dp + dov) * 2);
    return bn + ol_cols <= 256 && need <= (size_t)113 * 1024 - 2048;
  };
  // measured on B200 (tools/dev_attn_variants.py): 2 CTAs/SM x 128 keys beats 2 x 64 by 1-2 % at 4096 tokens;
  // 1 CTA/SM x 128 keys beats 64 by 23 % when the grid has no second CTA per SM to offer (1024 tokens, d = 80);
  // 3 CTAs/SM x 64 keys sustains 19 % more (444 vs 373 TFLOP/s) but only pays once the grid is several waves long
  // (a 512-CTA grid on 444 slots runs one slow wave plus a full-length tail)
  const long grid_ctas = (long)ceil_div_i(Tq, BQ) * NH * B;
  auto fits1 = [&](int bn) {
    const size_t need = (size_t)BQ * dp * 2 + (size_t)BQ * bn * 2 + 4096 + 2 * (size_t)(bn * (dp + dov) * 2);
    return bn + ol_cols <= 512 && need <= (size_t)227 * 1024 - 2048;
  };
  auto fits3 = [&]() {
    const size_t need = (size_t)BQ * dp * 2 + (size_t)BQ * 64 * 2 + 4096 + 2 * (size_t)(64 * (dp + dov) * 2);
    return 64 + ol_cols <= 170 && need <= (size_t)75 * 1024 - 1024;   // 3 x 170 columns <= 512 (allocations are powers of two: 128)
  };
  int BN = 64, occ = 2;
  if (Tk >= 512 && (fits2(128) || (grid_ctas <= tf_num_sms() && fits1(128)))) BN = 128;
  if (Tk >= 512 && fits3() && 64 + ol_cols <= 128 && grid_ctas >= 1024) { BN = 64; occ = 3; }
  if (g_force_attn_bn == 64 || g_force_attn_bn == 128) BN = g_force_attn_bn;
  if (g_force_attn_occ == 2 || (g_force_attn_occ == 3 && BN == 64 && fits3() && 64 + ol_cols <= 128)) occ = g_force_attn_occ;
  if (BN == 128) occ = 2;
  if (BN == 128 && 128 + ol_cols > 512) BN = 64;

  AttnParams p{};
  p.B = B; p.NH = NH; p.Tq = Tq; p.Tk = Tk; p.Tk_pad = Tk_pad; p.d = d; p.dp = dp; p.dov = dov;
  p.l_off = l_off; p.ol_cols = ol_cols; p.v_atom = v_atom;
  p.nkv = ceil_div_i(Tk, BN);
  p.causal = causal ? 1 : 0;
  p.timeline = g_attn_timeline;
  uint32_t need = BN + ol_cols, cols = 32;   // S, O, L
  while (cols < need) cols <<= 1;
  p.tmem_cols = cols;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__half*>(out);
  p.osb = out_stride_b; p.osh = out_stride_h; p.ost = out_stride_t;

  const size_t q_bytes = (size_t)BQ * dp * 2, k_bytes = (size_t)BN * dp * 2, v_bytes = (size_t)BN * dov * 2,
               p_bytes = (size_t)BQ * BN * 2;
  // half an SM's shared memory (two CTAs per SM) when that still holds a K/V ring of >= 2 stages - the MMA warp
  // issues S_{j+1} before PV_j, so block j+1 must land while block j's stage is still in use - else the whole SM
  auto ring = [&](size_t budget) {
    const long room = (long)budget - (long)q_bytes - (long)p_bytes - 2048;
    return room < 0 ? 0 : (int)(room / (long)(k_bytes + v_bytes));
  };
  int stages = occ == 3 ? ring((size_t)75 * 1024 - 1024) : (cols <= 256 ? ring((size_t)113 * 1024 - 2048) : 0);
  if (stages < 2 && stages < p.nkv) stages = ring((size_t)227 * 1024 - 2048);
  if (stages > 4) stages = 4;
  if (stages > p.nkv) stages = p.nkv < 1 ? 1 : p.nkv;
  TF_CHECK_ARG(stages >= 2 || (stages == 1 && p.nkv == 1), "tf_attention_f16: head dim %d does not fit shared memory", dp);
  p.stages = stages;
  const size_t smem = q_bytes + p_bytes + 2048 + (size_t)stages * (k_bytes + v_bytes) + 2048;

  CUtensorMap tmQ, tmK, tmV;
  {
    uint64_t dims[3] = {16, (uint64_t)B * Tq, (uint64_t)(ldq / 16)};
    uint64_t strides[2] = {(uint64_t)ldq * 2, 32};
    uint32_t box[3] = {16, BQ, (uint32_t)(dp / 16)};
    uint32_t es[3] = {1, 1, 1};
    int rc = tf_encode_tmap(&tmQ, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, q, dims, strides, box, es,
                            CU_TENSOR_MAP_SWIZZLE_32B);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {16, (uint64_t)B * Tk_pad, (uint64_t)(ldk / 16)};
    uint64_t strides[2] = {(uint64_t)ldk * 2, 32};
    uint32_t box[3] = {16, (uint32_t)BN, (uint32_t)(dp / 16)};
    uint32_t es[3] = {1, 1, 1};
    int rc = tf_encode_tmap(&tmK, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, k, dims, strides, box, es,
                            CU_TENSOR_MAP_SWIZZLE_32B);
    if (rc) return rc;
  }
  if (vnat) {
    uint64_t dims[2] = {(uint64_t)ldvt, (uint64_t)B * Tk_pad};
    uint64_t strides[1] = {(uint64_t)ldvt * 2};
    uint32_t box[2] = {(uint32_t)v_atom, (uint32_t)BN};
    uint32_t es[2] = {1, 1};
    int rc = tf_encode_tmap(&tmV, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, vt, dims, strides, box, es,
                            v_atom == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B);
    if (rc) return rc;
  } else {
    uint64_t dims[2] = {(uint64_t)B * Tk_pad, (uint64_t)NH * dp};
    uint64_t strides[1] = {(uint64_t)ldvt * 2};
    uint32_t box[2] = {64, (uint32_t)dp};
    uint32_t es[2] = {1, 1};
    int rc = tf_encode_tmap(&tmV, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, vt, dims, strides, box, es,
                            CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  dim3 grid(ceil_div_i(Tq, BQ), NH, B);
  static bool attr_set = false;
  if (!attr_set) {
    TF_CUDA(cudaFuncSetAttribute(tf_attention_kernel<64, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    TF_CUDA(cudaFuncSetAttribute(tf_attention_kernel<64, 3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    TF_CUDA(cudaFuncSetAttribute(tf_attention_kernel<128, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    TF_CUDA(cudaFuncSetAttribute(tf_attention_kernel<64, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    TF_CUDA(cudaFuncSetAttribute(tf_attention_kernel<64, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    TF_CUDA(cudaFuncSetAttribute(tf_attention_kernel<128, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
#define TF_ATT_LAUNCH(BN_, OCC_)                                                                                          \
  do {                                                                                                                    \
    if (vnat) TF_LAUNCH((tf_attention_kernel<BN_, OCC_, true>), grid, kAttThreads, smem, stream, tmQ, tmK, tmV, p);        \
    else TF_LAUNCH((tf_attention_kernel<BN_, OCC_, false>), grid, kAttThreads, smem, stream, tmQ, tmK, tmV, p);            \
  } while (0)
  if (BN == 64 && occ == 3) TF_ATT_LAUNCH(64, 3);
  else if (BN == 64) TF_ATT_LAUNCH(64, 2);
  else TF_ATT_LAUNCH(128, 2);
#undef TF_ATT_LAUNCH
  TF_LAUNCH_CHECK();
  tf_launch_count_add(1);
  return TF_OK;
}

extern "C" int tf_attention_f16(const void* q, int ldq, const void* k, int ldk, const void* vt, int ldvt, void* out,
                                long long out_stride_b, long long out_stride_h, long long out_stride_t, int B,
                                int NH, int Tq, int Tk, int Tk_pad, int d, int dp, float scale, void* stream) {
  return attention_impl(q, ldq, k, ldk, vt, ldvt, out, out_stride_b, out_stride_h, out_stride_t, B, NH, Tq, Tk, Tk_pad, d, dp,
                        scale, 0, 0, 0, stream);
}

extern "C" int tf_attention_v_f16(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out,
                                  long long out_stride_b, long long out_stride_h, long long out_stride_t, int B, int NH,
                                  int Tq, int Tk, int Tk_pad, int d, int dp, int dvp, float scale, int causal, void* stream) {
  const int ones_col = (causal >> 1) & 1;   // flags: bit 0 causal, bit 1 TF_ATTN_V_ONES_COLUMN
  causal &= 1;
  if (causal) TF_CHECK_ARG(Tq == Tk, "tf_attention_v_f16: causal masking needs Tq == Tk (got %d, %d)", Tq, Tk);
  return attention_impl(q, ldq, k, ldk, v, ldv, out, out_stride_b, out_stride_h, out_stride_t, B, NH, Tq, Tk, Tk_pad, d, dp,
                        scale, causal, 1, dvp, stream, ones_col);
}

extern "C" int tf_attention_causal_f16(const void* q, int ldq, const void* k, int ldk, const void* vt, int ldvt, void* out,
                                       long long out_stride_b, long long out_stride_h, long long out_stride_t, int B,
                                       int NH, int Tq, int Tk, int Tk_pad, int d, int dp, float scale, void* stream) {
  TF_CHECK_ARG(Tq == Tk, "tf_attention_causal_f16: causal masking needs Tq == Tk (got %d, %d)", Tq, Tk);
  return attention_impl(q, ldq, k, ldk, vt, ldvt, out, out_stride_b, out_stride_h, out_stride_t, B, NH, Tq, Tk, Tk_pad, d, dp,
                        scale, 1, 0, 0, stream);
}
