/* tinyfusers_b200 — C ABI of libtinyfusers_b200.so (hand-written sm_100a kernels).
 *
 * This header is the drop-in boundary for the denoising hot path of Fatlonder/tinyfusers: every
 * entry point below replaces one library call (cuDNN / cuBLAS / CuPy-JIT) that the reference's L2
 * operators make today. Citations are relative to the reference repository root.
 *
 * Conventions (same style as the reference's own ctypes bindings, tinyfusers/native/cublas/ops.py:3-54):
 *   - plain C types only: device pointers as void*, sizes as int, the CUDA stream as void*
 *     (cudaStream_t); no torch / CuPy types cross this boundary;
 *   - every function returns int: 0 = success, < 0 = argument / support error, > 0 = cudaError_t;
 *     tf_last_error() returns the thread-local message (the reference raises RuntimeError from the
 *     status code, tinyfusers/ff/linear.py:100-103 — the Python wrappers here do the same);
 *   - the callee never allocates: outputs and workspaces are caller-owned; nothing synchronises the
 *     device; all work is enqueued on `stream`, so every call is CUDA-graph capturable;
 *   - activations are fp16, NHWC for images ( == (B, T, C) for token sequences ), accumulation,
 *     statistics and softmax are fp32. Weights are fp16 in the layouts stated per function.
 *   - kernels are chained by programmatic dependent launch: each one lets its successor start launching on entry and
 *     waits for its predecessor before it reads ACTIVATIONS. PARAMETERS (norm gamma / beta, the Cin = 4 conv weights,
 *     and W with TF_GEMM_W_STATIC) are fetched before that wait, so they must be complete in memory when the call is
 *     enqueued - i.e. not written by the kernel enqueued immediately before it on the same stream (the Python layer
 *     synchronises once after packing weights);
 *   - there is NO CPU fallback. On a machine without an sm_100 GPU tf_init() fails.
 */
#ifndef TINYFUSERS_B200_H_
#define TINYFUSERS_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library state ------------------------------------------------------------------------- */
int tf_version(void);
/* Selects `device`, verifies compute capability 10.x. Replaces the per-module handle creation in
 * tinyfusers/vision/conv2d.py:7 and tinyfusers/ff/layer_norm.py:6 (cudnn.create_handle()). */
int tf_init(int device);
const char* tf_last_error(void);
/* number of kernels this library has launched since the last reset (bench.py: gpu_launches) */
/* programmatic dependent launch between this library's kernels (default on; env TINYFUSERS_B200_PDL=0 disables) */
int tf_set_pdl(int enable);
long long tf_launch_count(void);
void tf_launch_count_reset(void);

/* ---- CUDA-graph capture without a host framework ------------------------------------------------------------------
 * Every tf_* call is graph-capturable. The package captures with torch.cuda.CUDAGraph; a host that binds the library from
 * CuPy / ctypes (the reference's world: tinyfusers/native/cudart/ops.py binds cudart the same way) captures and replays the
 * step through these: begin on a non-default stream, enqueue tf_* calls on it, end -> an instantiated graph, launch it any
 * number of times. Replaces nothing in the reference (it launches every operator eagerly: example/sd1.py:68-73). */
int tf_graph_begin_capture(void* stream);
int tf_graph_end_capture(void* stream, void** graph_exec_out);
int tf_graph_launch(void* graph_exec, void* stream);
int tf_graph_destroy(void* graph_exec);

/* ---- epilogue flags for tf_gemm_f16 / tf_conv2d_nhwc_f16 ----------------------------------------- */
#define TF_EPI_NONE 0
#define TF_EPI_OUT_F32 1 /* write fp32 instead of fp16 */
#define TF_EPI_GEGLU 2   /* out[m, j] = (acc[m, v_j] + b) * gelu_tanh(acc[m, g_j] + b); weight rows packed
                            by tf_pack_geglu_rows: every 32 rows = 16 value rows then 16 gate rows */
#define TF_GEMM_W_STATIC 4 /* promise: the W operand is a static weight, complete in memory before this call is enqueued
                            (not written by a kernel still running on the stream). The kernel then starts fetching its
                            first W tiles BEFORE it waits for the preceding kernel (programmatic dependent launch):
                            every layer's weights arrive cold from HBM. Never set it for swapped-operand GEMMs that put
                            an activation in the W slot (V^T = Wv . X^T). */

/* D[M,N] = A[M,K] · W[N,K]^T (+ bias[N]) (+ residual[M,N]); fp16 in, fp32 accumulate (tcgen05/TMEM).
 * Replaces: Linear.__call__  cp.dot(x, W.T) + b          tinyfusers/ff/linear.py:116-121
 *           1x1 Conv2d (cuDNN conv_fprop + bias add)      tinyfusers/vision/conv2d.py:9-28,55-59
 *           GEGLU (with TF_EPI_GEGLU)                      tinyfusers/ff/nn.py:5-12
 * A: row-major, leading dim lda (elements); W: row-major (out_features, in_features), leading dim ldw.
 * K, lda, ldw, N, ldc, ldr multiples of 8; all pointers 16-byte aligned. bias is fp32 (may be NULL).
 * workspace (may be NULL): fp32 scratch for split-K partials; enables split-K for small-M / deep-K. */
int tf_gemm_f16(const void* A, int lda, const void* W, int ldw, void* out, int ldc, int M, int N, int K,
                const float* bias, const void* residual, int ldr, int flags, void* workspace,
                size_t ws_bytes, void* stream);

/* 2-D cross-correlation, NHWC fp16, as an implicit GEMM on tcgen05 (3x3 pad 1 stride 1|2, or 1x1).
 * Replaces: conv_2d / Conv2d.__call__ (cuDNN conv_fprop graph, NHWC->NCHW transpose, bias add)
 *           tinyfusers/vision/conv2d.py:9-28,48-59; Downsample tinyfusers/vision/unet.py:86-90.
 * x: (NI, H, W, Cin) with pixel stride x_pixel_stride >= Cin (elements); w: (Cout, kh, kw, Cin) fp16,
 * i.e. the reference's OIHW weight permuted to OHWI; out: (NI, Ho, Wo, *) with pixel stride ldc;
 * residual (optional) has the geometry of out with pixel stride ldr. Cin % 64 == 0 for 3x3. */
int tf_conv2d_nhwc_f16(const void* x, int NI, int H, int W, int Cin, int x_pixel_stride, const void* w,
                       int Cout, int ksize, int stride, void* out, int ldc, const float* bias,
                       const void* residual, int ldr, int flags, void* workspace, size_t ws_bytes,
                       void* stream);

/* Same as tf_gemm_f16 / tf_conv2d_nhwc_f16, and additionally leaves the GroupNorm statistics of the fp16 OUTPUT in
 * gn_stats[image][slot][unit] = {sum, sumsq} (float2; slot = 32 consecutive rows of one image, unit = gn_unit
 * consecutive channels), computed in the epilogue from the rounded values, so that the GroupNorm that consumes
 * this tensor (tf_groupnorm_fused_nhwc_f16) is ONE launch instead of statistics + finalize + apply.
 * Replaces the same reference calls as the plain versions; the statistics replace the mean / var passes of
 * group_norm (tinyfusers/ff/group_norm.py:5-10). gn_stats: NI * (rows_per_image/32) * (N/gn_unit) float2.
 * Needs rows_per_image % 32 == 0, N % gn_unit == 0, lcm(32, gn_unit) <= 256 and the plain fp16 epilogue;
 * tf_gn_stats_supported() answers whether a geometry qualifies (TF_ERR_UNSUPPORTED otherwise). */
int tf_gemm_gn_f16(const void* A, int lda, const void* W, int ldw, void* out, int ldc, int M, int N, int K,
                   const float* bias, const void* residual, int ldr, int flags, void* workspace, size_t ws_bytes,
                   void* gn_stats, int gn_unit, int rows_per_image, void* stream);
int tf_conv2d_nhwc_gn_f16(const void* x, int NI, int H, int W, int Cin, int x_pixel_stride, const void* w, int Cout,
                          int ksize, int stride, void* out, int ldc, const float* bias, const void* residual, int ldr,
                          int flags, void* workspace, size_t ws_bytes, void* gn_stats, int gn_unit, void* stream);
int tf_gn_stats_supported(int NI, int Ho, int Wo, int C, int gn_unit, int is_conv3x3);

/* tf_gemm_f16 with optional fused extras (NULL pointers = off):
 *   gn_stats / gn_unit / gn_rows_per_image : as tf_gemm_gn_f16;
 *   row_stats_out : float2 buffer of >= M * (N/32) entries; receives {sum, sumsq} of every output row per N-tile of this
 *                   launch (the library remembers the per-row entry count for the consumer; N % 32 == 0, plain fp16
 *                   epilogue, no split-K) - the statistics a LayerNorm over the row needs;
 *   ln_stats / ln_chunks / ln_c1 / ln_eps : LayerNorm FOLDED onto the A operand. A holds the un-normalised rows (K = 32 *
 *                   ln_chunks columns), ln_stats is the row_stats_out of the GEMM that produced A, W must have been
 *                   pre-multiplied by gamma (W' = W diag(gamma)), ln_c1[n] = sum_k W'[n,k], and `bias` must carry
 *                   W beta (+ the layer's own bias). The epilogue computes rstd[m] * (acc - mean[m] * c1[n]) + bias[n]:
 *                   exactly Linear(LayerNorm(A)) without the LayerNorm launch and its read + write pass.
 * Replaces: LayerNorm.__call__ followed by Linear / GEGLU (tinyfusers/attention/attention.py:50-55, ff/layer_norm.py:8-49). */
typedef struct tf_gemm_extras {
  void* gn_stats;
  int gn_unit;
  int gn_rows_per_image;
  void* row_stats_out;
  const void* ln_stats;
  int ln_chunks;
  const float* ln_c1;
  float ln_eps;
} tf_gemm_extras;
int tf_gemm_ex_f16(const void* A, int lda, const void* W, int ldw, void* out, int ldc, int M, int N, int K,
                   const float* bias, const void* residual, int ldr, int flags, void* workspace, size_t ws_bytes,
                   const tf_gemm_extras* ex, void* stream);
/* out = conv3x3(x, w[:, :9*Cin]) + conv1x1(x2, w[:, 9*Cin:]) + bias in ONE launch (stride 1, pad 1): the 1x1 convolution
 * of the second source is appended to the implicit GEMM along K (C2/64 extra k-blocks read through a second tensor map).
 * w: (Cout, 9*Cin + C2) fp16 - OHWI rows of the 3x3 weight followed by the 1x1 weight's row. gn_stats may be NULL.
 * Replaces: ResBlock `skip_connection(x) + h` with a Conv2d skip (tinyfusers/vision/resnet.py:22,29-30) and ResnetBlock
 *           `nin_shortcut(x) + h` (resnet.py:39,43-44): a separate cuDNN conv + an elementwise add. */
int tf_conv2d_nhwc_skip_f16(const void* x, int NI, int H, int W, int Cin, int x_pixel_stride, const void* x2, int C2,
                            int x2_pixel_stride, const void* w, int Cout, void* out, int ldc, const float* bias, int flags,
                            void* workspace, size_t ws_bytes, void* gn_stats, int gn_unit, void* stream);

/* test / tuning hook: force the N tile and split-K factor of the next GEMM/conv calls (0 = auto) */
int tf_gemm_set_tuning(int force_bn, int force_splits);
/* test / tuning hook: 0 = auto, 1 = single-CTA tiles only, 2 = CTA-pair (cta_group::2, 256-row) tiles where M > 128 */
int tf_gemm_set_ctas(int force_ctas);
/* Split-K folded inside a thread-block cluster through distributed shared memory (1) or through an fp32 workspace + fold
   kernel (0, the default: measured faster on every shape of the step); < 0 restores the default (environment
   TINYFUSERS_B200_CLUSTER_SPLITK). */
int tf_gemm_set_cluster_splitk(int on);
/* measured tile choices: (is_conv, M, N, K, class) -> (BN, split-K factor, CTAs per tile) overrides the built-in cost
 * model for that shape. class bits: 1 GEGLU, 2 fp32 out, 4 GroupNorm statistics, 8 residual, 16 stride-2 conv. For a
 * conv, M = NI*Ho*Wo, N = Cout, K = 9*Cin. Entries that do not fit a call (workspace too small, ...) are ignored.
 * tools/autotune_gemm.py writes native/b200/gemm_tuning.json on a B200; the Python binding loads it in init(). */
int tf_gemm_tuning_add(int is_conv, int M, int N, int K, int klass, int bn, int splits, int ctas);
int tf_gemm_tuning_clear(void);
/* the (BN, split-K, CTAs) the most recent GEMM / conv call on this process used */
int tf_gemm_last_choice(int* bn, int* splits, int* ctas);
/* measurement hook: cap the shared-memory ring depth of the next GEMM/conv launches (0 = as many stages as fit) */
int tf_gemm_set_max_stages(int max_stages);
/* debug hook: per-CTA clock64 stamps of the next GEMM/conv launches ([grid][8] int64; NULL = off) */
int tf_gemm_set_timeline(long long* dev_buf);
/* Next-layer weight prefetch into L2. Every layer's weights arrive cold from HBM (SD 1.5: 1.7 GB of fp16 weights per step
 * through a 126 MB L2), and a launch cannot know which layer follows it - but a denoising step is the same launch sequence
 * every time. mode 1: record the (pointer, bytes) of every static-weight GEMM / conv launch from now on (run ONE eager step);
 * mode 2: replay - launch i of the same sequence (in practice: the capture of the step's CUDA graph) asks L2 for the weights
 * of launch i + 1 once its own operand loads are in flight (cp.async.bulk.prefetch.L2, one slice per CTA), the last launch for
 * those of launch 0 (the next step); mode 0: off. A launch whose weights differ from the recorded ones ends the replay.
 * No reference counterpart (the reference's cuDNN / cuBLAS calls stream weights per call: tinyfusers/vision/conv2d.py:31-46,
 * tinyfusers/ff/linear.py:119-120). Results are unaffected: a prefetch only warms the cache. */
int tf_weight_prefetch_mode(int mode);
/* weights smaller than min_bytes or larger than max_bytes are not prefetched (defaults 1 MiB / 96 MiB) */
int tf_weight_prefetch_limits(long long min_bytes, long long max_bytes);
/* launches recorded by the last mode-1 pass; launches (and bytes) that were handed a hint since the last mode change */
int tf_weight_prefetch_stats(int* recorded, int* hinted, long long* hinted_bytes);

/* Nearest-neighbour 2x upsampling followed by a 3x3 convolution (pad 1), in ONE launch: out = conv3x3(upsample2x(x)).
 * Output pixel (2y+a, 2x+b) sees only the 2 x 2 input pixels (y+a-1.., x+b-1..), so each output phase (a, b) is a 2 x 2
 * convolution of the ORIGINAL image with the 3x3 taps that fall on the same input pixel summed beforehand: 4/9 of the
 * multiply-adds of the reference's conv over the 4x tensor, and the 4x tensor is never written.
 * x: (NI, H, W, Cin) NHWC fp16, Cin % 64 == 0. w4: (4, wrows, 4 Cin) fp16 - phase p = 2a + b, row = output channel,
 * k = (2i + j) Cin + c = the summed taps for input pixel (y + a - 1 + i, x + b - 1 + j) (tinyfusers_b200/packing.py
 * conv_up2x_weight builds it from the OIHW weight, summing in fp32). out: (NI, 2H, 2W, ldc >= Cout) NHWC fp16. bias fp32 or
 * NULL. gn_stats / gn_unit as in tf_conv2d_nhwc_gn_f16 (NULL: none; needs H * W % 32 == 0).
 * Replaces: Upsample.__call__ = reshape/expand 2x + Conv2d  tinyfusers/vision/unet.py:78-84. */
int tf_conv2d_up2x_nhwc_f16(const void* x, int NI, int H, int W, int Cin, int x_pixel_stride, const void* w4, int wrows,
                            int Cout, void* out, int ldc, const float* bias, int flags, void* gn_stats, int gn_unit,
                            void* stream);

/* ---- normalisation -------------------------------------------------------------------------------- */
/* GroupNorm (+ optional SiLU), NHWC fp16 -> NHWC fp16, statistics fp32, biased variance.
 * Replaces: group_norm / GroupNorm.__call__ (>= 8 CuPy launches)   tinyfusers/ff/group_norm.py:3-21
 *           + Tensor.silu that always follows it in the UNet         tinyfusers/storage/tensor.py:68-70,
 *                                                                    vision/resnet.py:8-10,17-19, vision/unet.py:45-47
 * The input may be given as TWO channel slices (x: Cx channels, x2: Cx2 channels, each with its own pixel
 * stride) that are normalised as their channel concatenation — this is how
 * `cp.concatenate((x, saved_inputs.pop()), axis=1)` (vision/unet.py:72) is consumed without a copy.
 * x2 may be NULL. gamma/beta: fp32 (C) or NULL. stats_ws: scratch of tf_groupnorm_workspace_bytes(NI, groups)
 * bytes. Statistics are reduced in a fixed order (no atomics): results are bit-reproducible. */
size_t tf_groupnorm_workspace_bytes(int NI, int groups);
int tf_groupnorm_nhwc_f16(const void* x, int x_pixel_stride, int Cx, const void* x2, int x2_pixel_stride,
                          int Cx2, void* out, int out_pixel_stride, int NI, int HW, int groups,
                          const float* gamma, const float* beta, float eps, int apply_silu, float* stats_ws,
                          void* stream);

/* GroupNorm (+ optional SiLU) in ONE launch for inputs whose producer left statistics (tf_gemm_gn_f16 /
 * tf_conv2d_nhwc_gn_f16): folds the slots, finalises mean / rstd per (image, group) and normalises. Same
 * two-source concatenation semantics as tf_groupnorm_nhwc_f16; each source has its own unit size, and unit
 * boundaries must coincide with group and source boundaries ((C/groups) % unit == 0, Cx % x2_unit == 0).
 * Replaces: group_norm / GroupNorm.__call__ + Tensor.silu (tinyfusers/ff/group_norm.py:3-21). */
int tf_groupnorm_fused_nhwc_f16(const void* x, int x_pixel_stride, int Cx, const void* x_stats, int x_unit,
                                const void* x2, int x2_pixel_stride, int Cx2, const void* x2_stats, int x2_unit,
                                void* out, int out_pixel_stride, int NI, int HW, int groups, const float* gamma,
                                const float* beta, float eps, int apply_silu, void* stream);

/* LayerNorm over the last dim of a (rows, C) fp16 matrix; gamma/beta fp32.
 * Replaces: layer_norm / LayerNorm.__call__ (cuDNN layernorm graph, rebuilt per call)
 *           tinyfusers/ff/layer_norm.py:8-49.
 * interleave == 1: canonical. interleave == B > 1: the reference's stride declaration at batch B
 * (layer_norm.py:10): the buffer is viewed as (rows/B, C, B) and normalised over C. */
int tf_layernorm_f16(const void* x, void* out, int rows, int C, const float* gamma, const float* beta, float eps,
                     int interleave, void* stream);

/* ---- attention ------------------------------------------------------------------------------------ */
/* softmax(scale * Q K^T) V per (batch, head), flash-style on tcgen05/TMEM; scores never reach HBM.
 * Replaces: scaled_dot_product_attention (2 cuBLAS batched SGEMMs + softmax_kernel)
 *           tinyfusers/attention/sdpa.py:53-77, tinyfusers/native/cuda/softmax.cu:24-112.
 * q: (B*Tq, ldq), k: (B*Tk_pad, ldk): head h at columns [h*dp, (h+1)*dp), dp = d padded to 16 with zeros;
 * vt: (NH*dp, ldvt >= B*Tk_pad) = V transposed, ldvt % 8 == 0; batch b owns key rows/columns
 * [b*Tk_pad, b*Tk_pad + Tk); keys [Tk, Tk_pad) of each batch are padding (ignored).
 * out element (b,h,t,j<d) at b*out_stride_b + h*out_stride_h + t*out_stride_t + j — the strides select
 * the reference's head-major reshape (attention/attention.py:39) or the canonical head merge. */
int tf_attention_f16(const void* q, int ldq, const void* k, int ldk, const void* vt, int ldvt, void* out,
                     long long out_stride_b, long long out_stride_h, long long out_stride_t, int B, int NH,
                     int Tq, int Tk, int Tk_pad, int d, int dp, float scale, void* stream);
/* Same attention with V in its NATURAL layout: v is (B*Tk_pad, ldv), head h at columns [h*dvp, (h+1)*dvp), dvp = head
 * dim padded with zero columns to a multiple of 16 (tiles use 32-byte swizzle atoms) or of 64 (128-byte atoms) - what a
 * fused [Q | K | V] projection GEMM writes directly, so no transposed copy of V is produced by anyone (V tiles are the
 * MN-major B operand of O += P V). causal = flags: bit 0 (TF_ATTN_CAUSAL) query t only sees keys <= t; bit 1
 * (TF_ATTN_V_ONES_COLUMN) column d of every V head holds 1.0 in every key row (dvp > d; the projection GEMM's bias writes
 * it), so the softmax denominator is column d of P V and needs no reduction of its own. */
#define TF_ATTN_CAUSAL 1
#define TF_ATTN_V_ONES_COLUMN 2
int tf_attention_v_f16(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out,
                       long long out_stride_b, long long out_stride_h, long long out_stride_t, int B, int NH, int Tq,
                       int Tk, int Tk_pad, int d, int dp, int dvp, float scale, int causal, void* stream);
int tf_attention_set_tuning(int force_bn);
/* Kernel variant for tf_attention_v_f16: version 0 = automatic (the split-row kernel where its grid is one wave, the
   two-query-tile ping-pong kernel where whole pairs of 128-row query tiles fill the GPU, the one-tile kernel elsewhere),
   1 = one-tile kernel only, 2 = ping-pong wherever it applies, 3 = split-row wherever it applies (head dim <= 48, V padded
   to 64 columns); emu = exponentials per 8 evaluated on the FMA pipe instead of MUFU.EX2 (0, 2, 4; < 0: each kernel's
   measured best). */
int tf_attention_set_variant(int version, int emu);
/* debug hook: per-block clock64 stamps of one softmax warp ([64][8] int64; NULL = off; TF_ATT_TRACE builds only) */
int tf_attention_set_timeline(long long* dev_buf);
/* debug: pinned host buffer (>= 64 bytes, zeroed) in which a timed-out barrier wait of the attention kernels records
   {1, code, block x, y, z, warp, key block, parity} before it traps; NULL switches it off */
int tf_attention_set_debug(void* host_mapped_buf);
/* Same, under a causal mask: query t attends to keys <= t only (Tq == Tk). Replaces CLIPAttention's
 * scaled_dot_product_attention(q, k, v, attn_mask = triu(full(-inf), k=1))
 *           tinyfusers/attention/attention.py:88-99, tinyfusers/vae/encoder.py:79, attention/sdpa.py:67-68. */
int tf_attention_causal_f16(const void* q, int ldq, const void* k, int ldk, const void* vt, int ldvt, void* out,
                            long long out_stride_b, long long out_stride_h, long long out_stride_t, int B, int NH,
                            int Tq, int Tk, int Tk_pad, int d, int dp, float scale, void* stream);

/* ---- small ops of the step --------------------------------------------------------------------------- */
/* [cos(t f_i) | sin(t f_i)], f_i = exp(-ln(max_period) i / (dim/2)); t = timesteps_dev[*index_dev]
 * (index_dev may be NULL -> element 0). Angles in fp64 like the reference. out: fp32 (dim).
 * Replaces: timestep_embedding  tinyfusers/vision/unet.py:92-97. */
int tf_timestep_embedding_f32(const float* timesteps_dev, const int* index_dev, int dim, float max_period,
                              float* out, void* stream);
/* out[n] = sum_k act(x[k]) W[n,k] + bias[n] + bias2[n]; x fp32 (K), W fp16 (N,K), act = SiLU if silu_input.
 * Replaces: Linear at M = 1 — time_embed (vision/unet.py:11,54) and ResBlock.emb_layers
 *           (vision/resnet.py:13-16,27-28). */
int tf_gemv_f16w(const float* x, const void* W, const float* bias, const float* bias2, float* out, int N, int K,
                 int silu_input, void* stream);
/* 3x3 pad-1 conv, Cin = 4, fp32 NCHW in, fp32 OIHW weights, fp16 NHWC out.
 * Output image n reads input image n % x_images: the CFG batch [latent ; latent] (variants/sd.py:31,
 * cp.broadcast_to) is formed without materialising the copy.
 * Replaces: UNetModel.input_blocks[0] Conv2d(4,320)  tinyfusers/vision/unet.py:13. */
int tf_conv3x3_smallcin_f32nchw(const float* x, int x_images, const float* w, const float* bias, void* out, int NI,
                                int Cin, int H, int W, int Cout, int out_pixel_stride, void* stream);
/* Same, and additionally leaves the GroupNorm statistics of the fp16 output (layout of tf_conv2d_nhwc_gn_f16: slot = 32
 * consecutive pixels of one image, unit = gn_unit channels), so the first ResBlock's GroupNorm is one launch. H*W % 32 == 0. */
int tf_conv3x3_smallcin_gn_f32nchw(const float* x, int x_images, const float* w, const float* bias, void* out, int NI,
                                   int Cin, int H, int W, int Cout, int out_pixel_stride, void* gn_stats, int gn_unit,
                                   void* stream);
/* nearest x2. Replaces: Upsample.__call__ broadcast/reshape  tinyfusers/vision/unet.py:81-83. */
int tf_upsample_nearest2x_nhwc_f16(const void* x, int x_pixel_stride, void* out, int out_pixel_stride, int NI,
                                   int H, int W, int C, void* stream);
/* layout edges of the drop-in API (the reference is NCHW fp32 everywhere, vision/conv2d.py:27) */
int tf_nchw_to_nhwc_f16(const void* x, int x_is_f32, void* out, int NI, int C, int HW, int out_pixel_stride,
                        void* stream);
int tf_nhwc_to_nchw(const void* x, int x_pixel_stride, void* out, int out_is_f32, int NI, int C, int HW,
                    void* stream);
/* prompt context (B,T,C) fp32 -> (B,Tpad,C) fp16, zero rows appended (77 -> 80 tokens) */
int tf_pad_tokens_f32_to_f16(const float* x, void* out, int B, int T, int Tpad, int C, void* stream);
/* e_t = u + g (c - u); x_prev = sqrt(a_prev) (x - sqrt(1-a_t) e_t)/sqrt(a_t) + sqrt(1-a_prev) e_t.
 * eps: fp32 NHWC (2B images: [uncond ; cond]), latent: fp32 NCHW; a_t = alphas_dev[*index_dev] etc.
 * e_t_out may be NULL. Replaces: StableDiffusion.get_model_output :44-45 and get_x_prev_and_pred_x0
 * (tinyfusers/variants/sd.py:14-25,44-45) — ~12 CuPy launches fused into one. */
int tf_cfg_ddim_step_f32(const float* eps_nhwc, int eps_pixel_stride, const float* latent, float* latent_out,
                         float* e_t_out, const float* alphas_dev, const float* alphas_prev_dev,
                         const int* index_dev, float guidance, int B, int C, int HW, void* stream);
/* ---- CFG halves on two GPUs (one image; lower latency): GPU 0 evaluates the unconditional half, GPU 1 the conditional half of
 * get_model_output's batch (tinyfusers/variants/sd.py:27-46) and ONE kernel per GPU exchanges the two noise predictions through
 * peer-mapped memory over NVLink and applies CFG + DDIM (no NCCL launch on the step path; csrc/tf_p2p.cu).
 *   tf_p2p_alloc / tf_p2p_free : exchange buffer from cudaMalloc + its 64-byte CUDA IPC handle (zero-filled);
 *   tf_p2p_open / tf_p2p_close : map the PEER process's buffer from its handle (peer access enabled lazily);
 *   tf_p2p_blocks              : thread blocks (= flags per slot) the step kernel uses for C*HW elements.
 * Mailbox = 2 * C*HW floats, flags = 2 * tf_p2p_blocks ints, both per rank and zero-initialised. seq_dev: device counter the
 * caller increments after every step (tf_add_int). mode: 1 send | 2 wait + update (3 in production). rank: 0 = holds the
 * unconditional half. eps_nhwc: this rank's UNet output, ONE image, fp32 NHWC. */
int tf_p2p_blocks(int C, int HW);
int tf_p2p_alloc(size_t bytes, void** dev_ptr, void* handle64);
int tf_p2p_open(const void* handle64, void** dev_ptr);
int tf_p2p_close(void* dev_ptr);
int tf_p2p_free(void* dev_ptr);
int tf_cfg_ddim_step_split_f32(const float* eps_nhwc, int eps_pixel_stride, const float* latent, float* latent_out,
                               float* e_t_out, const float* alphas_dev, const float* alphas_prev_dev, const int* index_dev,
                               float guidance, int C, int HW, int rank, const void* my_mailbox, void* peer_mailbox,
                               const void* my_flags, void* peer_flags, const int* seq_dev, int mode, void* stream);
/* *p_dev += delta (device-resident sampler step counter, so a captured graph can be replayed) */
int tf_add_int(int* p_dev, int delta, void* stream);
/* elementwise activation, op: 0 sigmoid, 1 silu/swish, 2 gelu (tanh approx), 3 quick_gelu; fp32 or fp16.
 * Replaces: Tensor.sigmoid/silu/gelu/quick_gelu static methods  tinyfusers/storage/tensor.py:64-86. */
int tf_unary(const void* x, void* out, long long n, int op, int is_f32, void* stream);
/* fp32 NHWC (pixel stride >= C) -> fp32 NCHW: returns the UNet's eps in the reference layout */
int tf_nhwc_f32_to_nchw_f32(const float* x, int x_pixel_stride, float* out, int NI, int C, int HW, void* stream);
/* DDIM eta=0 update on its own; a_t / a_prev are device scalars; pred_x0 may be NULL.
 * Replaces: StableDiffusion.get_x_prev_and_pred_x0  tinyfusers/variants/sd.py:14-25. */
int tf_ddim_step_f32(const float* x, const float* e_t, const float* a_t_dev, const float* a_prev_dev, float* x_prev,
                     float* pred_x0, long long n, void* stream);


/* ---- rows next to the hot path: VAE decoder, CLIP text encoder --------------------- */
/* Attention of the reference's AttnBlock as the reference executes it: 4-D (B,C,H,W) q/k/v handed to
 * scaled_dot_product_attention are read as (B, NH = C, T = H, HS = W), i.e. per channel plane
 * softmax(scale * Q K^T) V with Q, K, V the H x W planes.  q, k, v, out: (planes, H, W) fp16 contiguous (NCHW).
 * Replaces: AttnBlock.__call__ -> scaled_dot_product_attention   tinyfusers/attention/attention.py:19-24,
 *           tinyfusers/attention/sdpa.py:53-77. */
int tf_plane_attention_f16(const void* q, const void* k, const void* v, void* out, int planes, int H, int W, float scale,
                           void* stream);
/* probs[r, :] = softmax(scale * scores[r, :]), fp32 (rows, lds) -> fp16 (rows, ldp): the softmax between the QK^T and PV
 * GEMMs (tf_gemm_f16) of the CANONICAL AttnBlock - one head over the H*W pixels, head dim = channels (512), which the
 * fused attention kernel (head dim <= 256) does not cover.
 * Replaces: softmax_kernel  tinyfusers/native/cuda/softmax.cu:24-112 inside  tinyfusers/attention/sdpa.py:53-77. */
int tf_softmax_rows_f32_to_f16(const float* scores, long long lds, void* probs, long long ldp, long long rows, int cols,
                               float scale, void* stream);
/* out[n,co,p] = bias[co] + sum_ci w[co,ci] * (scale * x[n,ci,p]); fp32 NCHW, Cin, Cout <= 8.
 * Replaces: post_quant_conv(1/0.18215 * x)   tinyfusers/variants/sd.py:49, tinyfusers/vae/vae.py:10. */
int tf_conv1x1_small_f32nchw(const float* x, const float* w, const float* bias, float* out, int NI, int Cin, int Cout,
                             int HW, float scale, void* stream);
/* clip((x + 1) / 2, 0, 1) * 255 -> uint8 (truncating cast); fp32 NHWC (pixel stride >= C) -> uint8 HWC.
 * Replaces: StableDiffusion.decode post-processing   tinyfusers/variants/sd.py:51-53. */
int tf_image_to_u8(const float* x_nhwc, int x_pixel_stride, void* out_u8, long long pixels, int C, void* stream);
/* out[r, :] = table[ids[r], :] + pos_table[r % T, :] (pos_table may be NULL); fp32 tables -> fp16 rows.
 * Replaces: Embedding.__call__ (one-hot GEMM)   tinyfusers/ff/embedding.py:15-23, CLIPTextEmbeddings
 *           tinyfusers/vae/encoder.py:66-70. */
int tf_embedding_f16(const int* ids, const float* table, const float* pos_table, void* out, int rows, int T, int E,
                     int vocab, void* stream);
int tf_cast_f16_to_f32(const void* x, float* out, long long n, void* stream);

/* ---- fp32 parity mode (BASELINE.json configs[0]; north star: per-op error <= 1e-5 in fp32 mode) -------------
 * The reference computes in fp32 throughout; these are the same operators as plain-fp32 CUDA-core kernels in the
 * reference's own layouts (NCHW / OIHW / (B,T,C)), selected with tinyfusers_b200.set_precision("fp32").
 * A correctness mode: never on the measured path (csrc/tf_fp32.cu). */
/* out[z][m][n] = alpha * sum_k A[z][m][k] * W[z][n][k] (+bias[n]) (+residual[z][m][n]).  w_kn != 0: W is [K][N]
 * (the P.V product).  out_nchw_hw > 0: M = images * hw rows are written (and the residual read) as NCHW (images, N, hw)
 * - the 1x1 proj_out of SpatialTransformer reading tokens.  Batches: z = zo * batch_inner + zi with element strides
 * batch_strides = {A_outer, A_inner, W_outer, W_inner, out_outer, out_inner} (host array; heads inside (B,T,C) tokens).
 * Replaces: cp.dot(x, W.T) + b  tinyfusers/ff/linear.py:119-120; cp.matmul  tinyfusers/attention/sdpa.py:66,76. */
int tf_gemm_f32(const float* A, long long lda, const float* W, long long ldw, int w_kn, const float* bias,
                const float* residual, long long ldr, float* out, long long ldc, int M, int N, int K, float alpha,
                int out_nchw_hw, int batch_outer, int batch_inner, const long long* batch_strides, void* stream);
/* NCHW x OIHW cross-correlation, square stride / padding, no dilation; + bias[o] + bias_img[image][o] + residual.
 * out_tokens != 0 writes (images * Ho * Wo, O) rows instead of NCHW (proj_in feeding the transformer block).
 * Replaces: conv_2d + bias add  tinyfusers/vision/conv2d.py:9-28,55-59; `h + emb_out`  tinyfusers/vision/resnet.py:27. */
int tf_conv2d_nchw_f32(const float* x, const float* w, const float* bias, const float* bias_img, const float* residual,
                       float* out, int NI, int C, int H, int W, int O, int R, int S, int stride, int pad, int out_tokens,
                       void* stream);
/* literal two-pass GroupNorm on NCHW: (x - mean) * (1 / sqrt(mean((x - mean)^2) + eps)) [* gamma + beta] [-> SiLU].
 * Replaces: group_norm / GroupNorm.__call__  tinyfusers/ff/group_norm.py:3-21. */
int tf_groupnorm_nchw_f32(const float* x, const float* gamma, const float* beta, float* out, int NI, int C, int HW,
                          int groups, float eps, int silu, void* stream);
/* LayerNorm over the last dimension of (rows, C).  Replaces: layer_norm  tinyfusers/ff/layer_norm.py:8-32. */
int tf_layernorm_f32(const float* x, const float* gamma, const float* beta, float* out, long long rows, int C, float eps,
                     void* stream);
/* in-place max-subtracted row softmax; causal_tq > 0: row r is query r % causal_tq and sees keys 0..query only (the
 * triu(-inf, k=1) mask of tinyfusers/vae/encoder.py:79).  Replaces: softmax_kernel  tinyfusers/native/cuda/softmax.cu:24-112. */
int tf_softmax_rows_f32(float* x, long long rows, int cols, int causal_tq, void* stream);
/* out[r, :] = table[ids[r], :] + pos_table[r % T, :] in fp32 (pos_table may be NULL).
 * Replaces: Embedding.__call__  tinyfusers/ff/embedding.py:15-23, CLIPTextEmbeddings  tinyfusers/vae/encoder.py:66-70. */
int tf_embedding_f32(const int* ids, const float* table, const float* pos_table, float* out, int rows, int T, int E,
                     int vocab, void* stream);
/* op: 0 sigmoid, 1 silu/swish, 2 gelu (tanh approx), 3 quick_gelu with IEEE expf / tanhf.
 * Replaces: Tensor.sigmoid/silu/gelu/quick_gelu  tinyfusers/storage/tensor.py:64-86. */
int tf_unary_f32(const float* x, float* out, long long n, int op, void* stream);
/* out[m][j] = y[m][j] * gelu(y[m][H + j]).  Replaces: GEGLU.__call__  tinyfusers/ff/nn.py:10-12. */
int tf_geglu_f32(const float* y, long long ldy, float* out, long long M, int H, void* stream);
/* out = uncond + guidance * (cond - uncond).  Replaces: the CFG combine of get_model_output  tinyfusers/variants/sd.py:44-45. */
int tf_cfg_combine_f32(const float* uncond, const float* cond, float guidance, float* out, long long n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TINYFUSERS_B200_H_ */
