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
long long tf_launch_count(void);
void tf_launch_count_reset(void);

/* ---- epilogue flags for tf_gemm_f16 / tf_conv2d_nhwc_f16 ----------------------------------------- */
#define TF_EPI_NONE 0
#define TF_EPI_OUT_F32 1 /* write fp32 instead of fp16 */
#define TF_EPI_GEGLU 2   /* out[m, j] = (acc[m, v_j] + b) * gelu_tanh(acc[m, g_j] + b); weight rows packed
                            by tf_pack_geglu_rows: every 32 rows = 16 value rows then 16 gate rows */

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

/* test / tuning hook: force the N tile and split-K factor of the next GEMM/conv calls (0 = auto) */
int tf_gemm_set_tuning(int force_bn, int force_splits);

#ifdef __cplusplus
}
#endif
#endif /* TINYFUSERS_B200_H_ */
