"""Host-side runtime shared by the drop-in module classes.

PyTorch is used here strictly as a container: `torch.empty` for device memory, `data_ptr()` for raw
pointers, `torch.cuda.current_stream()` for the stream, `torch.cuda.CUDAGraph` for capture. All
arithmetic happens in libtinyfusers_b200.so (tinyfusers_b200/native/b200/ops.py).

Activations on the fast path are fp16 NHWC views (`Act`) carved from a stack arena whose addresses are
a deterministic function of the call sequence — so a whole UNet step can be captured into a CUDA graph
and replayed.
"""
import math
import os

import torch

import ctypes

from .native.b200.ops import GemmExtras, b200

F16 = torch.float16
F32 = torch.float32


def _align(n, a=256):
    return (n + a - 1) // a * a


def gn_unit(c, groups=32):
    """Statistics unit (channels) for a c-channel tensor: divides c/groups AND the group sizes of the UNet's
    channel concatenations that contain it (320/640/1280 -> 10), so one set of statistics serves every consumer."""
    base = c // groups if c % groups == 0 else 0
    if base <= 0:
        return 0
    g = math.gcd(base, 10)
    return g if g > 1 else base


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def require_cuda(t, name="tensor"):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA torch tensor (tinyfusers_b200 has no CPU path)")


class Act:
    """A (n, h, w, c) fp16 NHWC activation view: pixel p of image i starts at ptr + 2*((i*h*w + p)*stride)."""

    __slots__ = ("ptr", "n", "h", "w", "c", "stride", "keep", "valid", "gn")

    def __init__(self, ptr, n, h, w, c, stride=None, keep=None):
        self.ptr, self.n, self.h, self.w, self.c = ptr, n, h, w, c
        self.stride = c if stride is None else stride
        self.keep = keep  # torch tensor owning the memory when not arena-backed
        self.valid = None  # for padded token sequences: number of real tokens per batch (<= h)
        # GroupNorm statistics left by the producer(s): list of (c0, c1, stats_ptr, unit) channel parts, or None
        self.gn = None

    @property
    def rows(self):
        return self.n * self.h * self.w

    def channels(self, c0, c1):
        return Act(self.ptr + 2 * c0, self.n, self.h, self.w, c1 - c0, self.stride, self.keep)

    def as_tokens(self):
        return Act(self.ptr, self.n, self.h * self.w, 1, self.c, self.stride, self.keep)

    def reshaped(self, n, h, w):
        assert n * h * w == self.rows
        return Act(self.ptr, n, h, w, self.c, self.stride, self.keep)


class Arena:
    """Stack allocator over one device buffer. `dry` mode only measures the peak."""

    def __init__(self):
        self.buf = None
        self.base = 0
        self.top = 0
        self.peak = 0
        self.capacity = 0
        self.dry = False

    def reserve(self, nbytes, device):
        if self.buf is None or self.capacity < nbytes:
            self.buf = torch.empty(_align(nbytes, 1 << 20), dtype=torch.uint8, device=device)
            self.capacity = self.buf.numel()
            self.base = self.buf.data_ptr()
            assert self.base % 1024 == 0

    def reset(self):
        self.top = 0

    def mark(self):
        return self.top

    def release(self, mark):
        self.top = mark

    def alloc(self, nbytes):
        off = self.top
        self.top = off + _align(nbytes, 1024)
        self.peak = max(self.peak, self.top)
        if self.dry:
            return 1024 + off  # fake but well-aligned address; nothing is launched in dry mode
        if self.top > self.capacity:
            raise RuntimeError(f"tinyfusers_b200 arena exhausted ({self.top} > {self.capacity} bytes)")
        return self.base + off


class Context:
    """Execution context handed down the module tree on the fast path."""

    def __init__(self, device, quirks=True, ln_strided=False):
        self.device = torch.device(device)
        self.quirks = quirks            # CrossAttention head-major reshape (reference attention.py:39)
        self.ln_strided = ln_strided    # literal reading of the reference's LayerNorm stride declaration
        self.arena = Arena()
        self.ws_bytes = 192 << 20
        self.ws = None
        self.dry = False
        self.only = None           # bench instrumentation: if a set, only kernels of these kinds are launched
        self.prof = None           # dev instrumentation: list of (key, start_event, end_event) per kernel call
        # per-forward values set by the UNet
        self.emb_bias = None       # dict: id(ResBlock) -> device pointer of its (conv bias + emb) fp32 vector
        self.context = None        # Act view of the zero-padded fp16 prompt context (B, Tpad, 768)
        self.context_tokens = 0
        self.fuse_gn = os.environ.get("TINYFUSERS_B200_FUSE_GN", "1") != "0"
        # LayerNorm folded into the consuming GEMM (row statistics from the producing GEMM's epilogue, gamma folded into the
        # weights, mean / rstd applied per row in the consumer's epilogue). Correct and tested, opt-in:
        # TINYFUSERS_B200_FUSE_LN=1 (norm2 + sub-4096-token norm1: 30 LayerNorm launches fewer, 376 -> 346) measures +0.4 % at
        # batch 2 (279.2 -> 280.2 steps/s), +0.7 % at batch 16 and +0.2 % at 768^2 in same-box A/B runs - inside the box-to-box
        # spread - by moving 0.09 ms of norm time into 0.10 ms of GEMM-epilogue time; =2 (all three norms, also the
        # epilogue-bound GEGLU projection) is slower (271.1 steps/s). Not the default: no measurable step gain.
        self.fuse_ln = os.environ.get("TINYFUSERS_B200_FUSE_LN", "0") in ("1", "2")
        self.fuse_ln_all = os.environ.get("TINYFUSERS_B200_FUSE_LN", "0") == "2"   # also norm1 at 4096 tokens and norm3 (GEGLU)
        self.ctx_kv = None         # dict: id(CrossAttention) -> (k_ptr, ldk, vt_ptr, ldvt) projected once per forward
        self.gn_fuse_max_hw = 16384

    def ensure_workspaces(self):
        if self.ws is None:
            self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.device)

    # -- allocation ------------------------------------------------------------------------------
    def new_act(self, n, h, w, c, stride=None, gn=False, gn_unit=None):
        stride = c if stride is None else stride
        ptr = self.arena.alloc(2 * n * h * w * stride)
        a = Act(ptr, n, h, w, c, stride)
        if gn:
            self.attach_gn(a, gn_unit)
        return a

    def attach_gn(self, a, unit=None):
        """Give `a` a statistics buffer its producer (a GEMM / conv epilogue) will fill, if the geometry qualifies.
        Every block of the one-launch GroupNorm folds all (rows/32) slots of its image, which stops paying beyond
        ~16k pixels per image (VAE decoder at 256^2 / 512^2): those tensors take the statistics-pass path."""
        unit = gn_unit(a.c) if unit is None else unit
        if a.h * a.w > self.gn_fuse_max_hw:
            return a
        if self.fuse_gn and unit and b200.tf_gn_stats_supported(a.n, a.h, a.w, a.c, unit, 1):
            ptr = self.arena.alloc(8 * a.n * (a.h * a.w // 32) * (a.c // unit))
            a.gn = [(0, a.c, ptr, unit)]
        return a

    def new_f32(self, numel):
        return self.arena.alloc(4 * numel)

    # -- kernel wrappers (each is one C-ABI call) ---------------------------------------------------
    def _timed(self, key, fn):
        if self.prof is None:
            return fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        self.prof.append((key, e0, e1, fn))
        return r

    def skip(self, kind):
        """True when the kernel must not be launched (arena dry run, or bench's per-class timing filter)."""
        return self.dry or (self.only is not None and kind not in self.only)

    def gemm(self, a_ptr, lda, M, K, w, N, out_ptr, ldc, bias=None, residual_ptr=None, ldr=0, flags=0, ldw=None, gn=None,
             w_static=True, row_stats=None, ln=None):
        """gn = (stats_ptr, unit, rows_per_image): also emit the GroupNorm statistics of the output.
        row_stats = ptr: also emit {sum, sumsq} per 32-column chunk of every output row (float2 [M][N/32]).
        ln = (stats_ptr, chunks, c1_ptr, eps): LayerNorm folded onto A (see tf_gemm_ex_f16); `w` / `bias` are the folded ones.
        w_static: `w` is a packed weight (packing.cached synchronises after building it), so the kernel may fetch it
        before waiting for the preceding kernel; False for swapped-operand calls whose W slot holds an activation."""
        if self.skip("gemm"):
            return
        if w_static:
            flags |= b200.TF_GEMM_W_STATIC
        ldw = K if ldw is None else ldw
        if row_stats is not None or ln is not None:
            ex = GemmExtras(gn[0] if gn else None, gn[1] if gn else 0, gn[2] if gn else 0, row_stats,
                            ln[0] if ln else None, ln[1] if ln else 0, ln[2] if ln else None, ln[3] if ln else 0.0)
            fn = lambda: b200.tf_gemm_ex_f16(a_ptr, lda, w, ldw, out_ptr, ldc, M, N, K, bias, residual_ptr, ldr, flags,
                                             self.ws.data_ptr(), self.ws_bytes, ctypes.byref(ex), stream_ptr())
        elif gn is None:
            fn = lambda: b200.tf_gemm_f16(a_ptr, lda, w, ldw, out_ptr, ldc, M, N, K, bias,
                                          residual_ptr, ldr, flags, self.ws.data_ptr(), self.ws_bytes, stream_ptr())
        else:
            fn = lambda: b200.tf_gemm_gn_f16(a_ptr, lda, w, ldw, out_ptr, ldc, M, N, K, bias,
                                             residual_ptr, ldr, flags, self.ws.data_ptr(), self.ws_bytes, gn[0], gn[1], gn[2],
                                             stream_ptr())
        st = self._timed(("gemm", M, N, K, flags & 3, residual_ptr is not None, gn[1] if gn else 0, gn[2] if gn else 0,
                          (1 if row_stats is not None else 0) | (2 if ln is not None else 0)), fn)
        b200.check(st, "tf_gemm_f16")

    def conv3x3(self, x, w, cout, out, bias=None, residual=None, stride=1, flags=0, gn=None, w_static=True):
        if self.skip("gemm"):
            return
        if w_static:
            flags |= b200.TF_GEMM_W_STATIC
        rp = residual.ptr if residual is not None else None
        rs = residual.stride if residual is not None else 0
        if gn is None:
            fn = lambda: b200.tf_conv2d_nhwc_f16(x.ptr, x.n, x.h, x.w, x.c, x.stride, w, cout, 3, stride, out.ptr, out.stride,
                                                 bias, rp, rs, flags, self.ws.data_ptr(), self.ws_bytes, stream_ptr())
        else:
            fn = lambda: b200.tf_conv2d_nhwc_gn_f16(x.ptr, x.n, x.h, x.w, x.c, x.stride, w, cout, 3, stride, out.ptr,
                                                    out.stride, bias, rp, rs, flags, self.ws.data_ptr(), self.ws_bytes,
                                                    gn[0], gn[1], stream_ptr())
        st = self._timed(("conv3x3", x.n, x.h, x.w, x.c, cout, stride, residual is not None, gn[1] if gn else 0), fn)
        b200.check(st, "tf_conv2d_nhwc_f16")

    def conv_up2x(self, x, w4, wrows, cout, out, bias=None, gn=None):
        """out (n, 2h, 2w, cout) = conv3x3(upsample2x(x)) in one launch; w4 = packing.conv_up2x_weight."""
        if self.skip("gemm"):
            return
        fn = lambda: b200.tf_conv2d_up2x_nhwc_f16(x.ptr, x.n, x.h, x.w, x.c, x.stride, w4, wrows, cout, out.ptr, out.stride, bias,
                                                  b200.TF_GEMM_W_STATIC, gn[0] if gn else None, gn[1] if gn else 0, stream_ptr())
        st = self._timed(("conv_up2x", x.n, x.h, x.w, x.c, cout, gn[1] if gn else 0), fn)
        b200.check(st, "tf_conv2d_up2x_nhwc_f16")

    def conv3x3_skip(self, x, x2, w, cout, out, bias=None, gn=None):
        """out = conv3x3(x) + conv1x1(x2) + bias in one launch; w = [3x3 weight rows | 1x1 weight rows] per output channel."""
        if self.skip("gemm"):
            return
        fn = lambda: b200.tf_conv2d_nhwc_skip_f16(x.ptr, x.n, x.h, x.w, x.c, x.stride, x2.ptr, x2.c, x2.stride, w, cout, out.ptr,
                                                  out.stride, bias, b200.TF_GEMM_W_STATIC, self.ws.data_ptr(), self.ws_bytes,
                                                  gn[0] if gn else None, gn[1] if gn else 0, stream_ptr())
        st = self._timed(("conv3x3", x.n, x.h, x.w, x.c, cout, 1, False, gn[1] if gn else 0, x2.c), fn)
        b200.check(st, "tf_conv2d_nhwc_skip_f16")

    def groupnorm(self, x, out, gamma, beta, eps, silu, groups=32):
        parts = x.gn
        if parts is not None and len(parts) in (1, 2) and x.c % groups == 0:
            cpg = x.c // groups
            ok = parts[0][0] == 0 and parts[-1][1] == x.c and all(cpg % pt[3] == 0 and pt[0] % pt[3] == 0 for pt in parts)
            if len(parts) == 2:
                ok = ok and parts[0][1] == parts[1][0]
            if ok:
                if self.skip("norm"):
                    return
                p0 = parts[0]
                p1 = parts[1] if len(parts) == 2 else None
                st = self._timed(("groupnorm", x.n, x.h * x.w, x.c),
                                 lambda: b200.tf_groupnorm_fused_nhwc_f16(
                                     x.ptr, x.stride, p0[1], p0[2], p0[3],
                                     x.ptr + 2 * p1[0] if p1 else None, x.stride, p1[1] - p1[0] if p1 else 0,
                                     p1[2] if p1 else None, p1[3] if p1 else 1, out.ptr, out.stride, x.n, x.h * x.w,
                                     groups, gamma, beta, eps, 1 if silu else 0, stream_ptr()))
                b200.check(st, "tf_groupnorm_fused_nhwc_f16")
                return
        mark = self.arena.mark()
        stats = self.arena.alloc(b200.tf_groupnorm_workspace_bytes(x.n, groups))   # partials + {mean, rstd}
        self.arena.release(mark)                          # stream order keeps the reuse safe
        if self.skip("norm"):
            return
        st = self._timed(("groupnorm", x.n, x.h * x.w, x.c),
                         lambda: b200.tf_groupnorm_nhwc_f16(x.ptr, x.stride, x.c, None, 0, 0, out.ptr, out.stride, x.n,
                                                            x.h * x.w, groups, gamma, beta, eps, 1 if silu else 0, stats,
                                                            stream_ptr()))
        b200.check(st, "tf_groupnorm_nhwc_f16")

    def layernorm(self, x_ptr, out_ptr, rows, C, gamma, beta, eps, interleave):
        if self.skip("norm"):
            return
        st = self._timed(("layernorm", rows, C, interleave),
                         lambda: b200.tf_layernorm_f16(x_ptr, out_ptr, rows, C, gamma, beta, eps, interleave, stream_ptr()))
        b200.check(st, "tf_layernorm_f16")

    def attention(self, q_ptr, ldq, k_ptr, ldk, vt_ptr, ldvt, out_ptr, B, NH, Tq, Tk, Tk_pad, d, dp, head_major):
        if self.skip("attention"):
            return
        if head_major:   # reference reshape quirk: (B,NH,T,d) memory read back as (B,T,NH*d)
            osb, osh, ost = NH * Tq * d, Tq * d, d
        else:
            osb, osh, ost = Tq * NH * d, d, NH * d
        st = self._timed(("attention", B, NH, Tq, Tk, d),
                         lambda: b200.tf_attention_f16(q_ptr, ldq, k_ptr, ldk, vt_ptr, ldvt, out_ptr, osb, osh, ost, B, NH, Tq,
                                                       Tk, Tk_pad, d, dp, 1.0 / math.sqrt(d), stream_ptr()))
        b200.check(st, "tf_attention_f16")

    def attention_v(self, q_ptr, ldq, k_ptr, ldk, v_ptr, ldv, out_ptr, B, NH, Tq, Tk, Tk_pad, d, dp, dvp, head_major, causal=False):
        """Attention with V in its natural layout (B*Tk_pad rows, head h at columns [h*dvp, (h+1)*dvp))."""
        if self.skip("attention"):
            return
        if head_major:   # reference reshape quirk: (B,NH,T,d) memory read back as (B,T,NH*d)
            osb, osh, ost = NH * Tq * d, Tq * d, d
        else:
            osb, osh, ost = Tq * NH * d, d, NH * d
        st = self._timed(("attention", B, NH, Tq, Tk, d),
                         lambda: b200.tf_attention_v_f16(q_ptr, ldq, k_ptr, ldk, v_ptr, ldv, out_ptr, osb, osh, ost, B, NH, Tq,
                                                         Tk, Tk_pad, d, dp, dvp, 1.0 / math.sqrt(d), 1 if causal else 0,
                                                         stream_ptr()))
        b200.check(st, "tf_attention_v_f16")

    def attention_causal(self, q_ptr, ldq, k_ptr, ldk, vt_ptr, ldvt, out_ptr, B, NH, T, T_pad, d, dp):
        """Causal self-attention with the canonical head merge (CLIPAttention, reference attention.py:88-99)."""
        if self.skip("attention"):
            return
        st = self._timed(("attention_causal", B, NH, T, T, d),
                         lambda: b200.tf_attention_causal_f16(q_ptr, ldq, k_ptr, ldk, vt_ptr, ldvt, out_ptr, T * NH * d, d,
                                                              NH * d, B, NH, T, T, T_pad, d, dp, 1.0 / math.sqrt(d),
                                                              stream_ptr()))
        b200.check(st, "tf_attention_causal_f16")

    def plane_attention(self, q_ptr, k_ptr, v_ptr, out_ptr, planes, H, W):
        if self.skip("attention"):
            return
        st = self._timed(("plane_attention", planes, H, W),
                         lambda: b200.tf_plane_attention_f16(q_ptr, k_ptr, v_ptr, out_ptr, planes, H, W, 1.0 / math.sqrt(W),
                                                             stream_ptr()))
        b200.check(st, "tf_plane_attention_f16")

    def to_nchw_f16(self, x, out_ptr):
        """fp16 NHWC Act -> fp16 NCHW planes at out_ptr."""
        if self.skip("misc"):
            return
        st = b200.tf_nhwc_to_nchw(x.ptr, x.stride, out_ptr, 0, x.n, x.c, x.h * x.w, stream_ptr())
        b200.check(st, "tf_nhwc_to_nchw")

    def from_nchw_f16(self, src_ptr, out):
        """fp16 NCHW planes at src_ptr -> fp16 NHWC Act."""
        if self.skip("misc"):
            return
        st = b200.tf_nchw_to_nhwc_f16(src_ptr, 0, out.ptr, out.n, out.c, out.h * out.w, out.stride, stream_ptr())
        b200.check(st, "tf_nchw_to_nhwc_f16")

    def unary(self, ptr, n, op):
        if self.skip("misc"):
            return
        b200.check(b200.tf_unary(ptr, ptr, n, op, 0, stream_ptr()), "tf_unary")

    def upsample2x(self, x, out):
        if self.skip("misc"):
            return
        st = b200.tf_upsample_nearest2x_nhwc_f16(x.ptr, x.stride, out.ptr, out.stride, x.n, x.h, x.w, x.c, stream_ptr())
        b200.check(st, "tf_upsample_nearest2x_nhwc_f16")


# ------------------------------------------------------------------------------------------------
# layout edges for the stand-alone (reference-signature) operator wrappers
# ------------------------------------------------------------------------------------------------

def nchw_to_act(x, c_pad_to=8):
    """(N,C,H,W) fp32/fp16 CUDA tensor -> fp16 NHWC Act (channels zero-padded to a multiple of c_pad_to)."""
    require_cuda(x, "x")
    if x.dtype not in (F16, F32):
        x = x.to(F32)
    x = x.contiguous()
    N, C, H, W = x.shape
    Cp = _align(C, c_pad_to)
    buf = torch.zeros((N, H, W, Cp), dtype=F16, device=x.device) if Cp != C else torch.empty((N, H, W, Cp), dtype=F16, device=x.device)
    st = b200.tf_nchw_to_nhwc_f16(x.data_ptr(), 1 if x.dtype == F32 else 0, buf.data_ptr(), N, C, H * W, Cp, stream_ptr())
    b200.check(st, "tf_nchw_to_nhwc_f16")
    return Act(buf.data_ptr(), N, H, W, Cp, Cp, keep=(buf, x))


def act_to_nchw(a, C=None, dtype=F32):
    C = a.c if C is None else C
    dev = a.keep[0].device if isinstance(a.keep, tuple) else (a.keep.device if a.keep is not None else torch.device("cuda", torch.cuda.current_device()))
    out = torch.empty((a.n, C, a.h, a.w), dtype=dtype, device=dev)
    st = b200.tf_nhwc_to_nchw(a.ptr, a.stride, out.data_ptr(), 1 if dtype == F32 else 0, a.n, C, a.h * a.w, stream_ptr())
    b200.check(st, "tf_nhwc_to_nchw")
    return out


def as_f16(x, rows_pad_to=1):
    """fp32 / fp16 CUDA activation (..., C) -> NEW contiguous fp16 tensor (same shape; with rows_pad_to > 1 a 2-D (rows, C) input
    gets zero rows appended up to a multiple of it). The conversion is a tf_* launch (tf_pad_tokens_f32_to_f16), not a torch op:
    the stand-alone operator wrappers use torch for memory only."""
    require_cuda(x, "x")
    if x.dtype not in (F16, F32):
        raise RuntimeError(f"tinyfusers_b200: activations must be fp32 or fp16, got {x.dtype}")
    xc = x.contiguous()
    C = xc.shape[-1]
    rows = xc.numel() // C
    rows_p = (rows + rows_pad_to - 1) // rows_pad_to * rows_pad_to
    shape = tuple(xc.shape) if rows_p == rows else (rows_p, C)
    out = torch.empty(shape, dtype=F16, device=xc.device)
    if xc.dtype == F16:
        if rows_p != rows:
            out.zero_()
        out.view(-1)[:xc.numel()].copy_(xc.view(-1))          # device copy, no arithmetic
        return out
    st = b200.tf_pad_tokens_f32_to_f16(xc.data_ptr(), out.data_ptr(), 1, rows, rows_p, C, stream_ptr())
    b200.check(st, "tf_pad_tokens_f32_to_f16")
    return out


def as_f32(x):
    """fp16 CUDA tensor -> new fp32 tensor of the same shape through tf_cast_f16_to_f32 (fp32 input: returned as is)."""
    require_cuda(x, "x")
    if x.dtype == F32:
        return x
    xc = x.contiguous()
    out = torch.empty(xc.shape, dtype=F32, device=xc.device)
    b200.check(b200.tf_cast_f16_to_f32(xc.data_ptr(), out.data_ptr(), xc.numel(), stream_ptr()), "tf_cast_f16_to_f32")
    return out


def tokens_to_act(x):
    """(B,T,C) CUDA tensor -> fp16 Act (n=B, h=T, w=1)."""
    xh = as_f16(x)
    B, T, C = xh.shape
    return Act(xh.data_ptr(), B, T, 1, C, C, keep=xh)


def new_act_tensor(n, h, w, c, device="cuda"):
    buf = torch.empty((n, h, w, c), dtype=F16, device=device)
    return Act(buf.data_ptr(), n, h, w, c, c, keep=buf)


_standalone_ctx = {}


def standalone_context(quirks=None):
    """Context for per-op calls outside a UNet forward (allocates from torch, not the arena)."""
    from . import get_layernorm_strided, get_quirks
    dev = torch.cuda.current_device()
    quirks = get_quirks() if quirks is None else quirks
    key = (dev, quirks, get_layernorm_strided())
    ctx = _standalone_ctx.get(key)
    if ctx is None:
        b200.init(dev)
        ctx = Context(torch.device("cuda", dev), quirks, get_layernorm_strided())
        ctx.ensure_workspaces()
        ctx.arena.reserve(int(os.environ.get("TINYFUSERS_B200_SCRATCH_MB", "1024")) << 20, ctx.device)
        _standalone_ctx[key] = ctx
    return ctx
