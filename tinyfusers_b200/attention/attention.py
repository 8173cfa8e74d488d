"""SpatialTransformer / BasicTransformerBlock / CrossAttention with the reference's signatures
(reference: tinyfusers/attention/attention.py:26-76).

Fast path per transformer block (reference: 9 cuBLAS GEMMs, 2 materialised-score SDPAs, 3 cuDNN
LayerNorm graph builds, ~25 elementwise launches):
  LN -> [QK proj GEMM, V^T proj GEMM (operands swapped so V arrives transposed)] -> fused attention
     -> out-proj GEMM (+bias +residual, in place)           x2 (self, cross)
  LN -> GEGLU GEMM (fused gate) -> out GEMM (+bias +residual, in place)
In NHWC the reference's (B,C,HW)->(B,HW,C) transposes (attention.py:70,73) are free reinterpretations,
and the 1x1 proj_in / proj_out convs are plain GEMMs with bias / residual epilogues."""
import math

import torch

from .. import fp32, get_quirks, packing
from ..native.b200.ops import b200
from ..ff.group_norm import GroupNorm
from ..ff.layer_norm import LayerNorm
from ..ff.linear import Linear
from ..ff.nn import FeedForward
from ..runtime import (F16, F32, Act, act_to_nchw, as_f16, as_f32, nchw_to_act, new_act_tensor, require_cuda,
                       standalone_context, stream_ptr, tokens_to_act)
from ..vision.conv2d import Conv2d
from .sdpa import scaled_dot_product_attention  # noqa: F401  (re-exported like the reference)


def _pad16(d):
    return (d + 15) // 16 * 16


def _pad64(d):
    """Padded V head dim: a multiple of 64 (128-byte swizzle atoms, one TMA box per key block) unless that is more than 1.5x
    the 16-multiple (32-byte atoms): 40 -> 64, 64 -> 64, 80 -> 80, 160 -> 192. Measured at d = 40: 64 columns give 260.8
    steps/s at batch 2, 48 columns 260.2 (three boxes per block, slower attention) but +0.7 % at batch 16."""
    p64, p16 = (d + 63) // 64 * 64, (d + 15) // 16 * 16
    return p64 if p64 * 2 <= p16 * 3 else p16


def fold_norm1(rows):
    """LayerNorm -> [Q|K|V] projection fold: wins up to 2048-row (batch 2 x 1024 tokens) inputs, loses at 8192 rows."""
    return rows <= 4096


class CrossAttention:
    def __init__(self, query_dim, context_dim, n_heads, d_head):
        self.to_q = Linear(query_dim, n_heads * d_head, bias=False)
        self.to_k = Linear(context_dim, n_heads * d_head, bias=False)
        self.to_v = Linear(context_dim, n_heads * d_head, bias=False)
        self.num_heads = n_heads
        self.head_size = d_head
        self.to_out = [Linear(n_heads * d_head, query_dim)]

    def _packed(self):
        """wq, wk: each head's rows zero-padded to dp (multiple of 16: the QK^T slabs); wv: to dvp (multiple of 64: V is the
        MN-major operand of P.V, read in its natural layout); wkv = [wk ; wv]; wqkv = [wq ; wk ; wv] (self-attention)."""
        nh, d = self.num_heads, self.head_size
        dp, dvp = _pad16(d), _pad64(d)
        def build():
            wq = packing.head_pad(self.to_q.weight, nh, d, dp)
            wk = packing.head_pad(self.to_k.weight, nh, d, dp)
            wv = packing.head_pad(self.to_v.weight, nh, d, dvp)
            wkv = torch.cat((wk, wv), dim=0).contiguous()
            wqkv = torch.cat((wq, wk, wv), dim=0).contiguous() if wq.shape[1] == wk.shape[1] else None
            return wq, wk, wv, wkv, wqkv
        return packing.cached(self, "attn", (self.to_q.weight, self.to_k.weight, self.to_v.weight), build)

    def _packed_ln(self, norm, self_attention):
        """The first projection with `norm` (the LayerNorm in front of this attention) folded in: (W', c1, c2)."""
        def build():
            packs = self._packed()
            w = packs[4] if self_attention else packs[0]      # [wq ; wk ; wv] or wq
            return packing.ln_fold(w, norm.weight, norm.bias)
        return packing.cached(self, "attn_ln_self" if self_attention else "attn_ln_q",
                              (self.to_q.weight, self.to_k.weight, self.to_v.weight, norm.weight, norm.bias), build)

    def __call__(self, x, context=None):
        require_cuda(x, "x")
        if fp32.enabled():
            return fp32.cross_attention(self, x, context)
        ctx = standalone_context()
        ctx.arena.reset()
        xa = tokens_to_act(x)
        ca = None if context is None else _pad_context(ctx, context)
        B, T, C = x.shape
        out = torch.empty((B, T, self.to_out[0].weight.shape[0]), dtype=F16, device=x.device)
        self._run(ctx, xa.ptr, B, T, C, out.data_ptr(), context=ca, residual=False)
        return as_f32(out)

    # h_ptr (B*T, C) is updated:  h <- to_out(attention(...)) (+ h if residual);  xn_ptr = normalised input
    # ln = (row statistics of the UN-normalised input at xn_ptr, LayerNorm module): the norm is folded into the first GEMM.
    # h_stats: where the out-projection leaves the row statistics of the updated h (for the next folded LayerNorm).
    def _run(self, ctx, xn_ptr, B, T, C, h_ptr, context=None, residual=True, ln=None, h_stats=None):
        nh, d = self.num_heads, self.head_size
        dp, dvp = _pad16(d), _pad64(d)
        wq, wk, wv, wkv, wqkv = self._packed()
        mark = ctx.arena.mark()
        M = B * T
        lnx = None
        if ln is not None:
            wf, c1, c2 = self._packed_ln(ln[1], context is None)
            lnx = (ln[0], C // 32, c1.data_ptr(), float(ln[1].eps.reshape(-1)[0]))
        if context is None:
            # self-attention: ONE GEMM for [Q | K | V]; the attention kernel reads V in this natural layout
            n_all = 2 * nh * dp + nh * dvp
            qkv_ptr = ctx.arena.alloc(2 * M * n_all)
            if lnx is not None:
                ctx.gemm(xn_ptr, C, M, C, wf.data_ptr(), n_all, qkv_ptr, n_all, bias=c2.data_ptr(), ln=lnx)
            else:
                ctx.gemm(xn_ptr, C, M, C, wqkv.data_ptr(), n_all, qkv_ptr, n_all)
            q_ptr, k_ptr, v_ptr = qkv_ptr, qkv_ptr + 2 * nh * dp, qkv_ptr + 4 * nh * dp
            ldq = ldk = ldv = n_all
            Tk, Tkp = T, T
        else:
            Tkp, Cc = context.h, context.c
            Tk = context.valid if context.valid is not None else context.h
            Mc = B * Tkp
            q_ptr = ctx.arena.alloc(2 * M * nh * dp)
            if lnx is not None:
                ctx.gemm(xn_ptr, C, M, C, wf.data_ptr(), nh * dp, q_ptr, nh * dp, bias=c2.data_ptr(), ln=lnx)
            else:
                ctx.gemm(xn_ptr, C, M, C, wq.data_ptr(), nh * dp, q_ptr, nh * dp)
            ldq = nh * dp
            pre = ctx.ctx_kv.get(id(self)) if ctx.ctx_kv else None
            if pre is not None:      # projected once per step for all blocks (UNetModel._ctx_kv_pack)
                k_ptr, ldk, v_ptr, ldv = pre
            else:
                n_kv = nh * (dp + dvp)
                kv_ptr = ctx.arena.alloc(2 * Mc * n_kv)
                ctx.gemm(context.ptr, context.stride, Mc, Cc, wkv.data_ptr(), n_kv, kv_ptr, n_kv)
                k_ptr, v_ptr, ldk, ldv = kv_ptr, kv_ptr + 2 * nh * dp, n_kv, n_kv
        a_ptr = ctx.arena.alloc(2 * M * nh * d)
        ctx.attention_v(q_ptr, ldq, k_ptr, ldk, v_ptr, ldv, a_ptr, B, nh, T, Tk, Tkp, d, dp, dvp, head_major=ctx.quirks)
        wo, bo = self.to_out[0]._packed()
        ctx.gemm(a_ptr, nh * d, M, nh * d, wo.data_ptr(), wo.shape[0], h_ptr, wo.shape[0],
                 bias=bo.data_ptr() if bo is not None else None, residual_ptr=h_ptr if residual else None,
                 ldr=wo.shape[0], row_stats=h_stats)
        ctx.arena.release(mark)


class BasicTransformerBlock:
    def __init__(self, dim, context_dim, n_heads, d_head):
        self.attn1 = CrossAttention(dim, dim, n_heads, d_head)
        self.ff = FeedForward(dim)
        self.attn2 = CrossAttention(dim, context_dim, n_heads, d_head)
        self.norm1 = LayerNorm(dim)
        self.norm2 = LayerNorm(dim)
        self.norm3 = LayerNorm(dim)

    def __call__(self, x, context=None):
        require_cuda(x, "x")
        if fp32.enabled():
            return fp32.basic_transformer_block(self, x, context)
        ctx = standalone_context()
        ctx.arena.reset()
        B, T, C = x.shape
        h = as_f16(x)            # a new fp16 buffer (tf_* cast): the block updates it in place
        ca = None
        if context is not None:
            ca = _pad_context(ctx, context)
        self._run(ctx, h.data_ptr(), B, T, C, ca)
        return as_f32(h)

    # h (B*T, C) fp16 updated in place. h_stats: row statistics of h left by its producer (tf_gemm_ex_f16 row_stats_out):
    # with them the three LayerNorms are folded into the GEMMs that consume them (no LayerNorm launch, no xn buffer).
    def _run(self, ctx, h_ptr, B, T, C, context, h_stats=None):
        mark = ctx.arena.mark()
        if h_stats is not None and ctx.fuse_ln and not ctx.ln_strided and C % 32 == 0:
            # Measured per shape (tools/shape_table.py): the fold costs the consumer's epilogue a per-tile statistics fold
            # plus two FMAs per element, which pays for the narrow Q projection everywhere (norm2) and for the [Q|K|V]
            # projection below 4096 tokens (norm1), but not for the epilogue-bound GEGLU projection (norm3 stays a kernel).
            if ctx.fuse_ln_all:
                self.attn1._run(ctx, h_ptr, B, T, C, h_ptr, ln=(h_stats, self.norm1), h_stats=h_stats)
                self.attn2._run(ctx, h_ptr, B, T, C, h_ptr, context=context, ln=(h_stats, self.norm2), h_stats=h_stats)
                self.ff._run(ctx, h_ptr, h_ptr, B * T, C, ln=(h_stats, self.norm3))
                ctx.arena.release(mark)
                return
            xn = None
            if fold_norm1(B * T):
                self.attn1._run(ctx, h_ptr, B, T, C, h_ptr, ln=(h_stats, self.norm1), h_stats=h_stats)
            else:
                xn = ctx.arena.alloc(2 * B * T * C)
                self.norm1._run(ctx, h_ptr, xn, B, T, C)
                self.attn1._run(ctx, xn, B, T, C, h_ptr, h_stats=h_stats)
            self.attn2._run(ctx, h_ptr, B, T, C, h_ptr, context=context, ln=(h_stats, self.norm2))
            xn = ctx.arena.alloc(2 * B * T * C) if xn is None else xn
            self.norm3._run(ctx, h_ptr, xn, B, T, C)
            self.ff._run(ctx, xn, h_ptr, B * T, C)
        else:
            xn = ctx.arena.alloc(2 * B * T * C)
            self.norm1._run(ctx, h_ptr, xn, B, T, C)
            self.attn1._run(ctx, xn, B, T, C, h_ptr)
            self.norm2._run(ctx, h_ptr, xn, B, T, C)
            self.attn2._run(ctx, xn, B, T, C, h_ptr, context=context)
            self.norm3._run(ctx, h_ptr, xn, B, T, C)
            self.ff._run(ctx, xn, h_ptr, B * T, C)
        ctx.arena.release(mark)


def _pad_context(ctx, context):
    """(B,Tk,C) fp32/fp16 prompt embeddings -> zero-padded fp16 Act (B, Tk_pad, C) with Tk remembered."""
    from ..native.b200.ops import b200
    from ..runtime import stream_ptr
    B, Tk, Cc = context.shape
    Tkp = (Tk + 7) // 8 * 8
    buf = torch.empty((B, Tkp, Cc), dtype=F16, device=context.device)
    c32 = as_f32(context).contiguous()
    st = b200.tf_pad_tokens_f32_to_f16(c32.data_ptr(), buf.data_ptr(), B, Tk, Tkp, Cc, stream_ptr())
    b200.check(st, "tf_pad_tokens_f32_to_f16")
    act = Act(buf.data_ptr(), B, Tkp, 1, Cc, Cc, keep=(buf, c32))
    act.valid = Tk
    return act


class SpatialTransformer:
    def __init__(self, channels, context_dim, n_heads, d_head):
        self.norm = GroupNorm(32, channels)
        assert channels == n_heads * d_head
        self.proj_in = Conv2d(channels, n_heads * d_head, kernel_size=[1, 1])
        self.transformer_blocks = [BasicTransformerBlock(channels, context_dim, n_heads, d_head)]
        self.proj_out = Conv2d(n_heads * d_head, channels, kernel_size=[1, 1])

    def __call__(self, x, context=None):
        require_cuda(x, "x")
        if fp32.enabled():
            return fp32.spatial_transformer(self, x, context)
        ctx = standalone_context()
        ctx.arena.reset()
        a = nchw_to_act(x, c_pad_to=8)
        ca = _pad_context(ctx, context) if context is not None else None
        out = new_act_tensor(a.n, a.h, a.w, a.c, device=x.device)
        self._run(ctx, a, ca, out)
        return act_to_nchw(out, x.shape[1])

    def _run(self, ctx, x, context, out):
        mark = ctx.arena.mark()
        B, T, C = x.n, x.h * x.w, x.c
        hn = ctx.new_act(x.n, x.h, x.w, C)
        self.norm._run(ctx, x, hn, silu=False)
        h = ctx.new_act(x.n, x.h, x.w, C)
        h_stats = None
        if ctx.fuse_ln and not ctx.ln_strided and C % 32 == 0:
            h_stats = ctx.arena.alloc(8 * B * T * (C // 32))     # float2 per 32-column chunk of every token row
        self.proj_in._run(ctx, hn, h, row_stats=h_stats if (fold_norm1(B * T) or ctx.fuse_ln_all) else None)
        for block in self.transformer_blocks:
            block._run(ctx, h.ptr, B, T, C, context, h_stats=h_stats)
        self.proj_out._run(ctx, h, out, residual=x)
        ctx.arena.release(mark)
        return out


class AttnBlock:
    """VAE attention block (reference: tinyfusers/attention/attention.py:10-24).

    The reference passes the 4-D (B,C,H,W) q/k/v straight into scaled_dot_product_attention, which reads them as
    (B, NH = C, T = H, HS = W): every channel plane attends over its own rows (SURVEY.md section 8 parity note 3).
    That is what runs here under `set_quirks(True)` (default): GroupNorm -> ONE GEMM for [q | k | v] (bias fused)
    -> NHWC->NCHW planes -> tf_plane_attention_f16 -> NCHW->NHWC -> proj_out GEMM (+bias +x).
    The canonical LDM block (`set_quirks(False)`: one head over H*W pixels, head dim C = 512) runs as two tcgen05 GEMMs
    around a row softmax (`_attn_block_canonical`): the fused attention kernel stops at head dim 256."""

    def __init__(self, in_channels):
        self.norm = GroupNorm(32, in_channels)
        self.q = Conv2d(in_channels, in_channels, kernel_size=[1, 1])
        self.k = Conv2d(in_channels, in_channels, kernel_size=[1, 1])
        self.v = Conv2d(in_channels, in_channels, kernel_size=[1, 1])
        self.proj_out = Conv2d(in_channels, in_channels, kernel_size=[1, 1])
        self.in_channels = in_channels

    def _packed(self):
        def build():
            w = torch.cat([packing.conv1x1_weight(m.weight, 8) for m in (self.q, self.k, self.v)], dim=0).contiguous()
            b = torch.cat([packing.f32(m.bias) for m in (self.q, self.k, self.v)]).contiguous()
            return w, b
        return packing.cached(self, "qkv", (self.q.weight, self.k.weight, self.v.weight, self.q.bias, self.k.bias,
                                            self.v.bias), build)

    def __call__(self, x):
        require_cuda(x, "x")
        if fp32.enabled():
            return fp32.attn_block(self, x)
        ctx = standalone_context()
        ctx.arena.reset()
        a = nchw_to_act(x, c_pad_to=8)
        out = new_act_tensor(a.n, a.h, a.w, a.c, device=x.device)
        self._run(ctx, a, out)
        return act_to_nchw(out, x.shape[1])

    def _run(self, ctx, x, out):
        if not ctx.quirks:
            return self._run_canonical(ctx, x, out)
        C, H, W = x.c, x.h, x.w
        if C % 8 != 0 or W % 2 != 0:
            raise RuntimeError(f"tinyfusers_b200 AttnBlock: needs channels % 8 == 0 and an even width (C={C}, W={W})")
        mark = ctx.arena.mark()
        hn = ctx.new_act(x.n, H, W, C)
        self.norm._run(ctx, x, hn, silu=False)
        w, b = self._packed()
        qkv = ctx.new_act(x.n, H, W, 3 * C)
        ctx.gemm(hn.ptr, hn.stride, hn.rows, C, w.data_ptr(), 3 * C, qkv.ptr, 3 * C, bias=b.data_ptr())
        planes = ctx.arena.alloc(2 * x.n * 3 * C * H * W)       # (n, 3C, H, W): q planes, k planes, v planes per image
        ctx.to_nchw_f16(qkv, planes)
        o_planes = ctx.arena.alloc(2 * x.n * C * H * W)
        pb = 2 * C * H * W
        for i in range(x.n):
            base = planes + i * 3 * pb
            ctx.plane_attention(base, base + pb, base + 2 * pb, o_planes + i * pb, C, H, W)
        o = ctx.new_act(x.n, H, W, C)
        ctx.from_nchw_f16(o_planes, o)
        self.proj_out._run(ctx, o, out, residual=x)
        ctx.arena.release(mark)
        return out


def _attn_block_canonical(self, ctx, x, out):
    """The canonical LDM AttnBlock (`set_quirks(False)`, real checkpoints): ONE head over the H*W pixels, head dim = C.
    C = 512 is beyond the fused attention kernel (head dim <= 256), so it runs as the reference structures SDPA
    (attention/sdpa.py:62-76) on the tcgen05 GEMM kernel: S = Q K^T (fp32 out), row softmax -> fp16 P, O = P V with V^T
    from the NHWC->NCHW transpose; per image. 34 GFLOP at 64x64 / C = 512."""
    C, H, W = x.c, x.h, x.w
    T = H * W
    if C % 8 != 0 or T % 8 != 0:
        raise RuntimeError(f"tinyfusers_b200 AttnBlock (canonical): needs channels % 8 == 0 and H*W % 8 == 0 (C={C}, HW={T})")
    mark = ctx.arena.mark()
    hn = ctx.new_act(x.n, H, W, C)
    self.norm._run(ctx, x, hn, silu=False)
    w, b = self._packed()
    qkv = ctx.new_act(x.n, H, W, 3 * C)
    ctx.gemm(hn.ptr, hn.stride, hn.rows, C, w.data_ptr(), 3 * C, qkv.ptr, 3 * C, bias=b.data_ptr())
    vt = ctx.arena.alloc(2 * x.n * C * T)                      # (n, C, T): V^T per image
    ctx.to_nchw_f16(qkv.channels(2 * C, 3 * C), vt)
    scores = ctx.arena.alloc(4 * T * T)
    probs = ctx.arena.alloc(2 * T * T)
    o = ctx.new_act(x.n, H, W, C)
    scale = 1.0 / math.sqrt(C)
    for i in range(x.n):
        q_ptr = qkv.ptr + 2 * i * T * 3 * C
        ctx.gemm(q_ptr, 3 * C, T, C, q_ptr + 2 * C, T, scores, T, flags=b200.TF_EPI_OUT_F32, ldw=3 * C, w_static=False)
        if not ctx.skip("misc"):
            st = b200.tf_softmax_rows_f32_to_f16(scores, T, probs, T, T, T, scale, stream_ptr())
            b200.check(st, "tf_softmax_rows_f32_to_f16")
        ctx.gemm(probs, T, T, T, vt + 2 * i * C * T, C, o.ptr + 2 * i * T * o.stride, o.stride, ldw=T, w_static=False)
    self.proj_out._run(ctx, o, out, residual=x)
    ctx.arena.release(mark)
    return out


AttnBlock._run_canonical = _attn_block_canonical


def _check_causal_mask(mask, T):
    """None, or the (.., T, T) causal mask the reference builds (additive triu(-inf, 1) or boolean lower-triangular keep)."""
    if mask is None:
        return
    m = torch.as_tensor(mask)
    if m.numel() != T * T:
        raise RuntimeError(f"tinyfusers_b200 CLIPAttention: mask of {tuple(m.shape)} for {T} tokens; only the causal mask is built")
    m = m.reshape(T, T)
    tril = torch.ones((T, T), dtype=torch.bool, device=m.device).tril()
    if m.dtype == torch.bool:
        ok = torch.equal(m, tril)
    else:
        ok = bool((m[tril] == 0).all()) and bool(torch.isinf(m[~tril]).all()) and bool((m[~tril] < 0).all())
    if not ok:
        raise RuntimeError("tinyfusers_b200 CLIPAttention: only the causal mask (triu(-inf, k=1) or its boolean form) is built "
                           "into the attention kernel; use set_precision('fp32') for an arbitrary additive mask")


class CLIPAttention:
    """CLIP self-attention (reference: tinyfusers/attention/attention.py:78-99): 12 heads x 64, q/k/v/out Linears with
    bias, additive causal mask, heads merged canonically. Fast path: ONE GEMM for [q | k | v] (bias fused), causal
    tcgen05 flash attention reading V in that natural layout, out_proj GEMM (+bias +residual)."""

    def __init__(self):
        self.embed_dim = 768
        self.num_heads = 12
        self.head_dim = self.embed_dim // self.num_heads
        self.k_proj = Linear(self.embed_dim, self.embed_dim)
        self.v_proj = Linear(self.embed_dim, self.embed_dim)
        self.q_proj = Linear(self.embed_dim, self.embed_dim)
        self.out_proj = Linear(self.embed_dim, self.embed_dim)

    def _packed(self):
        mods = (self.q_proj, self.k_proj, self.v_proj, self.out_proj)
        def build():
            wqkv = torch.cat((self.q_proj.weight, self.k_proj.weight, self.v_proj.weight), dim=0).to(F16).contiguous()
            bqkv = torch.cat([packing.f32(m.bias) for m in (self.q_proj, self.k_proj, self.v_proj)]).contiguous()
            wo = self.out_proj.weight.to(F16).contiguous()
            bo = packing.f32(self.out_proj.bias)
            return wqkv, bqkv, wo, bo
        return packing.cached(self, "clipattn", tuple(m.weight for m in mods) + tuple(m.bias for m in mods), build)

    def __call__(self, hidden_states, causal_attention_mask=None):
        """(B, T, 768) -> (B, T, 768). The kernel applies the causal mask the reference always passes (vae/encoder.py:79:
        triu(-inf, k = 1)); `causal_attention_mask` may be that mask (or its boolean form) or None - any other mask raises,
        it is not silently replaced (arbitrary additive masks exist only in the fp32 parity mode)."""
        require_cuda(hidden_states, "hidden_states")
        if fp32.enabled():
            return fp32.clip_attention(self, hidden_states)
        _check_causal_mask(causal_attention_mask, hidden_states.shape[1])
        ctx = standalone_context()
        B, T, E = hidden_states.shape
        Tp = (T + 7) // 8 * 8
        out = torch.empty((B, T, E), dtype=F32, device=hidden_states.device)
        for i in range(B):
            ctx.arena.reset()
            xn = as_f16(hidden_states[i], rows_pad_to=8)          # (Tp, E), pad rows zero
            h = torch.empty((Tp, E), dtype=F16, device=hidden_states.device)
            self._run(ctx, xn.data_ptr(), h.data_ptr(), T, Tp, residual=False)
            out[i].copy_(as_f32(h)[:T])
        return out

    # xn: (Tp, E) fp16 normalised input; h (Tp, E): h <- out_proj(attn) (+ h)
    def _run(self, ctx, xn_ptr, h_ptr, T, Tp, residual=True):
        E, NH, D = self.embed_dim, self.num_heads, self.head_dim
        wqkv, bqkv, wo, bo = self._packed()
        mark = ctx.arena.mark()
        qkv = ctx.arena.alloc(2 * Tp * 3 * E)
        ctx.gemm(xn_ptr, E, T, E, wqkv.data_ptr(), 3 * E, qkv, 3 * E, bias=bqkv.data_ptr())     # [q | k | v], bias fused
        a = ctx.arena.alloc(2 * Tp * E)
        ctx.attention_v(qkv, 3 * E, qkv + 2 * E, 3 * E, qkv + 4 * E, 3 * E, a, 1, NH, T, T, T, D, D, D, head_major=False,
                        causal=True)
        ctx.gemm(a, E, T, E, wo.data_ptr(), E, h_ptr, E, bias=bo.data_ptr(), residual_ptr=h_ptr if residual else None, ldr=E)
        ctx.arena.release(mark)
