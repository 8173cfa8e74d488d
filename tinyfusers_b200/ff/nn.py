"""GEGLU / FeedForward with the reference's signatures (reference: tinyfusers/ff/nn.py:5-23).

Fast path: the GEGLU projection, its bias, the split and `value * gelu_tanh(gate)` are ONE tcgen05 GEMM
(TF_EPI_GEGLU epilogue over row-interleaved weights) — the (B,T,8C) intermediate of the reference
(84 MB fp32 at 64x64) never exists. The output projection fuses bias and the transformer residual."""
import torch

from .. import fp32, packing
from ..native.b200.ops import b200
from ..runtime import F16, F32, as_f16, as_f32, require_cuda, standalone_context
from .linear import Linear


class GEGLU:
    def __init__(self, dim_in, dim_out):
        self.proj = Linear(dim_in, dim_out * 2)
        self.dim_out = dim_out

    def _packed(self):
        return packing.cached(self, "geglu", (self.proj.weight, self.proj.bias),
                              lambda: packing.geglu_pack(self.proj.weight, self.proj.bias))

    def __call__(self, x):
        require_cuda(x, "x")
        if fp32.enabled():
            return fp32.geglu(self, x)
        ctx = standalone_context()
        a = as_f16(x.reshape(-1, x.shape[-1]))
        out = torch.empty((a.shape[0], self.dim_out), dtype=F16, device=x.device)
        self._run(ctx, a.data_ptr(), a.shape[1], a.shape[0], out.data_ptr())
        return as_f32(out).reshape(*x.shape[:-1], self.dim_out)

    def _packed_ln(self, norm):
        """The projection with the preceding LayerNorm folded in, in the TF_EPI_GEGLU row order: (W', c1, c2)."""
        def build():
            wf, c1, c2 = packing.ln_fold(self.proj.weight.to(packing.F16), norm.weight, norm.bias, self.proj.bias)
            wp, bp = packing.geglu_pack(wf, c2)
            _, c1p = packing.geglu_pack(wf, c1)
            return wp, c1p, bp
        return packing.cached(self, "geglu_ln", (self.proj.weight, self.proj.bias, norm.weight, norm.bias), build)

    def _run(self, ctx, a_ptr, lda, M, out_ptr, ln=None):
        if ln is not None:
            w, c1, c2 = self._packed_ln(ln[1])
            K = w.shape[1]
            ctx.gemm(a_ptr, lda, M, K, w.data_ptr(), 2 * self.dim_out, out_ptr, self.dim_out, bias=c2.data_ptr(),
                     flags=b200.TF_EPI_GEGLU, ln=(ln[0], K // 32, c1.data_ptr(), float(ln[1].eps.reshape(-1)[0])))
            return
        w, b = self._packed()
        K = w.shape[1]
        ctx.gemm(a_ptr, lda, M, K, w.data_ptr(), 2 * self.dim_out, out_ptr, self.dim_out,
                 bias=b.data_ptr() if b is not None else None, flags=b200.TF_EPI_GEGLU)


class FeedForward:
    def __init__(self, dim, mult=4):
        self.net = [
            GEGLU(dim, dim * mult),
            lambda x: x,  # keeps the checkpoint index of net.2 (reference: nn.py:18)
            Linear(dim * mult, dim)
        ]

    def __call__(self, x):
        if fp32.enabled():
            return fp32.feed_forward(self, x)
        h = self.net[0](x)
        return self.net[2](h)

    # fast path: h (M, dim) fp16 in place:  h <- Linear(GEGLU(xn)) + h
    def _run(self, ctx, xn_ptr, h_ptr, M, dim, ln=None):
        inner = self.net[0].dim_out
        mark = ctx.arena.mark()
        g_ptr = ctx.arena.alloc(2 * M * inner)
        self.net[0]._run(ctx, xn_ptr, dim, M, g_ptr, ln=ln)
        w, b = self.net[2]._packed()
        ctx.gemm(g_ptr, inner, M, inner, w.data_ptr(), dim, h_ptr, dim, bias=b.data_ptr() if b is not None else None,
                 residual_ptr=h_ptr, ldr=dim)
        ctx.arena.release(mark)


class CLIPMLP:
    """reference: tinyfusers/ff/nn.py:25-34 — fc1 -> quick_gelu -> fc2."""

    def __init__(self):
        self.fc1 = Linear(768, 3072)
        self.fc2 = Linear(3072, 768)

    def __call__(self, hidden_states):
        require_cuda(hidden_states, "hidden_states")
        if fp32.enabled():
            return fp32.clip_mlp(self, hidden_states)
        ctx = standalone_context()
        ctx.arena.reset()
        x2 = as_f16(hidden_states.reshape(-1, 768))
        h = torch.empty_like(x2)
        self._run(ctx, x2.data_ptr(), h.data_ptr(), x2.shape[0], residual=False)
        return as_f32(h).reshape(hidden_states.shape)

    # h (M, 768) fp16:  h <- fc2(quick_gelu(fc1(xn))) (+ h)
    def _run(self, ctx, xn_ptr, h_ptr, M, residual=True):
        w1, b1 = self.fc1._packed()
        w2, b2 = self.fc2._packed()
        mark = ctx.arena.mark()
        m = ctx.arena.alloc(2 * M * 3072)
        ctx.gemm(xn_ptr, 768, M, 768, w1.data_ptr(), 3072, m, 3072, bias=b1.data_ptr() if b1 is not None else None)
        ctx.unary(m, M * 3072, 3)
        ctx.gemm(m, 3072, M, 3072, w2.data_ptr(), 768, h_ptr, 768, bias=b2.data_ptr() if b2 is not None else None,
                 residual_ptr=h_ptr if residual else None, ldr=768)
        ctx.arena.release(mark)
