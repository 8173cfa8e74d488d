"""VAE mid block with the reference's signature (reference: tinyfusers/vae/mid.py:5-12)."""
from .. import fp32
from ..attention.attention import AttnBlock
from ..runtime import act_to_nchw, nchw_to_act, new_act_tensor, require_cuda, standalone_context
from ..vision.resnet import ResnetBlock


class Mid:
    def __init__(self, block_in):
        self.block_1 = ResnetBlock(block_in, block_in)
        self.attn_1 = AttnBlock(block_in)
        self.block_2 = ResnetBlock(block_in, block_in)
        self.block_in = block_in

    def __call__(self, x):
        require_cuda(x, "x")
        if fp32.enabled():
            return fp32.vae_mid(self, x)
        ctx = standalone_context()
        ctx.arena.reset()
        a = nchw_to_act(x, c_pad_to=8)
        out = new_act_tensor(a.n, a.h, a.w, a.c, device=x.device)
        self._run(ctx, a, out)
        return act_to_nchw(out, x.shape[1])

    def _run(self, ctx, x, out):
        mark = ctx.arena.mark()
        u = self.block_in // 32
        h1 = ctx.new_act(x.n, x.h, x.w, x.c, gn=True, gn_unit=u)
        self.block_1._run(ctx, x, h1)
        h2 = ctx.new_act(x.n, x.h, x.w, x.c, gn=True, gn_unit=u)
        self.attn_1._run(ctx, h1, h2)
        self.block_2._run(ctx, h2, out)
        ctx.arena.release(mark)
        return out
