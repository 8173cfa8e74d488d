"""ResBlock with the reference's signature (reference: tinyfusers/vision/resnet.py:6-31).

Reference: GN -> silu -> conv3x3 -> (+ Linear(silu(emb))) -> GN -> silu -> conv3x3 -> + skip(x), i.e.
2-3 cuDNN graphs + ~20 elementwise launches. Fast path: 2 x [GroupNorm-stats, GroupNorm-apply+SiLU],
2 implicit-GEMM convs with everything else in their epilogues: the emb projection is pre-added to the
first conv's bias, the residual (identity or 1x1 skip conv) is added by the second conv's epilogue."""
import torch

from .. import fp32
from ..ff.group_norm import GroupNorm
from ..ff.linear import Linear
from ..native.b200.ops import b200
from ..runtime import F32, act_to_nchw, nchw_to_act, new_act_tensor, require_cuda, standalone_context, stream_ptr
from ..storage.tensor import Tensor
from .conv2d import Conv2d, conv3x3_plus_skip


class ResBlock:
    def __init__(self, channels, emb_channels, out_channels):
        self.in_layers = [
            GroupNorm(32, channels),
            Tensor.silu,
            Conv2d(channels, out_channels, kernel_size=[3, 3], padding=[1, 1])
        ]
        self.emb_layers = [
            Tensor.silu,
            Linear(emb_channels, out_channels)
        ]
        self.out_layers = [
            GroupNorm(32, out_channels),
            Tensor.silu,
            lambda x: x,  # keeps the checkpoint index of out_layers.3 (reference: resnet.py:20)
            Conv2d(out_channels, out_channels, kernel_size=[3, 3], padding=[1, 1])
        ]
        self.skip_connection = Conv2d(channels, out_channels, kernel_size=[1, 1]) if channels != out_channels else lambda x: x
        self.channels, self.out_channels = channels, out_channels

    def __call__(self, x, emb):
        require_cuda(x, "x")
        if fp32.enabled():
            return fp32.res_block(self, x, emb)
        ctx = standalone_context()
        ctx.arena.reset()
        a = nchw_to_act(x, c_pad_to=8)
        out = new_act_tensor(a.n, a.h, a.w, self.out_channels, device=x.device)
        # emb projection + conv bias -> per-channel fp32 bias of the first conv
        lin = self.emb_layers[1]
        w, b = lin._packed()
        conv_b = self.in_layers[2]._packed()[1]
        eb = torch.empty(self.out_channels, dtype=F32, device=x.device)
        e = emb.reshape(-1).to(F32).contiguous()
        st = b200.tf_gemv_f16w(e.data_ptr(), w.data_ptr(), b.data_ptr() if b is not None else None,
                               conv_b.data_ptr() if conv_b is not None else None, eb.data_ptr(), self.out_channels,
                               w.shape[1], 1, stream_ptr())
        b200.check(st, "tf_gemv_f16w")
        self._run(ctx, a, eb.data_ptr(), out)
        return act_to_nchw(out, self.out_channels)

    def _run(self, ctx, x, emb_bias_ptr, out):
        mark = ctx.arena.mark()
        h0 = ctx.new_act(x.n, x.h, x.w, x.c)
        self.in_layers[0]._run(ctx, x, h0, silu=True)
        h1 = ctx.new_act(x.n, x.h, x.w, self.out_channels, gn=True)   # conv1's epilogue leaves GN2's statistics
        self.in_layers[2]._run(ctx, h0, h1, bias_ptr=emb_bias_ptr)
        h2 = ctx.new_act(x.n, x.h, x.w, self.out_channels)
        self.out_layers[0]._run(ctx, h1, h2, silu=True)
        if isinstance(self.skip_connection, Conv2d):
            # the 1x1 skip convolution rides along the second 3x3 convolution as extra k-blocks (one launch, no residual read)
            if not conv3x3_plus_skip(ctx, self, self.out_layers[3], self.skip_connection, h2, x, out):
                res = ctx.new_act(x.n, x.h, x.w, self.out_channels)
                self.skip_connection._run(ctx, x, res)
                self.out_layers[3]._run(ctx, h2, out, residual=res)
        else:
            self.out_layers[3]._run(ctx, h2, out, residual=x)
        ctx.arena.release(mark)
        return out


class ResnetBlock:
    """VAE residual block (reference: tinyfusers/vision/resnet.py:33-45): conv1(swish(norm1 x)) -> conv2(swish(norm2 .))
    + nin_shortcut(x) (1x1 conv when the channel count changes). Fast path: 2 x GroupNorm+SiLU, 2 implicit-GEMM convs;
    conv1's epilogue leaves norm2's statistics, the shortcut is conv2's epilogue residual."""

    def __init__(self, in_channels, out_channels=None):
        out_channels = in_channels if out_channels is None else out_channels
        self.norm1 = GroupNorm(32, in_channels)
        self.conv1 = Conv2d(in_channels, out_channels, kernel_size=[3, 3], padding=[1, 1])
        self.norm2 = GroupNorm(32, out_channels)
        self.conv2 = Conv2d(out_channels, out_channels, kernel_size=[3, 3], padding=[1, 1])
        self.nin_shortcut = Conv2d(in_channels, out_channels, kernel_size=[1, 1]) if in_channels != out_channels else lambda x: x
        self.in_channels, self.out_channels = in_channels, out_channels

    def __call__(self, x):
        require_cuda(x, "x")
        if fp32.enabled():
            return fp32.resnet_block(self, x)
        ctx = standalone_context()
        ctx.arena.reset()
        a = nchw_to_act(x, c_pad_to=8)
        out = new_act_tensor(a.n, a.h, a.w, self.out_channels, device=x.device)
        self._run(ctx, a, out)
        return act_to_nchw(out, self.out_channels)

    def _run(self, ctx, x, out):
        mark = ctx.arena.mark()
        h0 = ctx.new_act(x.n, x.h, x.w, x.c)
        self.norm1._run(ctx, x, h0, silu=True)
        h1 = ctx.new_act(x.n, x.h, x.w, self.out_channels, gn=True, gn_unit=self.out_channels // 32)
        self.conv1._run(ctx, h0, h1)
        h2 = ctx.new_act(x.n, x.h, x.w, self.out_channels)
        self.norm2._run(ctx, h1, h2, silu=True)
        if isinstance(self.nin_shortcut, Conv2d):
            if not conv3x3_plus_skip(ctx, self, self.conv2, self.nin_shortcut, h2, x, out):
                res = ctx.new_act(x.n, x.h, x.w, self.out_channels)
                self.nin_shortcut._run(ctx, x, res)
                self.conv2._run(ctx, h2, out, residual=res)
        else:
            self.conv2._run(ctx, h2, out, residual=x)
        ctx.arena.release(mark)
        return out
