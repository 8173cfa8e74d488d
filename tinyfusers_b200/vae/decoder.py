"""VAE decoder with the reference's signature (reference: tinyfusers/vae/decoder.py:8-34).

conv_in (4 -> 512) -> Mid -> 4 up stages of 3 ResnetBlocks (512, 512, 256, 128 channels; nearest x2 + 3x3 conv between
stages) -> GroupNorm + swish -> conv_out (128 -> 3). 2.47 TFLOP per 512^2 image, 99 % of it in 3x3 convolutions
that run on the same tcgen05 implicit-GEMM kernel as the UNet (M up to 262 144 pixels). Activations are fp16 NHWC
in a stack arena sized by a dry run; conv_in reads the fp32 NCHW latent directly, conv_out writes fp32.
"""
import torch

from .. import fp32, packing
from ..ff.group_norm import GroupNorm
from ..native.b200.ops import b200
from ..runtime import F32, Act, Context, require_cuda, stream_ptr
from ..vision.conv2d import Conv2d
from ..vision.resnet import ResnetBlock
from .mid import Mid


class Decoder:
    def __init__(self):
        sz = [(128, 256), (256, 512), (512, 512), (512, 512)]
        self.conv_in = Conv2d(4, 512, kernel_size=[3, 3], padding=[1, 1])
        self.mid = Mid(512)
        arr = []
        for i, s in enumerate(sz):
            arr.append({"block": [ResnetBlock(s[1], s[0]), ResnetBlock(s[0], s[0]), ResnetBlock(s[0], s[0])]})
            if i != 0:
                arr[-1]['upsample'] = {"conv": Conv2d(s[0], s[0], kernel_size=[3, 3], padding=[1, 1])}
        self.up = arr
        self.norm_out = GroupNorm(32, 128)
        self.conv_out = Conv2d(128, 3, kernel_size=[3, 3], padding=[1, 1])
        self._engines = {}

    def __call__(self, x):
        """(B, 4, h, w) fp32 NCHW latent (already through post_quant_conv) -> (B, 3, 8h, 8w) fp32 NCHW."""
        require_cuda(x, "x")
        if fp32.enabled():
            return fp32.vae_decoder(self, x)
        eng = self._engine(tuple(x.shape))
        return eng.decode_nchw(x)

    def _engine(self, shape):
        from .. import get_quirks
        key = (shape, get_quirks(), torch.cuda.current_device())
        eng = self._engines.get(key)
        if eng is None:
            eng = DecoderEngine(self, shape[0], shape[2], shape[3], get_quirks())
            self._engines = {key: eng}      # one resident configuration (the arena is ~1.5 GB at 512^2)
        return eng

    # latent: fp32 NCHW (n, 4, H, W) device pointer; img: fp32 NHWC (n, 8H, 8W, 8) device pointer (channels 0..2 valid)
    def _run(self, ctx, latent_ptr, n, H, W, img_ptr):
        ar = ctx.arena
        ar.reset()
        x = ctx.new_act(n, H, W, 512, gn=True, gn_unit=16)      # conv_in emits the first GroupNorm's statistics too
        self.conv_in._run_smallcin(ctx, _F32View(latent_ptr, (n, 4, H, W)), n, x)
        h = ctx.new_act(n, H, W, 512, gn=True, gn_unit=16)
        self.mid._run(ctx, x, h)
        x = h
        for l in self.up[::-1]:
            for b in l['block']:
                out = ctx.new_act(x.n, x.h, x.w, b.out_channels, gn=True, gn_unit=b.out_channels // 32)
                x = b._run(ctx, x, out)
            if 'upsample' in l:
                up = ctx.new_act(x.n, 2 * x.h, 2 * x.w, x.c)
                ctx.upsample2x(x, up)
                out = ctx.new_act(x.n, 2 * x.h, 2 * x.w, x.c, gn=True, gn_unit=x.c // 32)
                x = l['upsample']['conv']._run(ctx, up, out)
        hn = ctx.new_act(x.n, x.h, x.w, x.c)
        self.norm_out._run(ctx, x, hn, silu=True)
        img = Act(img_ptr, n, x.h, x.w, 8, 8)
        self.conv_out._run(ctx, hn, img, flags=b200.TF_EPI_OUT_F32)
        return img


class _F32View:
    def __init__(self, ptr, shape):
        self._ptr, self.shape = ptr, shape

    def data_ptr(self):
        return self._ptr


class DecoderEngine:
    """Static buffers + arena for one (batch, latent H, latent W) decoder configuration."""

    def __init__(self, model, n, H, W, quirks):
        dev = torch.device("cuda", torch.cuda.current_device())
        b200.init(dev.index)
        self.model, self.n, self.H, self.W = model, n, H, W
        self.ctx = Context(dev, quirks)
        self.ctx.ensure_workspaces()
        self.latent = torch.zeros((n, 4, H, W), dtype=F32, device=dev)
        self.img = torch.zeros((n, 8 * H, 8 * W, 8), dtype=F32, device=dev)
        self.ctx.dry = self.ctx.arena.dry = True
        self._enqueue()
        self.ctx.dry = self.ctx.arena.dry = False
        self.arena_bytes = self.ctx.arena.peak
        self.ctx.arena.reserve(self.arena_bytes, dev)

    def _enqueue(self):
        with packing.domain("vae"):
            return self.model._run(self.ctx, self.latent.data_ptr(), self.n, self.H, self.W, self.img.data_ptr())

    def decode_nhwc_f32(self, latent):
        """-> the engine's (n, 8H, 8W, 8) fp32 NHWC buffer (channels 0..2 valid)."""
        self.latent.copy_(latent.reshape(self.latent.shape))
        self._enqueue()
        return self.img

    def decode_nchw(self, latent):
        self.decode_nhwc_f32(latent)
        out = torch.empty((self.n, 3, 8 * self.H, 8 * self.W), dtype=F32, device=self.latent.device)
        st = b200.tf_nhwc_f32_to_nchw_f32(self.img.data_ptr(), 8, out.data_ptr(), self.n, 3, 64 * self.H * self.W, stream_ptr())
        b200.check(st, "tf_nhwc_f32_to_nchw_f32")
        return out
