"""UNetModel / Upsample / Downsample / timestep_embedding with the reference's signatures
(reference: tinyfusers/vision/unet.py:9-97).

Fast path of one forward (batch 2B for CFG):
  * activations fp16 NHWC in a stack arena with call-sequence-deterministic addresses (graph capturable);
  * `cp.concatenate((x, saved_inputs.pop()), axis=1)` (unet.py:72) costs nothing: every input block writes its
    output straight into the second channel half of the buffer its output block will read, and the
    producer of `x` writes the first half (all kernels take a pixel stride);
  * the 22 ResBlock emb projections are one GEMV over their row-concatenated weights, folded into the
    first conv's bias; conv_in reads the fp32 NCHW latent directly and forms the CFG batch by index;
  * conv_out writes fp32 NHWC (channels padded to 16) that the fused CFG+DDIM kernel consumes.
"""
import os

import numpy as np
import torch

from .. import fp32, get_layernorm_strided, get_quirks, packing
from ..attention.attention import SpatialTransformer
from ..ff.group_norm import GroupNorm
from ..ff.linear import Linear
from ..native.b200.ops import b200
from ..runtime import F16, F32, Act, Context, act_to_nchw, nchw_to_act, new_act_tensor, require_cuda, standalone_context, stream_ptr
from ..storage.tensor import Tensor
from .conv2d import Conv2d
from .resnet import ResBlock


FUSED_UPSAMPLE = os.environ.get("TINYFUSERS_B200_FUSED_UPSAMPLE", "1") != "0"
NVTX = os.environ.get("TINYFUSERS_B200_NVTX", "0") == "1"


def _pow2_tiling(h, w):
    """a 128-pixel tile TW x TH x TN (powers of two) with TW | w, TH | h and >= 32 pixels of one image exists"""
    for tw in (128, 64, 32, 16, 8, 4, 2, 1):
        for th in (128, 64, 32, 16, 8, 4, 2, 1):
            if tw * th <= 128 and tw * th >= 32 and w % tw == 0 and h % th == 0:
                return True
    return False


class Upsample:
    def __init__(self, channels):
        self.conv = Conv2d(channels, channels, kernel_size=[3, 3], padding=[1, 1])

    def __call__(self, x):
        require_cuda(x, "x")
        if fp32.enabled():
            return fp32.upsample(self, x)
        ctx = standalone_context()
        ctx.arena.reset()
        a = nchw_to_act(x, c_pad_to=64)
        out = new_act_tensor(a.n, a.h * 2, a.w * 2, self.conv.weight.shape[0], device=x.device)
        self._run(ctx, a, out)
        return act_to_nchw(out, self.conv.weight.shape[0])

    def _packed_up2x(self):
        def build():
            w4 = packing.conv_up2x_weight(self.conv.weight, 64, 8)
            b = packing.f32(self.conv.bias)
            if b is not None and b.shape[0] != w4.shape[1]:
                b = torch.nn.functional.pad(b, (0, w4.shape[1] - b.shape[0]))
            return w4, b
        return packing.cached(self, "up2x", (self.conv.weight, self.conv.bias), build)

    def _run(self, ctx, x, out):
        """Upsampling folded into the convolution (tf_conv2d_up2x_nhwc_f16: four 2 x 2 phase convolutions of the original image,
        4/9 of the multiply-adds, no 4x tensor) wherever the input tiles exactly; the two-launch path otherwise."""
        if FUSED_UPSAMPLE and (x.h * x.w) % 32 == 0 and x.c % 64 == 0 and _pow2_tiling(x.h, x.w):
            w4, b = self._packed_up2x()
            cout_p = w4.shape[1]
            gn = None
            if out.gn is not None and len(out.gn) == 1 and out.gn[0][0] == 0 and out.gn[0][1] == cout_p:
                gn = (out.gn[0][2], out.gn[0][3])
            if gn is not None or out.gn is None:
                ctx.conv_up2x(x, w4.data_ptr(), cout_p, cout_p, out, bias=b.data_ptr() if b is not None else None, gn=gn)
                return out
        mark = ctx.arena.mark()
        up = ctx.new_act(x.n, x.h * 2, x.w * 2, x.c)
        ctx.upsample2x(x, up)
        self.conv._run(ctx, up, out)
        ctx.arena.release(mark)
        return out


class Downsample:
    def __init__(self, channels):
        self.op = Conv2d(channels, channels, stride=[2, 2], kernel_size=[3, 3], padding=[1, 1])

    def __call__(self, x):
        return self.op(x)

    def _run(self, ctx, x, out):
        return self.op._run(ctx, x, out)


def timestep_embedding(timesteps, dim, max_period=10000):
    """(1, dim) fp32 [cos | sin] embedding of ONE timestep (reference: unet.py:92-97)."""
    dev = torch.device("cuda", torch.cuda.current_device())
    b200.init(dev.index)
    if isinstance(timesteps, torch.Tensor):
        t = timesteps.reshape(-1)[:1].to(device=dev, dtype=F32)
    else:
        t = torch.tensor([float(np.asarray(timesteps).reshape(-1)[0])], dtype=F32, device=dev)
    out = torch.empty((1, dim), dtype=F32, device=dev)
    st = b200.tf_timestep_embedding_f32(t.data_ptr(), None, dim, float(max_period), out.data_ptr(), stream_ptr())
    b200.check(st, "tf_timestep_embedding_f32")
    return out


class UNetModel:
    def __init__(self):
        self.time_embed = [Linear(320, 1280), Tensor.silu, Linear(1280, 1280), ]
        self.input_blocks = [
            [Conv2d(4, 320, kernel_size=[3, 3], padding=[1, 1])],
            [ResBlock(320, 1280, 320), SpatialTransformer(320, 768, 8, 40)],
            [ResBlock(320, 1280, 320), SpatialTransformer(320, 768, 8, 40)],
            [Downsample(320)],
            [ResBlock(320, 1280, 640), SpatialTransformer(640, 768, 8, 80)],
            [ResBlock(640, 1280, 640), SpatialTransformer(640, 768, 8, 80)],
            [Downsample(640)],
            [ResBlock(640, 1280, 1280), SpatialTransformer(1280, 768, 8, 160)],
            [ResBlock(1280, 1280, 1280), SpatialTransformer(1280, 768, 8, 160)],
            [Downsample(1280)],
            [ResBlock(1280, 1280, 1280)],
            [ResBlock(1280, 1280, 1280)]
        ]
        self.middle_block = [
            ResBlock(1280, 1280, 1280),
            SpatialTransformer(1280, 768, 8, 160),
            ResBlock(1280, 1280, 1280)
        ]
        self.output_blocks = [
            [ResBlock(2560, 1280, 1280)],
            [ResBlock(2560, 1280, 1280)],
            [ResBlock(2560, 1280, 1280), Upsample(1280)],
            [ResBlock(2560, 1280, 1280), SpatialTransformer(1280, 768, 8, 160)],
            [ResBlock(2560, 1280, 1280), SpatialTransformer(1280, 768, 8, 160)],
            [ResBlock(1920, 1280, 1280), SpatialTransformer(1280, 768, 8, 160), Upsample(1280)],
            [ResBlock(1920, 1280, 640), SpatialTransformer(640, 768, 8, 80)],
            [ResBlock(1280, 1280, 640), SpatialTransformer(640, 768, 8, 80)],
            [ResBlock(960, 1280, 640), SpatialTransformer(640, 768, 8, 80), Upsample(640)],
            [ResBlock(960, 1280, 320), SpatialTransformer(320, 768, 8, 40)],
            [ResBlock(640, 1280, 320), SpatialTransformer(320, 768, 8, 40)],
            [ResBlock(640, 1280, 320), SpatialTransformer(320, 768, 8, 40)],
        ]
        self.out = [
            GroupNorm(32, 320),
            Tensor.silu,
            Conv2d(320, 4, kernel_size=[3, 3], padding=[1, 1])
        ]
        self._engines = {}

    # ---- reference-signature call: x (NB,4,H,W), timesteps 1 element, context (NB,77,768) -> (NB,4,H,W) fp32 ----
    def __call__(self, x, timesteps=None, context=None):
        require_cuda(x, "x")
        if fp32.enabled():
            return fp32.unet_forward(self, x, timesteps, context)
        NB, _, H, W = x.shape
        eng = self.engine(NB, H, W, n_src=NB, ctx_tokens=context.shape[1])
        return eng.forward_nchw(x, timesteps, context)

    def engine(self, n_images, H, W, n_src=None, ctx_tokens=77):
        n_src = n_images if n_src is None else n_src
        key = (torch.cuda.current_device(), n_images, H, W, n_src, ctx_tokens, get_quirks(), get_layernorm_strided())
        eng = self._engines.get(key)
        if eng is None:
            eng = UNetEngine(self, n_images, H, W, n_src, ctx_tokens, get_quirks(), get_layernorm_strided())
            self._engines[key] = eng
        return eng

    def _res_blocks(self):
        blocks = []
        for group in (self.input_blocks, [self.middle_block], self.output_blocks):
            for b in group:
                for layer in b:
                    if isinstance(layer, ResBlock):
                        blocks.append(layer)
        return blocks

    def _emb_pack(self):
        """Row-concatenated emb_layers weights / biases / conv1 biases of all ResBlocks (one GEMV)."""
        rbs = self._res_blocks()
        tensors = []
        for rb in rbs:
            tensors += [rb.emb_layers[1].weight, rb.emb_layers[1].bias, rb.in_layers[2].bias]
        def build():
            w = torch.cat([rb.emb_layers[1]._packed()[0][:rb.out_channels] for rb in rbs], dim=0).contiguous()
            b = torch.cat([packing.f32(rb.emb_layers[1].bias) for rb in rbs]).contiguous()
            cb = torch.cat([packing.f32(rb.in_layers[2].bias) for rb in rbs]).contiguous()
            offs, o = {}, 0
            for rb in rbs:
                offs[id(rb)] = o
                o += rb.out_channels
            return w, b, cb, offs, o
        return packing.cached(self, "emb", tensors, build)

    def _cross_attns(self):
        out = []
        for group in (self.input_blocks, [self.middle_block], self.output_blocks):
            for b in group:
                for layer in b:
                    if isinstance(layer, SpatialTransformer):
                        out += [blk.attn2 for blk in layer.transformer_blocks]
        return out

    def _ctx_kv_pack(self):
        """Row-concatenated [to_k ; to_v] weights (head-padded) of every cross-attention: the prompt context is the same
        for all 16 transformer blocks, so their K and V projections are ONE GEMM per step, not 32."""
        atts = self._cross_attns()
        tensors = []
        for a in atts:
            tensors += [a.to_k.weight, a.to_v.weight]
        def build():
            wkv = torch.cat([a._packed()[3] for a in atts], dim=0).contiguous()
            offs, o = {}, 0
            for a in atts:
                offs[id(a)] = (o, a._packed()[1].shape[0])     # (first row of this block's K rows, number of K rows)
                o += a._packed()[3].shape[0]
            return wkv, offs, o
        return packing.cached(self, "ctxkv", tensors, build)

    # ---- fast path ------------------------------------------------------------------------------
    def _run(self, ctx, latent_ptr, n_src, n, H, W, t_ptr, idx_ptr, context_ptr, ctx_tokens, eps_ptr):
        """Enqueue one forward. latent: fp32 NCHW (n_src,4,H,W) (image i of the batch reads i % n_src);
        context: fp32 (n, ctx_tokens, 768); eps out: fp32 NHWC (n, H, W, 16), channels 0..3 valid."""
        ar = ctx.arena
        ar.reset()
        S = stream_ptr() if not ctx.dry else None
        misc = not ctx.skip("misc")
        # --- time embedding MLP and all ResBlock emb biases (fp32, M = 1) ---
        temb, e1, emb = ctx.new_f32(320), ctx.new_f32(1280), ctx.new_f32(1280)
        wemb, bemb, cbias, offs, total = self._emb_pack()
        allb = ctx.new_f32(total)
        w0, b0 = self.time_embed[0]._packed()
        w2, b2 = self.time_embed[2]._packed()
        if misc:
            b200.check(b200.tf_timestep_embedding_f32(t_ptr, idx_ptr, 320, 10000.0, temb, S), "tf_timestep_embedding_f32")
            b200.check(b200.tf_gemv_f16w(temb, w0.data_ptr(), b0.data_ptr(), None, e1, 1280, 320, 0, S), "tf_gemv_f16w")
            b200.check(b200.tf_gemv_f16w(e1, w2.data_ptr(), b2.data_ptr(), None, emb, 1280, 1280, 1, S), "tf_gemv_f16w")
            b200.check(b200.tf_gemv_f16w(emb, wemb.data_ptr(), bemb.data_ptr(), cbias.data_ptr(), allb, total, 1280, 1, S),
                       "tf_gemv_f16w")
        emb_bias = lambda rb: allb + 4 * offs[id(rb)]
        # --- prompt context -> fp16, tokens zero-padded to a multiple of 8 ---
        tkp = (ctx_tokens + 7) // 8 * 8
        cact = ctx.new_act(n, tkp, 1, 768)
        cact.valid = ctx_tokens
        if misc:
            b200.check(b200.tf_pad_tokens_f32_to_f16(context_ptr, cact.ptr, n, ctx_tokens, tkp, 768, S),
                       "tf_pad_tokens_f32_to_f16")
        # --- K and V of every cross-attention in one launch (same context for all blocks) ---
        wkv_all, kv_offs, kv_total = self._ctx_kv_pack()
        Mc = n * tkp
        kvall = ar.alloc(2 * Mc * kv_total)
        ctx.gemm(cact.ptr, 768, Mc, 768, wkv_all.data_ptr(), kv_total, kvall, kv_total)
        ctx.ctx_kv = {key: (kvall + 2 * off, kv_total, kvall + 2 * (off + nk), kv_total) for key, (off, nk) in kv_offs.items()}
        # --- plan the skip/concat buffers: output block j reads [x | saved[11-j]] ---
        res = [(H, W)]
        in_ch, in_hw = [], []
        h, w = H, W
        for blk in self.input_blocks:
            first = blk[0]
            if isinstance(first, Downsample):
                h, w = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
                c = first.op.weight.shape[0]
            elif isinstance(first, Conv2d):
                c = first.weight.shape[0]
            else:
                c = first.out_channels
            in_ch.append(c)
            in_hw.append((h, w))
        nb = len(self.output_blocks)
        cats = []
        for j, blk in enumerate(self.output_blocks):
            cat_c = blk[0].channels
            hh, ww = in_hw[nb - 1 - j]
            cats.append(ctx.new_act(n, hh, ww, cat_c))
        # persistent slice views; each gets its own statistics buffer, the concatenation lists both parts
        svs = [cats[nb - 1 - i].channels(cats[nb - 1 - i].c - in_ch[i], cats[nb - 1 - i].c) for i in range(nb)]
        xvs = [cats[j].channels(0, cats[j].c - in_ch[nb - 1 - j]) for j in range(nb)]
        for i in range(nb):
            ctx.attach_gn(svs[i])       # svs[0] is written by conv_in, which emits its statistics too
        for j in range(nb):
            ctx.attach_gn(xvs[j])
            xp, sp = xvs[j].gn, svs[nb - 1 - j].gn
            if xp is not None and sp is not None:
                c1 = xvs[j].c
                cats[j].gn = [xp[0], (c1, cats[j].c, sp[0][2], sp[0][3])]
        skip_view = lambda i: svs[i]
        x_view = lambda j: xvs[j]

        def run_layer(layer, x, final_out, last):
            if isinstance(layer, ResBlock):
                out = final_out if last else ctx.new_act(x.n, x.h, x.w, layer.out_channels, gn=True)
                return layer._run(ctx, x, emb_bias(layer), out)
            if isinstance(layer, SpatialTransformer):
                out = final_out if last else ctx.new_act(x.n, x.h, x.w, x.c, gn=True)
                return layer._run(ctx, x, cact, out)
            if isinstance(layer, Upsample):
                out = final_out if last else ctx.new_act(x.n, x.h * 2, x.w * 2, x.c)
                return layer._run(ctx, x, out)
            if isinstance(layer, Downsample):
                return layer._run(ctx, x, final_out)
            raise RuntimeError(f"unexpected layer {type(layer)}")

        def run_block(layers, x, final_out, tag=""):
            for li, layer in enumerate(layers):
                last = li == len(layers) - 1
                if NVTX:    # TINYFUSERS_B200_NVTX=1: one NVTX range per layer of an eagerly enqueued step (profilers group by it)
                    torch.cuda.nvtx.range_push(f"{tag}.{li}.{type(layer).__name__}")
                    try:
                        x = run_layer(layer, x, final_out, last)
                    finally:
                        torch.cuda.nvtx.range_pop()
                else:
                    x = run_layer(layer, x, final_out, last)
            return x

        # --- input blocks ---
        x = None
        for i, blk in enumerate(self.input_blocks):
            dst = skip_view(i)
            if i == 0:
                blk[0]._run_smallcin(ctx, _LatentView(latent_ptr, n_src, 4, H, W), n, dst)
                x = dst
            else:
                mark = ar.mark()
                x = run_block(blk, x, dst, f"input_blocks.{i}")
                ar.release(mark)
        # --- middle ---
        mark = ar.mark()
        x = run_block(self.middle_block, x, x_view(0), "middle_block")
        ar.release(mark)
        # --- output blocks ---
        for j, blk in enumerate(self.output_blocks):
            if j + 1 < nb:
                dst = x_view(j + 1)
            else:
                dst = ctx.new_act(n, H, W, blk[0].out_channels, gn=True)
            mark = ar.mark()
            x = run_block(blk, cats[j], dst, f"output_blocks.{j}")
            ar.release(mark)
        # --- out: GroupNorm + SiLU + conv 320 -> 4 (padded to 16, fp32) ---
        hn = ctx.new_act(n, H, W, x.c)
        self.out[0]._run(ctx, x, hn, silu=True)
        eps = Act(eps_ptr, n, H, W, 16, 16)
        self.out[2]._run(ctx, hn, eps, flags=b200.TF_EPI_OUT_F32)
        ctx.ctx_kv = None
        return eps


class _LatentView:
    """fp32 NCHW latent described by pointer + shape (what Conv2d._run_smallcin needs from a tensor)."""

    def __init__(self, ptr, n, c, h, w):
        self._ptr, self.shape = ptr, (n, c, h, w)

    def data_ptr(self):
        return self._ptr


class UNetEngine:
    """Static buffers + arena + (optionally) a captured CUDA graph for one (batch, H, W) configuration."""

    def __init__(self, model, n, H, W, n_src, ctx_tokens, quirks, ln_strided=False):
        dev = torch.device("cuda", torch.cuda.current_device())
        b200.init(dev.index)
        self.model, self.n, self.H, self.W, self.n_src, self.ctx_tokens = model, n, H, W, n_src, ctx_tokens
        self.ctx = Context(dev, quirks, ln_strided)
        self.ctx.ensure_workspaces()
        self.latent = torch.zeros((n_src, 4, H, W), dtype=F32, device=dev)
        self.context = torch.zeros((n, ctx_tokens, 768), dtype=F32, device=dev)
        self.t = torch.zeros(1, dtype=F32, device=dev)
        self.eps = torch.zeros((n, H, W, 16), dtype=F32, device=dev)
        self.out_nchw = torch.zeros((n, 4, H, W), dtype=F32, device=dev)
        # measure the arena with a dry run, then allocate it once
        self.ctx.dry = self.ctx.arena.dry = True
        self._enqueue()
        self.ctx.dry = self.ctx.arena.dry = False
        self.arena_bytes = self.ctx.arena.peak
        self.ctx.arena.reserve(self.arena_bytes, dev)
        self.graph = None

    def _enqueue(self, t_ptr=None, idx_ptr=None):
        with packing.domain("unet"):     # weights packed here only invalidate graphs that hold UNet weights
            return self.model._run(self.ctx, self.latent.data_ptr(), self.n_src, self.n, self.H, self.W,
                                   self.t.data_ptr() if t_ptr is None else t_ptr, idx_ptr, self.context.data_ptr(),
                                   self.ctx_tokens, self.eps.data_ptr())

    def forward_nchw(self, x, timesteps, context):
        self.latent.copy_(x.reshape(self.latent.shape))
        self.context.copy_(context.reshape(self.context.shape))
        if isinstance(timesteps, torch.Tensor):
            self.t.copy_(timesteps.reshape(-1)[:1])
        else:
            self.t.fill_(float(np.asarray(timesteps).reshape(-1)[0]))
        self._enqueue()
        st = b200.tf_nhwc_f32_to_nchw_f32(self.eps.data_ptr(), 16, self.out_nchw.data_ptr(), self.n, 4, self.H * self.W,
                                          stream_ptr())
        b200.check(st, "tf_nhwc_f32_to_nchw_f32")
        return self.out_nchw.clone()
