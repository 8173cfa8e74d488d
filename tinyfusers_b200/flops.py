"""Algorithmic work of one UNet forward (the numerators of bench.py's roofline figures).

Counted from the module tree exactly as SURVEY.md §8d does: 2*M*N*K per conv / linear,
4*B*NH*Tq*Tk*d per attention. Padding introduced by the kernels (head dim 40 -> 48, N tiles) is NOT
counted: these are the reference's FLOPs, not the machine's."""
from .attention.attention import SpatialTransformer
from .vision.conv2d import Conv2d
from .vision.resnet import ResBlock
from .vision.unet import Downsample, Upsample


def _conv(c, n, h, w):
    O, I, kh, kw = c.weight.shape
    s = c.stride[0]
    ho, wo = (h + 2 * c.padding[0] - kh) // s + 1, (w + 2 * c.padding[1] - kw) // s + 1
    return 2.0 * n * ho * wo * O * I * kh * kw, ho, wo


def unet_flops(unet, n, H, W, ctx_tokens=77):
    """-> dict(gemm=..., attention=..., total=...) in FLOPs for batch n at latent H x W."""
    gemm = attn = 0.0

    def lin(l, m):
        return 2.0 * m * l.weight.shape[0] * l.weight.shape[1]

    def res(rb, h, w):
        f = _conv(rb.in_layers[2], n, h, w)[0] + _conv(rb.out_layers[3], n, h, w)[0] + lin(rb.emb_layers[1], 1)
        if isinstance(rb.skip_connection, Conv2d):
            f += _conv(rb.skip_connection, n, h, w)[0]
        return f

    def st(s, h, w):
        nonlocal attn
        T = h * w
        blk = s.transformer_blocks[0]
        f = _conv(s.proj_in, n, h, w)[0] + _conv(s.proj_out, n, h, w)[0]
        for a, tk, m_kv in ((blk.attn1, T, n * T), (blk.attn2, ctx_tokens, n * ctx_tokens)):
            f += lin(a.to_q, n * T) + lin(a.to_k, m_kv) + lin(a.to_v, m_kv) + lin(a.to_out[0], n * T)
            attn += 4.0 * n * a.num_heads * T * tk * a.head_size
        f += lin(blk.ff.net[0].proj, n * T) + lin(blk.ff.net[2], n * T)
        return f

    gemm += lin(unet.time_embed[0], 1) + lin(unet.time_embed[2], 1)
    h, w = H, W
    for group in (unet.input_blocks, [unet.middle_block], unet.output_blocks):
        for blk in group:
            for layer in blk:
                if isinstance(layer, ResBlock):
                    gemm += res(layer, h, w)
                elif isinstance(layer, SpatialTransformer):
                    gemm += st(layer, h, w)
                elif isinstance(layer, Downsample):
                    f, h, w = _conv(layer.op, n, h, w)
                    gemm += f
                elif isinstance(layer, Upsample):
                    h, w = 2 * h, 2 * w
                    gemm += _conv(layer.conv, n, h, w)[0]
                elif isinstance(layer, Conv2d):
                    gemm += _conv(layer, n, h, w)[0]
    gemm += _conv(unet.out[2], n, h, w)[0]
    return {"gemm": gemm, "attention": attn, "total": gemm + attn}
