"""fp32 parity mode: the hot-path operators and blocks in plain fp32, in the reference's own layouts.

The reference computes in fp32 everywhere (SURVEY.md §8 banner); BASELINE.json's north star asks for per-op error
<= 1e-5 "in fp32 mode" and configs[0] is down-block-0 in fp32. `tinyfusers_b200.set_precision("fp32")` routes the
stand-alone operators (`conv_2d`, `Conv2d`, `Linear`, `group_norm`, `GroupNorm`, `layer_norm`, `LayerNorm`,
`scaled_dot_product_attention`, `GEGLU`, `FeedForward`, `Tensor.*`) and the blocks built from them (`ResBlock`,
`CrossAttention`, `BasicTransformerBlock`, `SpatialTransformer`, `Upsample`, `Downsample`, `UNetModel`) through the
kernels of csrc/tf_fp32.cu: CUDA-core FMA, IEEE expf/tanhf, NCHW images and (B,T,C) tokens, blocked K sums.

It is a correctness mode - the measured path is the fp16/tcgen05 one (bench.py never enters here) - but it is the
same boundary: every arithmetic operation is a C-ABI call into libtinyfusers_b200.so; torch only allocates, reshapes
and concatenates (the data movement `cp.concatenate` / `cp.reshape` do in the reference).
"""
import ctypes
import math

import torch

from .native.b200.ops import b200
from .runtime import require_cuda, stream_ptr

F32 = torch.float32
_PRECISION = "fp16"


def set_precision(mode: str):
    """"fp16" (default: tcgen05 kernels, fp16 operands / fp32 accumulate, <= 1e-2) or "fp32" (parity mode, <= 1e-5)."""
    global _PRECISION
    if mode not in ("fp16", "fp32"):
        raise ValueError(f"set_precision: unknown mode {mode!r} (fp16 | fp32)")
    _PRECISION = mode


def get_precision() -> str:
    return _PRECISION


def enabled() -> bool:
    return _PRECISION == "fp32"


def _prep(x, name="x"):
    require_cuda(x, name)
    b200.init(x.device.index if x.device.index is not None else torch.cuda.current_device())
    return x.to(F32).contiguous()


def _ptr(t):
    return None if t is None else t.data_ptr()


def _w(t, like):
    """A parameter as contiguous fp32 on the activation's device (parameters are fp32 masters already)."""
    return None if t is None or not isinstance(t, torch.Tensor) else t.to(device=like.device, dtype=F32).contiguous()


# ---- operators -----------------------------------------------------------------------------------------------------

def gemm(a2d, w, bias=None, residual=None, alpha=1.0, out=None):
    """a2d (M,K) @ w(N,K)^T (+bias) (+residual (M,N))  -> (M,N)   (reference: ff/linear.py:119-120)."""
    M, K = a2d.shape
    N = w.shape[0]
    out = torch.empty((M, N), dtype=F32, device=a2d.device) if out is None else out
    st = b200.tf_gemm_f32(a2d.data_ptr(), a2d.stride(0), w.data_ptr(), w.stride(0), 0, _ptr(bias), _ptr(residual), N,
                          out.data_ptr(), N, M, N, K, float(alpha), 0, 1, 1, None, stream_ptr())
    b200.check(st, "tf_gemm_f32")
    return out


def linear(x, weight, bias=None, residual=None):
    x = _prep(x)
    w, b = _w(weight, x), _w(bias, x)
    x2 = x.reshape(-1, x.shape[-1])
    r2 = None if residual is None else _prep(residual).reshape(-1, w.shape[0])
    return gemm(x2, w, b, r2).reshape(*x.shape[:-1], w.shape[0])


def conv2d(x, weight, bias=None, stride=1, padding=0, bias_img=None, residual=None, out_tokens=False):
    """NCHW x OIHW cross-correlation (+bias) (+per-image bias) (+residual)   (reference: vision/conv2d.py:9-28,55-59)."""
    x = _prep(x)
    w, b = _w(weight, x), _w(bias, x)
    NI, C, H, W = x.shape
    O, Cw, R, S = w.shape
    if Cw != C:
        raise RuntimeError(f"conv2d (fp32): input has {C} channels, weight expects {Cw}")
    Ho, Wo = (H + 2 * padding - R) // stride + 1, (W + 2 * padding - S) // stride + 1
    shape = (NI * Ho * Wo, O) if out_tokens else (NI, O, Ho, Wo)
    out = torch.empty(shape, dtype=F32, device=x.device)
    bi = None if bias_img is None else _prep(bias_img).reshape(NI, O)
    res = None if residual is None else _prep(residual)
    st = b200.tf_conv2d_nchw_f32(x.data_ptr(), w.data_ptr(), _ptr(b), _ptr(bi), _ptr(res), out.data_ptr(), NI, C, H, W, O,
                                 R, S, int(stride), int(padding), 1 if out_tokens else 0, stream_ptr())
    b200.check(st, "tf_conv2d_nchw_f32")
    return out


def _sq(v, what):
    v = [int(v), int(v)] if isinstance(v, int) else [int(t) for t in v]
    if v[0] != v[1]:
        raise RuntimeError(f"conv2d (fp32): non-square {what} {v} has no kernel")
    return v[0]


def conv_module(m, x, **kw):
    """A Conv2d module (vision/conv2d.py:48-59) on NCHW fp32."""
    if _sq(m.dilation, "dilation") != 1:
        raise RuntimeError("conv2d (fp32): dilation has no kernel")
    return conv2d(x, m.weight, m.bias, _sq(m.stride, "stride"), _sq(m.padding, "padding"), **kw)


def group_norm(x, num_groups, eps, weight=None, bias=None, silu=False):
    """reference: ff/group_norm.py:3-21 (+ the SiLU that follows it in ResBlock, resnet.py:10-11)."""
    x = _prep(x)
    N, C = x.shape[0], x.shape[1]
    HW = x.numel() // (N * C)
    out = torch.empty_like(x)
    st = b200.tf_groupnorm_nchw_f32(x.data_ptr(), _ptr(_w(weight, x)), _ptr(_w(bias, x)), out.data_ptr(), N, C, HW,
                                    int(num_groups), float(eps), 1 if silu else 0, stream_ptr())
    b200.check(st, "tf_groupnorm_nchw_f32")
    return out


def layer_norm(x, weight, bias, eps):
    """reference: ff/layer_norm.py:8-32 as cuDNN executes it (canonical, over the last dimension)."""
    x = _prep(x)
    C = x.shape[-1]
    out = torch.empty_like(x)
    st = b200.tf_layernorm_f32(x.data_ptr(), _ptr(_w(weight, x)), _ptr(_w(bias, x)), out.data_ptr(), x.numel() // C, C,
                               float(eps), stream_ptr())
    b200.check(st, "tf_layernorm_f32")
    return out


def unary(x, op):
    x = _prep(x)
    out = torch.empty_like(x)
    b200.check(b200.tf_unary_f32(x.data_ptr(), out.data_ptr(), x.numel(), int(op), stream_ptr()), "tf_unary_f32")
    return out


def _strides(*v):
    return (ctypes.c_longlong * 6)(*[int(t) for t in v])


def _attention(q, ldq, sq, k, ldk, sk, v, ldv, sv, B, NH, Tq, Tk, d, out, ldo, so, causal=False):
    """softmax(scale * Q K^T) V per (batch, head); operands addressed as base + b*s[0] + h*s[1] + t*ld + j.
    Scores are materialised in fp32 like the reference does (attention/sdpa.py:62-76)."""
    dev = q.device
    scale = float(torch.tensor(1.0 / math.sqrt(d), dtype=F32))   # cp.single scale, sdpa.py:62
    # bound the scores buffer: heads of one batch element at a time when B * NH * Tq * Tk would exceed ~2 GiB
    per_b = NH * Tq * Tk * 4
    bstep = max(1, min(B, (2 << 30) // max(per_b, 1)))
    scores = torch.empty((bstep * NH, Tq, Tk), dtype=F32, device=dev)
    for b0 in range(0, B, bstep):
        nb = min(bstep, B - b0)
        qp, kp, vp, op = (t.data_ptr() + 4 * b0 * s[0] for t, s in ((q, sq), (k, sk), (v, sv), (out, so)))
        st = b200.tf_gemm_f32(qp, ldq, kp, ldk, 0, None, None, 0, scores.data_ptr(), Tk, Tq, Tk, d, scale, 0, nb, NH,
                              _strides(sq[0], sq[1], sk[0], sk[1], NH * Tq * Tk, Tq * Tk), stream_ptr())
        b200.check(st, "tf_gemm_f32")
        b200.check(b200.tf_softmax_rows_f32(scores.data_ptr(), nb * NH * Tq, Tk, Tq if causal else 0, stream_ptr()),
                   "tf_softmax_rows_f32")
        st = b200.tf_gemm_f32(scores.data_ptr(), Tk, vp, ldv, 1, None, None, 0, op, ldo, Tq, d, Tk, 1.0, 0, nb, NH,
                              _strides(NH * Tq * Tk, Tq * Tk, sv[0], sv[1], so[0], so[1]), stream_ptr())
        b200.check(st, "tf_gemm_f32")
    return out


def scaled_dot_product_attention(q, k, v, causal=False):
    """(B,NH,Tq,HS) x (B,NH,Tk,HS) -> (B,NH,Tq,HS)   (reference: attention/sdpa.py:53-77); `causal` = the one mask the
    reference passes (CLIP's triu(-inf, k=1), recognised by the caller in attention/sdpa.py)."""
    q, k, v = _prep(q, "q"), _prep(k, "k"), _prep(v, "v")
    B, NH, Tq, d = q.shape
    Tk = k.shape[-2]
    out = torch.empty_like(q)
    return _attention(q, d, (NH * Tq * d, Tq * d), k, d, (NH * Tk * d, Tk * d), v, d, (NH * Tk * d, Tk * d), B, NH, Tq, Tk, d,
                      out, d, (NH * Tq * d, Tq * d), causal=causal)


def geglu(m, x):
    """reference: ff/nn.py:5-12."""
    x = _prep(x)
    y = linear(x, m.proj.weight, m.proj.bias)
    H = y.shape[-1] // 2
    y2 = y.reshape(-1, 2 * H)
    out = torch.empty((y2.shape[0], H), dtype=F32, device=x.device)
    b200.check(b200.tf_geglu_f32(y2.data_ptr(), 2 * H, out.data_ptr(), y2.shape[0], H, stream_ptr()), "tf_geglu_f32")
    return out.reshape(*x.shape[:-1], H)


# ---- blocks (same attribute walk as the fp16 classes; reference file:line per function) -------------------------------

def feed_forward(m, x, residual=None):
    """reference: ff/nn.py:14-23 (net = [GEGLU, identity, Linear])."""
    return linear(geglu(m.net[0], x), m.net[2].weight, m.net[2].bias, residual)


def cross_attention(m, x, context=None, residual=None, quirks=None):
    """reference: attention/attention.py:26-41, head-major reshape quirk at :39 (set_quirks)."""
    from . import get_quirks
    quirks = get_quirks() if quirks is None else quirks
    x = _prep(x)
    context = x if context is None else _prep(context, "context")
    B, T, C = x.shape
    Tk = context.shape[1]
    nh, d = m.num_heads, m.head_size
    q = linear(x, m.to_q.weight)
    k = linear(context, m.to_k.weight)
    v = linear(context, m.to_v.weight)
    I = nh * d
    o = torch.empty((B, T, I), dtype=F32, device=x.device)
    # quirk: (B,NH,T,HS) memory re-read as (B,T,NH*HS); canonical: heads interleaved back into the token rows
    so, ldo = ((T * I, T * d), d) if quirks else ((T * I, d), I)
    _attention(q, I, (T * I, d), k, I, (Tk * I, d), v, I, (Tk * I, d), B, nh, T, Tk, d, o, ldo, so)
    return linear(o, m.to_out[0].weight, m.to_out[0].bias, residual)


def basic_transformer_block(m, x, context=None):
    """reference: attention/attention.py:43-56."""
    ln = lambda n, t: layer_norm(t, n.weight, n.bias, float(torch.as_tensor(n.eps).reshape(-1)[0]))
    x = _prep(x)
    x = cross_attention(m.attn1, ln(m.norm1, x), None, residual=x)
    x = cross_attention(m.attn2, ln(m.norm2, x), context, residual=x)
    return feed_forward(m.ff, ln(m.norm3, x), residual=x)


def spatial_transformer(m, x, context=None):
    """reference: attention/attention.py:58-76. proj_in writes token rows, proj_out reads them and writes NCHW + x_in."""
    x = _prep(x)
    b, c, h, w = x.shape
    xn = group_norm(x, m.norm.num_groups, m.norm.eps, m.norm.weight, m.norm.bias)
    t = conv_module(m.proj_in, xn, out_tokens=True).reshape(b, h * w, -1)
    for blk in m.transformer_blocks:
        t = basic_transformer_block(blk, t, context)
    wo, bo = _w(m.proj_out.weight, x), _w(m.proj_out.bias, x)
    O = wo.shape[0]
    out = torch.empty((b, O, h, w), dtype=F32, device=x.device)
    t2 = t.reshape(b * h * w, -1)
    wo2 = wo.reshape(O, -1)
    st = b200.tf_gemm_f32(t2.data_ptr(), t2.stride(0), wo2.data_ptr(), wo2.stride(0), 0, _ptr(bo), x.data_ptr(), 0,
                          out.data_ptr(), 0, b * h * w, O, t2.shape[1], 1.0, h * w, 1, 1, None, stream_ptr())
    b200.check(st, "tf_gemm_f32")
    return out


def res_block(m, x, emb):
    """reference: vision/resnet.py:6-31."""
    x = _prep(x)
    gn1, conv1 = m.in_layers[0], m.in_layers[2]
    gn2, conv2 = m.out_layers[0], m.out_layers[3]
    h = group_norm(x, gn1.num_groups, gn1.eps, gn1.weight, gn1.bias, silu=True)
    lin = m.emb_layers[1]
    emb_out = linear(unary(_prep(emb, "emb").reshape(-1, lin.weight.shape[1]), 1), lin.weight, lin.bias)
    if emb_out.shape[0] != x.shape[0]:
        emb_out = emb_out.expand(x.shape[0], -1)
    h = conv_module(conv1, h, bias_img=emb_out)
    h = group_norm(h, gn2.num_groups, gn2.eps, gn2.weight, gn2.bias, silu=True)
    skip = m.skip_connection
    xs = conv_module(skip, x) if hasattr(skip, "weight") else x
    return conv_module(conv2, h, residual=xs)


def resnet_block(m, x):
    """reference: vision/resnet.py:33-45 (VAE)."""
    x = _prep(x)
    h = group_norm(x, m.norm1.num_groups, m.norm1.eps, m.norm1.weight, m.norm1.bias, silu=True)
    h = conv_module(m.conv1, h)
    h = group_norm(h, m.norm2.num_groups, m.norm2.eps, m.norm2.weight, m.norm2.bias, silu=True)
    xs = conv_module(m.nin_shortcut, x) if hasattr(m.nin_shortcut, "weight") else x
    return conv_module(m.conv2, h, residual=xs)


def upsample(m, x):
    """reference: vision/unet.py:78-84 (nearest x2 by broadcast + reshape, then the 3x3 conv)."""
    x = _prep(x)
    bs, c, py, px = x.shape
    x = x.reshape(bs, c, py, 1, px, 1).expand(bs, c, py, 2, px, 2).reshape(bs, c, py * 2, px * 2)
    return conv_module(m.conv, x)


def downsample(m, x):
    """reference: vision/unet.py:86-90."""
    return conv_module(m.op, x)


def _run_layer(layer, x, emb, context):
    name = type(layer).__name__
    if name == "ResBlock":
        return res_block(layer, x, emb)
    if name == "SpatialTransformer":
        return spatial_transformer(layer, x, context)
    if name == "Upsample":
        return upsample(layer, x)
    if name == "Downsample":
        return downsample(layer, x)
    if name == "Conv2d":
        return conv_module(layer, x)
    raise RuntimeError(f"fp32 UNet: no fp32 path for layer {name}")


def unet_forward(m, x, timesteps=None, context=None):
    """reference: vision/unet.py:51-76 - the block walk of the reference itself (no engine, no graph)."""
    from .vision.unet import timestep_embedding
    x = _prep(x)
    context = _prep(context, "context")
    t_emb = timestep_embedding(timesteps, 320)
    emb = linear(t_emb, m.time_embed[0].weight, m.time_embed[0].bias)
    emb = linear(unary(emb, 1), m.time_embed[2].weight, m.time_embed[2].bias)
    saved = []
    for block in m.input_blocks:
        for layer in block:
            x = _run_layer(layer, x, emb, context)
        saved.append(x)
    for layer in m.middle_block:
        x = _run_layer(layer, x, emb, context)
    for block in m.output_blocks:
        x = torch.cat((x, saved.pop()), dim=1)       # cp.concatenate, unet.py:72 (data movement only)
        for layer in block:
            x = _run_layer(layer, x, emb, context)
    gn, conv = m.out[0], m.out[2]
    x = group_norm(x, gn.num_groups, gn.eps, gn.weight, gn.bias, silu=True)
    return conv_module(conv, x)


def get_model_output(sd_model, unconditional_context, context, latent, timestep, guidance):
    """reference: variants/sd.py:27-46 - batch [uncond ; cond] through the UNet, e_t = u + g (c - u)."""
    latent = _prep(latent, "latent")
    n = latent.shape[0]
    lat2 = torch.cat((latent, latent), dim=0)
    ctx2 = torch.cat((_prep(unconditional_context, "unconditional_context"), _prep(context, "context")), dim=0)
    out = unet_forward(sd_model.model.diffusion_model, lat2, timestep, ctx2)
    u, c = out[0:n].contiguous(), out[n:2 * n].contiguous()
    e_t = torch.empty_like(u)
    g = float(torch.as_tensor(guidance).reshape(-1)[0])
    b200.check(b200.tf_cfg_combine_f32(u.data_ptr(), c.data_ptr(), g, e_t.data_ptr(), u.numel(), stream_ptr()),
               "tf_cfg_combine_f32")
    return e_t


# ---- rows next to the hot path: VAE decoder, CLIP text encoder --------------------------------------

def attn_block(m, x, quirks=None):
    """reference: attention/attention.py:10-24. quirks: the 4-D (B,C,H,W) q/k/v go straight into SDPA, i.e. NH = C,
    T = H, HS = W (parity note 3). Canonical (quirks off): one head over the H*W pixels, head dim C - built here (the
    fp16 attention kernel stops at head dim 256), so `set_quirks(False)` decodes real checkpoints in fp32 mode."""
    from . import get_quirks
    quirks = get_quirks() if quirks is None else quirks
    x = _prep(x)
    B, C, H, W = x.shape
    hn = group_norm(x, m.norm.num_groups, m.norm.eps, m.norm.weight, m.norm.bias)
    if quirks:
        q, k, v = (conv_module(c, hn) for c in (m.q, m.k, m.v))
        o = scaled_dot_product_attention(q, k, v)
        return conv_module(m.proj_out, o, residual=x)
    T = H * W
    q, k, v = (conv_module(c, hn, out_tokens=True) for c in (m.q, m.k, m.v))     # (B*T, C) token rows
    o = torch.empty_like(q)
    _attention(q, C, (T * C, 0), k, C, (T * C, 0), v, C, (T * C, 0), B, 1, T, T, C, o, C, (T * C, 0))
    wo, bo = _w(m.proj_out.weight, x), _w(m.proj_out.bias, x)
    wo2 = wo.reshape(wo.shape[0], -1)
    out = torch.empty_like(x)
    st = b200.tf_gemm_f32(o.data_ptr(), C, wo2.data_ptr(), wo2.stride(0), 0, _ptr(bo), x.data_ptr(), 0, out.data_ptr(), 0,
                          B * T, wo.shape[0], C, 1.0, T, 1, 1, None, stream_ptr())
    b200.check(st, "tf_gemm_f32")
    return out


def vae_mid(m, x):
    """reference: vae/mid.py:5-12."""
    return resnet_block(m.block_2, attn_block(m.attn_1, resnet_block(m.block_1, x)))


def vae_decoder(m, x):
    """reference: vae/decoder.py:22-34."""
    x = conv_module(m.conv_in, x)
    x = vae_mid(m.mid, x)
    for l in m.up[::-1]:
        for b in l["block"]:
            x = resnet_block(b, x)
        if "upsample" in l:
            bs, c, py, px = x.shape      # nearest x2 (decoder.py:30-31), data movement only
            x = x.reshape(bs, c, py, 1, px, 1).expand(bs, c, py, 2, px, 2).reshape(bs, c, py * 2, px * 2)
            x = conv_module(l["upsample"]["conv"], x)
    n = m.norm_out
    return conv_module(m.conv_out, group_norm(x, n.num_groups, n.eps, n.weight, n.bias, silu=True))


def decode(sd_model, x):
    """reference: variants/sd.py:48-54 (the hard-coded 512 generalised like the fp16 path): uint8 (H,W,3) | (B,H,W,3)."""
    fsm = sd_model.first_stage_model
    z = fsm.post_quant(_prep(x), 1 / 0.18215)
    img = vae_decoder(fsm.decoder, z)                         # (B, 3, H, W) fp32
    B, _, H, W = img.shape
    nhwc = img.permute(0, 2, 3, 1).contiguous()
    out = torch.empty((B, H, W, 3), dtype=torch.uint8, device=img.device)
    b200.check(b200.tf_image_to_u8(nhwc.data_ptr(), 3, out.data_ptr(), B * H * W, 3, stream_ptr()), "tf_image_to_u8")
    return out[0] if B == 1 else out


def clip_attention(m, x, residual=None):
    """reference: attention/attention.py:78-99 (12 heads x 64, biased projections, causal mask, canonical head merge)."""
    x = _prep(x)
    B, T, E = x.shape
    NH, D = m.num_heads, m.head_dim
    q, k, v = (linear(x, p.weight, p.bias) for p in (m.q_proj, m.k_proj, m.v_proj))
    o = torch.empty_like(q)
    s = (T * E, D)
    _attention(q, E, s, k, E, s, v, E, s, B, NH, T, T, D, o, E, s, causal=True)
    return linear(o, m.out_proj.weight, m.out_proj.bias, residual)


def clip_mlp(m, x, residual=None):
    """reference: ff/nn.py:25-34."""
    h = unary(linear(x, m.fc1.weight, m.fc1.bias), 3)
    return linear(h, m.fc2.weight, m.fc2.bias, residual)


def _ln(n, t):
    return layer_norm(t, n.weight, n.bias, float(torch.as_tensor(n.eps).reshape(-1)[0]))


def clip_encoder_layer(m, x):
    """reference: vae/encoder.py:53-64."""
    x = _prep(x)
    x = clip_attention(m.self_attn, _ln(m.layer_norm1, x), residual=x)
    return clip_mlp(m.mlp, _ln(m.layer_norm2, x), residual=x)


def clip_text_transformer(m, ids):
    """reference: vae/encoder.py:72-81; ids (B, T) int32 on the device."""
    B, T = ids.shape
    tok = m.embeddings.token_embedding.weight
    pos = m.embeddings.position_embedding.weight
    tok, pos = _w(tok, tok), _w(pos, tok)
    ids = ids.to(device=tok.device, dtype=torch.int32).contiguous()
    x = torch.empty((B, T, tok.shape[1]), dtype=F32, device=tok.device)
    st = b200.tf_embedding_f32(ids.data_ptr(), tok.data_ptr(), pos.data_ptr(), x.data_ptr(), B * T, T, tok.shape[1],
                               tok.shape[0], stream_ptr())
    b200.check(st, "tf_embedding_f32")
    for l in m.encoder.layers:
        x = clip_encoder_layer(l, x)
    return _ln(m.final_layer_norm, x)
