"""ORACLE — CPU restatement of the tinyfusers denoising hot path (TEST INFRASTRUCTURE ONLY).

This module restates, function by function, the arithmetic of the reference's SD1.x UNet step
(`/root/reference`, Fatlonder/tinyfusers) in plain torch-CPU / numpy. It exists to *check* the CUDA
path; nothing in the product (`tinyfusers_b200/`) imports it. Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s cpu_baseline / `--impl reference` legs may use it.

Parity pinning:
  * every function below is checked in tests/test_oracle_golden.py against golden vectors produced by
    running the reference's OWN Python (imported from /root/reference with numpy standing in for CuPy;
    script: oracle/make_golden.py, fixtures: tests/golden/*.npz);
  * the three ops whose arithmetic lives in cuDNN / cuBLAS / a CUDA kernel (conv2d, layer_norm,
    softmax) are additionally pinned the way the reference's own tests pin them — against
    torch.nn.functional (tests/conv2d.py:27-33, tests/layer_norm.py:33-41, tests/sdpa.py:97-100) — and
    LayerNorm against a real cuDNN run of the reference's graph (oracle/cudnn_probe.py, GPU box).

All functions take/return torch CPU tensors in the REFERENCE layouts: NCHW images, (B, T, C) tokens,
OIHW conv weights, (out, in) linear weights. `dtype` of the inputs decides the precision (fp32 like the
reference, or fp64 for a tighter yardstick).

`quirks=True` reproduces the reference's CrossAttention head-major reshape (SURVEY.md §8 parity note 2);
`quirks=False` is canonical Stable Diffusion. `ln_strided` (default False) selects the literal reading of the
LayerNorm stride declaration, which real cuDNN rejects at batch > 1 — see layer_norm().
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# ------------------------------------------------------------------------------------------------
# activations — reference: tinyfusers/storage/tensor.py:64-86
# ------------------------------------------------------------------------------------------------


def sigmoid(x):
    # tensor.py:65-66   1 / (1 + exp(-x))
    return 1 / (1 + torch.exp(-x))


def silu(x):
    # tensor.py:68-70   x * sigmoid(x)   (swish, tensor.py:84-86, is the same function)
    return x * sigmoid(x)


def quick_gelu(x):
    # tensor.py:76-78
    return x * sigmoid(x * 1.702)


def gelu(x):
    # tensor.py:80-82   tanh approximation
    return 0.5 * x * (1 + torch.tanh(x * 0.7978845608 * (1 + 0.044715 * x * x)))


# ------------------------------------------------------------------------------------------------
# operators
# ------------------------------------------------------------------------------------------------


def linear(x, weight, bias=None):
    # ff/linear.py:119-120   cp.dot(x, W.T) (+ bias)
    y = x @ weight.t()
    return y + bias if bias is not None else y


def conv2d(x, weight, bias=None, stride=(1, 1), padding=(0, 0), dilation=(1, 1)):
    # vision/conv2d.py:9-28 (cuDNN conv_fprop == cross-correlation; NHWC result re-read as NCHW at :27)
    # vision/conv2d.py:55-59 (bias broadcast add). The reference's own test pins this op to
    # torch.nn.functional.conv2d (tests/conv2d.py:27-33), which is what is used here.
    return F.conv2d(x, weight, bias, stride=tuple(stride), padding=tuple(padding), dilation=tuple(dilation))


def group_norm(x, num_groups, eps):
    # ff/group_norm.py:3-11 — literal two-pass, biased variance, no affine
    N, C, H, W = x.shape
    xg = x.reshape(N, num_groups, -1)
    mean = xg.mean(dim=-1, keepdim=True)
    yn = xg - mean
    yvar = torch.sqrt((yn * yn).mean(dim=-1, keepdim=True) + eps)
    return (yn * (1 / yvar)).reshape(N, C, H, W)


def group_norm_affine(x, num_groups, weight, bias, eps=1e-5):
    # ff/group_norm.py:18-21
    o = group_norm(x, num_groups, eps)
    shape = (1, -1) + (1,) * (x.dim() - 2)
    return o * weight.reshape(shape) + bias.reshape(shape)


def layer_norm(x, weight, bias, eps=1e-5, ln_strided=False):
    """ff/layer_norm.py:34-49 -> :8-32: LayerNorm over the last dimension (cuDNN inference layernorm graph).

    The reference hands cuDNN a contiguous (1,B,T,C) buffer but declares the strides
    [B*T*C, 1, B*C, B] (layer_norm.py:10). For B == 1 that IS the contiguous layout and the result is the
    canonical LayerNorm over C — confirmed against real cuDNN 9.x on the B200 box (oracle/cudnn_probe.py,
    tests/golden/cudnn_layernorm_probe.json, rel err 2e-7). For B > 1 the same probe shows cuDNN REJECTS the
    descriptor (CUDNN_STATUS_NOT_SUPPORTED at finalize): the reference has no result at CFG batch 2 on this
    cuDNN. The default here is therefore the semantics cuDNN does execute (canonical, per batch element).
    `ln_strided=True` is the literal reading of the declared strides (SURVEY.md §8 parity note 1): logical
    element (b,t,c) at memory offset b + t*B*C + c*B, i.e. view memory as (T, C, B) and normalise over axis 1
    — kept as an option of the kernel and tested, not the default.
    """
    B, T, C = x.shape
    if not ln_strided or B == 1:
        mean = x.mean(dim=-1, keepdim=True)
        var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
        return (x - mean) / torch.sqrt(var + eps) * weight + bias
    mem = x.contiguous().reshape(T, C, B)
    mean = mem.mean(dim=1, keepdim=True)
    var = ((mem - mean) ** 2).mean(dim=1, keepdim=True)
    y = (mem - mean) / torch.sqrt(var + eps) * weight.reshape(1, C, 1) + bias.reshape(1, C, 1)
    return y.reshape(B, T, C)


def softmax_rows(x):
    # native/cuda/softmax.cu:24-112 — per row: max, exp(x - max), sum, divide
    m = x.max(dim=-1, keepdim=True).values
    e = torch.exp(x - m)
    return e / e.sum(dim=-1, keepdim=True)


def scaled_dot_product_attention(q, k, v, attn_mask=None):
    # attention/sdpa.py:53-77 — scale * (q @ k^T) [+ mask] -> row softmax -> @ v
    HS = q.shape[-1]
    scale = torch.tensor(1.0 / math.sqrt(HS), dtype=torch.float32).to(device=q.device, dtype=q.dtype)  # cp.single, sdpa.py:62
    preatt = scale * (q @ k.transpose(-1, -2))
    if attn_mask is not None:
        if attn_mask.dtype == torch.bool:
            preatt = preatt + torch.where(attn_mask == 0, -float("inf"), 0.0).to(q.dtype)  # sdpa.py:68
        else:
            preatt = preatt + attn_mask
    att = softmax_rows(preatt)
    return att @ v


def timestep_embedding(timesteps, dim, max_period=10000):
    # vision/unet.py:92-97. `timesteps` is an int64 array of ONE element (example/sd1.py:73); int64 *
    # float promotes to float64 in CuPy, so the angles are formed in fp64 and only the result is fp32.
    half = dim // 2
    t = np.asarray(timesteps, dtype=np.float64).reshape(-1)
    freqs = np.exp(-np.log(float(max_period)) * np.arange(half, dtype=np.float32).astype(np.float64) / half)
    args = t * freqs
    out = np.concatenate((np.cos(args), np.sin(args))).reshape(1, -1).astype(np.float32)
    return torch.from_numpy(out)


def get_alphas_cumprod(beta_start=0.00085, beta_end=0.0120, n_training_steps=1000):
    # variants/sd.py:61-65 (fp32 throughout)
    betas = np.linspace(beta_start ** 0.5, beta_end ** 0.5, n_training_steps, dtype=np.float32) ** 2
    alphas = (1.0 - betas).astype(np.float32)
    return torch.from_numpy(np.cumprod(alphas, axis=0).astype(np.float32))


# ------------------------------------------------------------------------------------------------
# blocks. `sd` is a flat dict  "<prefix>.<attr path>.weight|bias" -> tensor, the reference's
# checkpoint naming (storage/state.py:4-23).
# ------------------------------------------------------------------------------------------------


def _w(sd, name):
    return sd[name]


def _b(sd, name):
    return sd.get(name)


def geglu(sd, p, x):
    # ff/nn.py:5-12: proj -> split(2, axis=-1): first half = value, second half = gate
    y = linear(x, _w(sd, p + ".proj.weight"), _b(sd, p + ".proj.bias"))
    a, gate = y.chunk(2, dim=-1)
    return a * gelu(gate)


def feed_forward(sd, p, x):
    # ff/nn.py:14-23: net = [GEGLU, identity, Linear]
    h = geglu(sd, p + ".net.0", x)
    return linear(h, _w(sd, p + ".net.2.weight"), _b(sd, p + ".net.2.bias"))


def cross_attention(sd, p, x, context, n_heads, d_head, quirks=True):
    # attention/attention.py:26-41
    context = x if context is None else context
    q = linear(x, _w(sd, p + ".to_q.weight"))
    k = linear(context, _w(sd, p + ".to_k.weight"))
    v = linear(context, _w(sd, p + ".to_v.weight"))
    B = x.shape[0]
    q, k, v = [y.reshape(B, -1, n_heads, d_head).permute(0, 2, 1, 3) for y in (q, k, v)]
    o = scaled_dot_product_attention(q, k, v)  # (B, NH, T, HS)
    if quirks:
        # attention.py:39 reshapes (B,NH,T,HS) straight to (B,T,NH*HS) WITHOUT moving heads back
        o = o.contiguous().reshape(B, -1, n_heads * d_head)
    else:
        o = o.permute(0, 2, 1, 3).reshape(B, -1, n_heads * d_head)
    return linear(o, _w(sd, p + ".to_out.0.weight"), _b(sd, p + ".to_out.0.bias"))


def basic_transformer_block(sd, p, x, context, n_heads, d_head, quirks=True, ln_strided=False):
    # attention/attention.py:43-56
    ln = lambda name, t: layer_norm(t, _w(sd, f"{p}.{name}.weight"), _b(sd, f"{p}.{name}.bias"), 1e-5, ln_strided)
    x = cross_attention(sd, p + ".attn1", ln("norm1", x), None, n_heads, d_head, quirks) + x
    x = cross_attention(sd, p + ".attn2", ln("norm2", x), context, n_heads, d_head, quirks) + x
    x = feed_forward(sd, p + ".ff", ln("norm3", x)) + x
    return x


def spatial_transformer(sd, p, x, context, n_heads, d_head, quirks=True, ln_strided=False):
    # attention/attention.py:58-76
    b, c, h, w = x.shape
    x_in = x
    x = group_norm_affine(x, 32, _w(sd, p + ".norm.weight"), _b(sd, p + ".norm.bias"), 1e-5)
    x = conv2d(x, _w(sd, p + ".proj_in.weight"), _b(sd, p + ".proj_in.bias"))
    x = x.reshape(b, c, h * w).permute(0, 2, 1)
    x = basic_transformer_block(sd, p + ".transformer_blocks.0", x, context, n_heads, d_head, quirks, ln_strided)
    x = x.permute(0, 2, 1).reshape(b, c, h, w)
    return conv2d(x, _w(sd, p + ".proj_out.weight"), _b(sd, p + ".proj_out.bias")) + x_in


def res_block(sd, p, x, emb):
    # vision/resnet.py:6-31
    h = group_norm_affine(x, 32, _w(sd, p + ".in_layers.0.weight"), _b(sd, p + ".in_layers.0.bias"))
    h = conv2d(silu(h), _w(sd, p + ".in_layers.2.weight"), _b(sd, p + ".in_layers.2.bias"), padding=(1, 1))
    emb_out = linear(silu(emb), _w(sd, p + ".emb_layers.1.weight"), _b(sd, p + ".emb_layers.1.bias"))
    h = h + emb_out.reshape(*emb_out.shape, 1, 1)
    h = group_norm_affine(h, 32, _w(sd, p + ".out_layers.0.weight"), _b(sd, p + ".out_layers.0.bias"))
    h = conv2d(silu(h), _w(sd, p + ".out_layers.3.weight"), _b(sd, p + ".out_layers.3.bias"), padding=(1, 1))
    if (p + ".skip_connection.weight") in sd:
        x = conv2d(x, _w(sd, p + ".skip_connection.weight"), _b(sd, p + ".skip_connection.bias"))
    return x + h


def upsample(sd, p, x):
    # vision/unet.py:78-84 — nearest x2 then 3x3 conv
    bs, c, py, px = x.shape
    x = x.reshape(bs, c, py, 1, px, 1).expand(bs, c, py, 2, px, 2).reshape(bs, c, py * 2, px * 2)
    return conv2d(x, _w(sd, p + ".conv.weight"), _b(sd, p + ".conv.bias"), padding=(1, 1))


def downsample(sd, p, x):
    # vision/unet.py:86-90 — 3x3 stride-2 pad-1 conv
    return conv2d(x, _w(sd, p + ".op.weight"), _b(sd, p + ".op.bias"), stride=(2, 2), padding=(1, 1))


# UNet structure, vision/unet.py:11-49. Entries: ("conv", cin, cout) | ("res", cin, cout) |
# ("st", channels, heads, d_head) | ("down", c) | ("up", c)
UNET_INPUT_BLOCKS = [
    [("conv", 4, 320)],
    [("res", 320, 320), ("st", 320, 8, 40)],
    [("res", 320, 320), ("st", 320, 8, 40)],
    [("down", 320)],
    [("res", 320, 640), ("st", 640, 8, 80)],
    [("res", 640, 640), ("st", 640, 8, 80)],
    [("down", 640)],
    [("res", 640, 1280), ("st", 1280, 8, 160)],
    [("res", 1280, 1280), ("st", 1280, 8, 160)],
    [("down", 1280)],
    [("res", 1280, 1280)],
    [("res", 1280, 1280)],
]
UNET_MIDDLE_BLOCK = [("res", 1280, 1280), ("st", 1280, 8, 160), ("res", 1280, 1280)]
UNET_OUTPUT_BLOCKS = [
    [("res", 2560, 1280)],
    [("res", 2560, 1280)],
    [("res", 2560, 1280), ("up", 1280)],
    [("res", 2560, 1280), ("st", 1280, 8, 160)],
    [("res", 2560, 1280), ("st", 1280, 8, 160)],
    [("res", 1920, 1280), ("st", 1280, 8, 160), ("up", 1280)],
    [("res", 1920, 640), ("st", 640, 8, 80)],
    [("res", 1280, 640), ("st", 640, 8, 80)],
    [("res", 960, 640), ("st", 640, 8, 80), ("up", 640)],
    [("res", 960, 320), ("st", 320, 8, 40)],
    [("res", 640, 320), ("st", 320, 8, 40)],
    [("res", 640, 320), ("st", 320, 8, 40)],
]
CONTEXT_DIM = 768
EMB_CHANNELS = 1280


def _run_layer(sd, p, layer, x, emb, context, quirks, ln_strided=False):
    kind = layer[0]
    if kind == "conv":
        return conv2d(x, _w(sd, p + ".weight"), _b(sd, p + ".bias"), padding=(1, 1))
    if kind == "res":
        return res_block(sd, p, x, emb)
    if kind == "st":
        return spatial_transformer(sd, p, x, context, layer[2], layer[3], quirks, ln_strided)
    if kind == "down":
        return downsample(sd, p, x)
    if kind == "up":
        return upsample(sd, p, x)
    raise ValueError(kind)


def unet_forward(sd, x, timesteps, context, prefix="model.diffusion_model", quirks=True, ln_strided=False):
    # vision/unet.py:51-76
    P = prefix
    t_emb = timestep_embedding(timesteps, 320).to(device=x.device, dtype=x.dtype)
    emb = linear(t_emb, _w(sd, P + ".time_embed.0.weight"), _b(sd, P + ".time_embed.0.bias"))
    emb = linear(silu(emb), _w(sd, P + ".time_embed.2.weight"), _b(sd, P + ".time_embed.2.bias"))
    saved = []
    for i, block in enumerate(UNET_INPUT_BLOCKS):
        for j, layer in enumerate(block):
            x = _run_layer(sd, f"{P}.input_blocks.{i}.{j}", layer, x, emb, context, quirks, ln_strided)
        saved.append(x)
    for j, layer in enumerate(UNET_MIDDLE_BLOCK):
        x = _run_layer(sd, f"{P}.middle_block.{j}", layer, x, emb, context, quirks, ln_strided)
    for i, block in enumerate(UNET_OUTPUT_BLOCKS):
        x = torch.cat((x, saved.pop()), dim=1)
        for j, layer in enumerate(block):
            x = _run_layer(sd, f"{P}.output_blocks.{i}.{j}", layer, x, emb, context, quirks, ln_strided)
    x = group_norm_affine(x, 32, _w(sd, P + ".out.0.weight"), _b(sd, P + ".out.0.bias"))
    return conv2d(silu(x), _w(sd, P + ".out.2.weight"), _b(sd, P + ".out.2.bias"), padding=(1, 1))


# ------------------------------------------------------------------------------------------------
# sampler step — reference: tinyfusers/variants/sd.py:14-59
# ------------------------------------------------------------------------------------------------


def get_x_prev_and_pred_x0(x, e_t, a_t, a_prev):
    # sd.py:14-25 (DDIM, eta = 0 => sigma_t = 0)
    sigma_t = 0
    sqrt_one_minus_at = torch.sqrt(1 - a_t)
    pred_x0 = (x - sqrt_one_minus_at * e_t) / torch.sqrt(a_t)
    dir_xt = torch.sqrt(1.0 - a_prev - sigma_t ** 2) * e_t
    x_prev = torch.sqrt(a_prev) * pred_x0 + dir_xt
    return x_prev, pred_x0


def get_model_output(sd, unconditional_context, context, latent, timestep, guidance, quirks=True, ln_strided=False):
    # sd.py:27-46: batch = [uncond ; cond], e_t = u + g (c - u)
    n = latent.shape[0]
    lat2 = torch.cat((latent, latent), dim=0) if n > 1 else latent.expand(2, *latent.shape[1:])
    ctx2 = torch.cat((unconditional_context, context), dim=0)
    out = unet_forward(sd, lat2.contiguous(), timestep, ctx2, quirks=quirks, ln_strided=ln_strided)
    u, c = out[0:n], out[n:2 * n]
    return u + guidance * (c - u)


def sampler_step(sd, unconditional_context, context, latent, timestep, a_t, a_prev, guidance, quirks=True, ln_strided=False):
    # sd.py:56-59
    e_t = get_model_output(sd, unconditional_context, context, latent, timestep, guidance, quirks, ln_strided)
    x_prev, _ = get_x_prev_and_pred_x0(latent, e_t, a_t, a_prev)
    return x_prev


def sampler_schedule(steps, alphas_cumprod=None):
    # example/sd1.py:54-57
    timesteps = list(range(1, 1000, 1000 // steps))
    ac = get_alphas_cumprod() if alphas_cumprod is None else alphas_cumprod
    alphas = ac[timesteps]
    alphas_prev = torch.cat((torch.tensor([1.0]), alphas[:-1])).float()
    return timesteps, alphas, alphas_prev


# ------------------------------------------------------------------------------------------------
# VAE decoder — reference: tinyfusers/vae/decoder.py:8-34, vae/mid.py:5-12,
# vision/resnet.py:33-45, attention/attention.py:10-24, vae/vae.py:5-18, variants/sd.py:48-54
# ------------------------------------------------------------------------------------------------


def _pw(sd, p):
    return sd[p + ".weight"]


def _pb(sd, p):
    return sd.get(p + ".bias")


def _gn(sd, p, x, eps=1e-5):
    return group_norm_affine(x, 32, _pw(sd, p), _pb(sd, p), eps)


def _conv(sd, p, x, padding=(0, 0), stride=(1, 1)):
    return conv2d(x, _pw(sd, p), _pb(sd, p), stride=stride, padding=padding)


def resnet_block(sd, p, x):
    # vision/resnet.py:41-45   conv1(swish(norm1 x)) -> conv2(swish(norm2 .)) ; + nin_shortcut(x) (1x1) | x
    h = _conv(sd, p + ".conv1", silu(_gn(sd, p + ".norm1", x)), padding=(1, 1))
    h = _conv(sd, p + ".conv2", silu(_gn(sd, p + ".norm2", h)), padding=(1, 1))
    sc = _conv(sd, p + ".nin_shortcut", x) if (p + ".nin_shortcut.weight") in sd else x
    return sc + h


def attn_block(sd, p, x, quirks=True):
    """attention/attention.py:19-24. The reference hands the 4-D (B,C,H,W) q/k/v straight to SDPA
    (attention.py:21-22), which reads them as (B, NH=C, T=H, HS=W): per channel, an H x H attention over the
    rows of that channel's H x W plane, scale 1/sqrt(W) (SURVEY.md §8 parity note 3). quirks=False is the
    canonical LDM AttnBlock: one head over the H*W pixels, head dim C."""
    h_ = _gn(sd, p + ".norm", x)
    q, k, v = (_conv(sd, f"{p}.{n}", h_) for n in ("q", "k", "v"))
    if quirks:
        h = scaled_dot_product_attention(q, k, v)
    else:
        B, C, H, W = q.shape
        tok = lambda t: t.reshape(B, C, H * W).transpose(1, 2).reshape(B, 1, H * W, C)
        h = scaled_dot_product_attention(tok(q), tok(k), tok(v)).reshape(B, H * W, C).transpose(1, 2).reshape(B, C, H, W)
    return x + _conv(sd, p + ".proj_out", h)


def vae_mid(sd, p, x, quirks=True):
    # vae/mid.py:11-12
    x = resnet_block(sd, p + ".block_1", x)
    x = attn_block(sd, p + ".attn_1", x, quirks)
    return resnet_block(sd, p + ".block_2", x)


VAE_DECODER_SZ = [(128, 256), (256, 512), (512, 512), (512, 512)]   # vae/decoder.py:10


def vae_decoder(sd, p, x, quirks=True):
    # vae/decoder.py:22-34
    x = _conv(sd, p + ".conv_in", x, padding=(1, 1))
    x = vae_mid(sd, p + ".mid", x, quirks)
    for i in (3, 2, 1, 0):
        for j in range(3):
            x = resnet_block(sd, f"{p}.up.{i}.block.{j}", x)
        if i != 0:
            bs, c, py, px = x.shape   # nearest x2 (decoder.py:30-31)
            x = x.reshape(bs, c, py, 1, px, 1).expand(bs, c, py, 2, px, 2).reshape(bs, c, py * 2, px * 2)
            x = _conv(sd, f"{p}.up.{i}.upsample.conv", x, padding=(1, 1))
    return _conv(sd, p + ".conv_out", silu(_gn(sd, p + ".norm_out", x)), padding=(1, 1))


def vae_decode_float(sd, x, prefix="first_stage_model", quirks=True):
    """variants/sd.py:48-51: post_quant_conv(x / 0.18215) -> decoder -> (x + 1) / 2, before clipping / uint8."""
    x = _conv(sd, prefix + ".post_quant_conv", (1 / 0.18215) * x)
    x = vae_decoder(sd, prefix + ".decoder", x, quirks)
    return (x + 1.0) / 2.0


def vae_decode(sd, x, prefix="first_stage_model", quirks=True):
    """variants/sd.py:48-54 with the hard-coded 512 (sd.py:52) generalised to the decoded size:
    clip(transpose(x.reshape(3,H,W), (1,2,0)), 0, 1) * 255 -> uint8 (truncation, like cp.astype)."""
    y = vae_decode_float(sd, x, prefix, quirks)
    _, _, H, W = y.shape
    img = torch.clamp(y.reshape(3, H, W).permute(1, 2, 0), 0, 1) * 255
    return img.to(torch.uint8)


# ------------------------------------------------------------------------------------------------
# CLIP text encoder — reference: tinyfusers/vae/encoder.py:36-81,
# attention/attention.py:78-99, ff/nn.py:25-34, ff/embedding.py:6-23
# ------------------------------------------------------------------------------------------------


def clip_attention(sd, p, x, mask):
    # attention/attention.py:88-99: 12 heads x 64, q/k/v/out Linear WITH bias, heads transposed back (canonical)
    B, T, E = x.shape
    NH, HD = 12, E // 12
    q, k, v = (linear(x, _pw(sd, f"{p}.{n}"), _pb(sd, f"{p}.{n}")) for n in ("q_proj", "k_proj", "v_proj"))
    q, k, v = (t.reshape(B, T, NH, HD).transpose(1, 2) for t in (q, k, v))
    o = scaled_dot_product_attention(q, k, v, attn_mask=mask)
    o = o.transpose(1, 2).reshape(B, T, E)
    return linear(o, _pw(sd, p + ".out_proj"), _pb(sd, p + ".out_proj"))


def clip_encoder_layer(sd, p, x, mask):
    # vae/encoder.py:53-64
    h = layer_norm(x, _pw(sd, p + ".layer_norm1"), _pb(sd, p + ".layer_norm1"))
    x = x + clip_attention(sd, p + ".self_attn", h, mask)
    h = layer_norm(x, _pw(sd, p + ".layer_norm2"), _pb(sd, p + ".layer_norm2"))
    h = linear(h, _pw(sd, p + ".mlp.fc1"), _pb(sd, p + ".mlp.fc1"))     # ff/nn.py:30-34
    h = linear(quick_gelu(h), _pw(sd, p + ".mlp.fc2"), _pb(sd, p + ".mlp.fc2"))
    return x + h


def clip_text_transformer(sd, input_ids, prefix="cond_stage_model.transformer.text_model", layers=12):
    """vae/encoder.py:72-81. token + position embedding -> 12 pre-LN layers under the additive causal mask
    triu(full((1,1,77,77), -inf), k=1) -> final LayerNorm. The reference's Embedding (ff/embedding.py:15-23) builds
    a one-hot matrix of the wrong shape (embed_sz x N instead of vocab x N) and cannot run; what it means —
    row lookup, weight[idx] — is restated here."""
    P = prefix
    ids = torch.as_tensor(input_ids, dtype=torch.long).reshape(1, -1)
    T = ids.shape[1]
    x = _pw(sd, P + ".embeddings.token_embedding")[ids[0]] + _pw(sd, P + ".embeddings.position_embedding")[torch.arange(T)]
    x = x.reshape(1, T, -1)
    mask = torch.triu(torch.full((1, 1, T, T), float("-inf")), diagonal=1).to(x.dtype)
    for i in range(layers):
        x = clip_encoder_layer(sd, f"{P}.encoder.layers.{i}", x, mask)
    return layer_norm(x, _pw(sd, P + ".final_layer_norm"), _pb(sd, P + ".final_layer_norm"))


# ------------------------------------------------------------------------------------------------
# synthetic weights: deterministic per key, identical for oracle and kernels.
# ------------------------------------------------------------------------------------------------


def _key_seed(key, seed):
    h = 2166136261
    for ch in key.encode():
        h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
    return (seed * 1000003 + h) & 0x7FFFFFFF


def _randn(key, seed, shape, std):
    rng = np.random.Generator(np.random.Philox(_key_seed(key, seed)))
    return torch.from_numpy((rng.standard_normal(size=shape, dtype=np.float32) * np.float32(std)))


def _add_conv(sd, p, cin, cout, k, seed):
    fan_in = cin * k * k
    sd[p + ".weight"] = _randn(p + ".weight", seed, (cout, cin, k, k), 1.0 / math.sqrt(fan_in))
    sd[p + ".bias"] = _randn(p + ".bias", seed, (cout,), 0.02)


def _add_linear(sd, p, cin, cout, seed, bias=True):
    sd[p + ".weight"] = _randn(p + ".weight", seed, (cout, cin), 1.0 / math.sqrt(cin))
    if bias:
        sd[p + ".bias"] = _randn(p + ".bias", seed, (cout,), 0.02)


def _add_norm(sd, p, c, seed):
    sd[p + ".weight"] = 1.0 + _randn(p + ".weight", seed, (c,), 0.02)
    sd[p + ".bias"] = _randn(p + ".bias", seed, (c,), 0.02)


def add_res_block(sd, p, cin, cout, seed=1234, emb=EMB_CHANNELS):
    _add_norm(sd, p + ".in_layers.0", cin, seed)
    _add_conv(sd, p + ".in_layers.2", cin, cout, 3, seed)
    _add_linear(sd, p + ".emb_layers.1", emb, cout, seed)
    _add_norm(sd, p + ".out_layers.0", cout, seed)
    _add_conv(sd, p + ".out_layers.3", cout, cout, 3, seed)
    if cin != cout:
        _add_conv(sd, p + ".skip_connection", cin, cout, 1, seed)


def add_spatial_transformer(sd, p, c, context_dim=CONTEXT_DIM, seed=1234):
    _add_norm(sd, p + ".norm", c, seed)
    _add_conv(sd, p + ".proj_in", c, c, 1, seed)
    t = p + ".transformer_blocks.0"
    for attn, cd in (("attn1", c), ("attn2", context_dim)):
        _add_linear(sd, f"{t}.{attn}.to_q", c, c, seed, bias=False)
        _add_linear(sd, f"{t}.{attn}.to_k", cd, c, seed, bias=False)
        _add_linear(sd, f"{t}.{attn}.to_v", cd, c, seed, bias=False)
        _add_linear(sd, f"{t}.{attn}.to_out.0", c, c, seed)
    _add_linear(sd, t + ".ff.net.0.proj", c, 8 * c, seed)
    _add_linear(sd, t + ".ff.net.2", 4 * c, c, seed)
    for n in ("norm1", "norm2", "norm3"):
        _add_norm(sd, f"{t}.{n}", c, seed)
    _add_conv(sd, p + ".proj_out", c, c, 1, seed)


def _add_layer(sd, p, layer, seed):
    kind = layer[0]
    if kind == "conv":
        _add_conv(sd, p, layer[1], layer[2], 3, seed)
    elif kind == "res":
        add_res_block(sd, p, layer[1], layer[2], seed)
    elif kind == "st":
        add_spatial_transformer(sd, p, layer[1], CONTEXT_DIM, seed)
    elif kind == "down":
        _add_conv(sd, p + ".op", layer[1], layer[1], 3, seed)
    elif kind == "up":
        _add_conv(sd, p + ".conv", layer[1], layer[1], 3, seed)


def make_unet_state_dict(seed=1234, prefix="model.diffusion_model"):
    """Seeded synthetic UNet weights with the reference's checkpoint key names (fp32, ~3.4 GB)."""
    sd = {}
    P = prefix
    _add_linear(sd, P + ".time_embed.0", 320, 1280, seed)
    _add_linear(sd, P + ".time_embed.2", 1280, 1280, seed)
    for i, block in enumerate(UNET_INPUT_BLOCKS):
        for j, layer in enumerate(block):
            _add_layer(sd, f"{P}.input_blocks.{i}.{j}", layer, seed)
    for j, layer in enumerate(UNET_MIDDLE_BLOCK):
        _add_layer(sd, f"{P}.middle_block.{j}", layer, seed)
    for i, block in enumerate(UNET_OUTPUT_BLOCKS):
        for j, layer in enumerate(block):
            _add_layer(sd, f"{P}.output_blocks.{i}.{j}", layer, seed)
    _add_norm(sd, P + ".out.0", 320, seed)
    _add_conv(sd, P + ".out.2", 320, 4, 3, seed)
    return sd


def add_resnet_block(sd, p, cin, cout, seed=1234):
    _add_norm(sd, p + ".norm1", cin, seed)
    _add_conv(sd, p + ".conv1", cin, cout, 3, seed)
    _add_norm(sd, p + ".norm2", cout, seed)
    _add_conv(sd, p + ".conv2", cout, cout, 3, seed)
    if cin != cout:
        _add_conv(sd, p + ".nin_shortcut", cin, cout, 1, seed)


def add_attn_block(sd, p, c, seed=1234):
    _add_norm(sd, p + ".norm", c, seed)
    for n in ("q", "k", "v", "proj_out"):
        _add_conv(sd, f"{p}.{n}", c, c, 1, seed)


def make_vae_decoder_state_dict(seed=4321, prefix="first_stage_model"):
    """Seeded synthetic post_quant_conv + Decoder weights under the reference's checkpoint key names (~198 MB fp32)."""
    sd = {}
    _add_conv(sd, prefix + ".post_quant_conv", 4, 4, 1, seed)
    D = prefix + ".decoder"
    _add_conv(sd, D + ".conv_in", 4, 512, 3, seed)
    add_resnet_block(sd, D + ".mid.block_1", 512, 512, seed)
    add_attn_block(sd, D + ".mid.attn_1", 512, seed)
    add_resnet_block(sd, D + ".mid.block_2", 512, 512, seed)
    for i, (lo, hi) in enumerate(VAE_DECODER_SZ):
        add_resnet_block(sd, f"{D}.up.{i}.block.0", hi, lo, seed)
        add_resnet_block(sd, f"{D}.up.{i}.block.1", lo, lo, seed)
        add_resnet_block(sd, f"{D}.up.{i}.block.2", lo, lo, seed)
        if i != 0:
            _add_conv(sd, f"{D}.up.{i}.upsample.conv", lo, lo, 3, seed)
    _add_norm(sd, D + ".norm_out", 128, seed)
    _add_conv(sd, D + ".conv_out", 128, 3, 3, seed)
    return sd


def make_clip_state_dict(seed=777, prefix="cond_stage_model.transformer.text_model", layers=12):
    """Seeded synthetic CLIP text-encoder weights under the reference's checkpoint key names (~490 MB fp32)."""
    sd = {}
    P = prefix
    sd[P + ".embeddings.token_embedding.weight"] = _randn(P + ".tok", seed, (49408, 768), 0.02)
    sd[P + ".embeddings.position_embedding.weight"] = _randn(P + ".pos", seed, (77, 768), 0.01)
    for i in range(layers):
        L = f"{P}.encoder.layers.{i}"
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            _add_linear(sd, f"{L}.self_attn.{n}", 768, 768, seed)
        _add_norm(sd, L + ".layer_norm1", 768, seed)
        _add_norm(sd, L + ".layer_norm2", 768, seed)
        _add_linear(sd, L + ".mlp.fc1", 768, 3072, seed)
        _add_linear(sd, L + ".mlp.fc2", 3072, 768, seed)
    _add_norm(sd, P + ".final_layer_norm", 768, seed)
    return sd


def make_inputs(batch=1, latent_hw=64, seed=42, ctx_seed=43):
    """SURVEY.md §8d synthetic inputs: latent ~ N(0,1) seed 42, prompt embeddings ~ N(0,1) seed 43."""
    g = np.random.Generator(np.random.Philox(seed))
    latent = torch.from_numpy(g.standard_normal((batch, 4, latent_hw, latent_hw), dtype=np.float32))
    g2 = np.random.Generator(np.random.Philox(ctx_seed))
    ctx = torch.from_numpy(g2.standard_normal((batch, 77, CONTEXT_DIM), dtype=np.float32))
    unc = torch.from_numpy(g2.standard_normal((batch, 77, CONTEXT_DIM), dtype=np.float32))
    return latent, unc, ctx


# ------------------------------------------------------------------------------------------------
# algorithmic FLOPs from the structure tables (used to scale a bounded CPU sample to the 64x64 step)
# ------------------------------------------------------------------------------------------------


def unet_step_flops(n, H, W, ctx_tokens=77):
    """2*M*N*K per conv/linear + 4*B*NH*Tq*Tk*d per attention for one UNet forward at batch n."""
    total = 2.0 * (320 * 1280 + 1280 * 1280)
    h, w = H, W

    def layer_flops(layer, h, w):
        kind = layer[0]
        if kind == "conv":
            return 2.0 * n * h * w * layer[2] * layer[1] * 9
        if kind == "res":
            cin, cout = layer[1], layer[2]
            f = 2.0 * n * h * w * cout * (cin + cout) * 9 + 2.0 * EMB_CHANNELS * cout
            return f + (2.0 * n * h * w * cin * cout if cin != cout else 0.0)
        if kind == "st":
            c, nh, d = layer[1], layer[2], layer[3]
            T = h * w
            f = 2 * 2.0 * n * T * c * c                                   # proj_in / proj_out
            f += 4 * 2.0 * n * T * c * c + 4.0 * n * nh * T * T * d       # self-attention
            f += 2 * 2.0 * n * T * c * c + 2 * 2.0 * n * ctx_tokens * CONTEXT_DIM * c + 4.0 * n * nh * T * ctx_tokens * d
            f += 2.0 * n * T * c * 8 * c + 2.0 * n * T * 4 * c * c        # GEGLU feed-forward
            return f
        return 0.0

    for group in (UNET_INPUT_BLOCKS, [UNET_MIDDLE_BLOCK], UNET_OUTPUT_BLOCKS):
        for block in group:
            for layer in block:
                if layer[0] == "down":
                    h, w = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
                    total += 2.0 * n * h * w * layer[1] * layer[1] * 9
                elif layer[0] == "up":
                    h, w = 2 * h, 2 * w
                    total += 2.0 * n * h * w * layer[1] * layer[1] * 9
                else:
                    total += layer_flops(layer, h, w)
    total += 2.0 * n * h * w * 4 * 320 * 9
    return total
