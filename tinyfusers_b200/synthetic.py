"""Seeded synthetic weights / inputs and the algorithmic FLOP count of one UNet step.

There is no checkpoint, dataset or network offline, so `example/sd1.py` (without --ckpt) and `bench.py` build the model from
these generators: every tensor is a deterministic function of (its checkpoint key, seed), under the reference's own
checkpoint key names (tinyfusers/storage/state.py:4-23 walks them; structure tables below = tinyfusers/vision/unet.py:11-49,
vae/decoder.py:8-34, vae/encoder.py:36-81). conv / linear W ~ N(0, 1/fan_in), biases N(0, 0.02^2), norm gamma = 1 + N(0, 0.02^2),
beta = N(0, 0.02^2). Host-side numpy only; nothing here touches the GPU.

The oracle (`oracle/ref_ops.py`, test infrastructure) carries its own copy of these generators so that it stays
self-contained; `tests/test_synthetic_cpu.py` holds the two to bit-identical output.
"""
import math

import numpy as np
import torch

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
VAE_DECODER_SZ = [(128, 256), (256, 512), (512, 512), (512, 512)]   # vae/decoder.py:10


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


def sampler_schedule(steps, alphas_cumprod=None):
    """timesteps, alphas, alphas_prev of the reference's sampler loop (example/sd1.py:54-57), host tensors."""
    timesteps = list(range(1, 1000, 1000 // steps))
    if alphas_cumprod is None:
        betas = np.linspace(0.00085 ** 0.5, 0.0120 ** 0.5, 1000, dtype=np.float32) ** 2      # variants/sd.py:61-65
        alphas_cumprod = torch.from_numpy(np.cumprod((1.0 - betas).astype(np.float32), axis=0).astype(np.float32))
    ac = alphas_cumprod.detach().cpu()
    alphas = ac[timesteps]
    alphas_prev = torch.cat((torch.tensor([1.0]), alphas[:-1])).float()
    return timesteps, alphas, alphas_prev


# ------------------------------------------------------------------------------------------------
# algorithmic FLOPs from the structure tables (scales a bounded CPU sample to the 64x64 step)
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
