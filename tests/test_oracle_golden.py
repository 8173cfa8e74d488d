"""The oracle (oracle/ref_ops.py) against (1) golden vectors produced by the reference's OWN Python
(oracle/make_golden.py -> tests/golden/reference_outputs.npz) and (2) the torch operators the reference's own
tests use as yardsticks (tests/conv2d.py:27-33, group_norm.py:33-40, layer_norm.py:38-41, sdpa.py:97-100)."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_outputs.npz"))
TOL = 2e-5   # fp32 on both sides (numpy vs torch reduction orders)


def rnd(seed, *shape, scale=1.0, shift=0.0):
    g = np.random.Generator(np.random.Philox(seed))
    return torch.from_numpy(g.standard_normal(shape, dtype=np.float32) * np.float32(scale) + np.float32(shift))


def gold(name):
    return torch.from_numpy(GOLD[name])


def test_activations(oracle):
    x = torch.linspace(-9, 9, 1801)
    for name in ("sigmoid", "silu", "swish", "gelu", "quick_gelu"):
        fn = getattr(oracle, "silu" if name == "swish" else name)
        assert (fn(x) - gold(f"act_{name}")).abs().max().item() < 2e-6, name


def test_group_norm(oracle):
    x = rnd(101, 2, 320, 8, 8, scale=1.7, shift=0.6)
    assert rel_err(oracle.group_norm(x, 32, 1e-5), gold("group_norm_noaffine")) < TOL
    w, b = 1 + 0.1 * rnd(102, 320), 0.1 * rnd(103, 320)
    out = oracle.group_norm_affine(x, 32, w, b, 1e-5)
    assert rel_err(out, gold("group_norm_affine")) < TOL
    assert rel_err(out, F.group_norm(x, 32, w, b, 1e-5)) < TOL           # reference tests/group_norm.py:33-40


def test_linear_and_feed_forward(oracle):
    sd = {}
    oracle._add_linear(sd, "lin", 320, 640, 201)
    assert rel_err(oracle.linear(rnd(202, 2, 16, 320), sd["lin.weight"], sd["lin.bias"]), gold("linear")) < TOL
    sd = {}
    oracle._add_linear(sd, "ff.net.0.proj", 320, 2560, 203)
    oracle._add_linear(sd, "ff.net.2", 1280, 320, 203)
    x = rnd(204, 2, 16, 320)
    assert rel_err(oracle.geglu(sd, "ff.net.0", x), gold("geglu")) < TOL
    assert rel_err(oracle.feed_forward(sd, "ff", x), gold("feed_forward")) < TOL


def test_timestep_embedding_schedule_ddim(oracle):
    te = torch.cat([oracle.timestep_embedding([t], 320) for t in (1, 21, 501, 981)])
    assert (te - gold("timestep_embedding")).abs().max().item() < 1e-6
    ac = oracle.get_alphas_cumprod()
    assert rel_err(ac, gold("alphas_cumprod")) < 1e-6
    xp, p0 = oracle.get_x_prev_and_pred_x0(rnd(301, 1, 4, 8, 8), rnd(302, 1, 4, 8, 8), ac[[501]], ac[[481]])
    assert rel_err(xp, gold("ddim_x_prev")) < TOL and rel_err(p0, gold("ddim_pred_x0")) < TOL


def test_sdpa(oracle):
    q, k, v = rnd(401, 2, 8, 64, 40), rnd(402, 2, 8, 77, 40), rnd(403, 2, 8, 77, 40)
    out = oracle.scaled_dot_product_attention(q, k, v)
    assert rel_err(out, gold("sdpa")) < TOL
    assert rel_err(out, F.scaled_dot_product_attention(q, k, v, scale=1 / math.sqrt(40))) < TOL  # tests/sdpa.py:97-100


def test_attention_blocks(oracle):
    sd = {}
    oracle.add_spatial_transformer(sd, "st", 320, 768, seed=501)
    tb = "st.transformer_blocks.0"
    xt, ctx = rnd(502, 2, 64, 320), rnd(503, 2, 77, 768)
    assert rel_err(oracle.cross_attention(sd, tb + ".attn1", xt, None, 8, 40, True), gold("cross_attention_self")) < TOL
    assert rel_err(oracle.cross_attention(sd, tb + ".attn2", xt, ctx, 8, 40, True), gold("cross_attention_ctx")) < TOL
    assert rel_err(oracle.basic_transformer_block(sd, tb, xt, ctx, 8, 40, True), gold("transformer_block_b2")) < 5e-5
    assert rel_err(oracle.basic_transformer_block(sd, tb, xt[:1], ctx[:1], 8, 40, True), gold("transformer_block_b1")) < 5e-5
    out = oracle.spatial_transformer(sd, "st", rnd(504, 2, 320, 8, 8), ctx, 8, 40, True)
    assert rel_err(out, gold("spatial_transformer")) < 5e-5


def test_quirks_are_observable(oracle):
    """the reference's head-major reshape differs from the canonical head merge: the goldens pin the literal one"""
    sd = {}
    oracle.add_spatial_transformer(sd, "st", 320, 768, seed=501)
    tb = "st.transformer_blocks.0"
    xt, ctx = rnd(502, 2, 64, 320), rnd(503, 2, 77, 768)
    canon = oracle.basic_transformer_block(sd, tb, xt, ctx, 8, 40, False)
    assert rel_err(canon, gold("transformer_block_b2")) > 1e-2
    canon_attn = oracle.cross_attention(sd, tb + ".attn1", xt, None, 8, 40, False)
    assert rel_err(canon_attn, gold("cross_attention_self")) > 1e-2


def test_layer_norm_matches_torch(oracle):
    """reference tests/layer_norm.py:38-41 pins the op to torch.nn.functional.layer_norm"""
    x = rnd(7, 1, 64, 320, scale=2.0, shift=0.3)
    w, b = 1 + 0.1 * rnd(8, 320), 0.1 * rnd(9, 320)
    assert rel_err(oracle.layer_norm(x, w, b, 1e-5), F.layer_norm(x, (320,), w, b, 1e-5)) < TOL
    assert rel_err(oracle.layer_norm(x, w, b, 1e-5, ln_strided=True), F.layer_norm(x, (320,), w, b, 1e-5)) < TOL  # B == 1
    x2 = rnd(10, 2, 64, 320)
    assert rel_err(oracle.layer_norm(x2, w, b, 1e-5), F.layer_norm(x2, (320,), w, b, 1e-5)) < TOL
    # the literal stride reading is a different function at B > 1 (normalises memory viewed as (T, C, B) over C)
    lit = oracle.layer_norm(x2, w, b, 1e-5, ln_strided=True)
    mem = x2.reshape(64, 320, 2)
    want = (F.layer_norm(mem.permute(0, 2, 1), (320,), w, b, 1e-5)).permute(0, 2, 1).reshape(2, 64, 320)
    assert rel_err(lit, want) < TOL and rel_err(lit, F.layer_norm(x2, (320,), w, b, 1e-5)) > 1e-2


def test_res_block_up_down(oracle):
    sd = {}
    oracle.add_res_block(sd, "rb", 320, 640, seed=601)
    assert rel_err(oracle.res_block(sd, "rb", rnd(602, 2, 320, 8, 8), rnd(603, 1, 1280)), gold("res_block")) < 5e-5
    sd = {}
    oracle._add_conv(sd, "u.conv", 64, 64, 3, 604)
    oracle._add_conv(sd, "d.op", 64, 64, 3, 604)
    assert rel_err(oracle.upsample(sd, "u", rnd(605, 1, 64, 6, 6)), gold("upsample")) < TOL
    assert rel_err(oracle.downsample(sd, "d", rnd(606, 1, 64, 12, 12)), gold("downsample")) < TOL


def test_unet_cfg_and_sampler_step(oracle, unet_sd):
    lat, unc, ctx = oracle.make_inputs(1, 16)
    x2, c2 = torch.cat([lat, lat]), torch.cat([unc, ctx])
    with torch.no_grad():
        assert rel_err(oracle.unet_forward(unet_sd, x2, [981], c2, quirks=True), gold("unet_16")) < 2e-4
        assert rel_err(oracle.get_model_output(unet_sd, unc, ctx, lat, [501], 7.5), gold("cfg_e_t_16")) < 5e-4
        ac = oracle.get_alphas_cumprod()
        out = oracle.sampler_step(unet_sd, unc, ctx, lat, [501], ac[[25]], ac[[24]], 7.5)
        assert rel_err(out, gold("sampler_step_16")) < 5e-4


def test_cudnn_layernorm_probe_recorded():
    """oracle/cudnn_probe.py ran the reference's layernorm graph on real cuDNN (GPU box): at B = 1 cuDNN equals the
    oracle's (canonical) LayerNorm; at B > 1 cuDNN rejects the reference's stride declaration."""
    import json
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cudnn_layernorm_probe.json")
    rec = json.load(open(path))
    b1 = [c for c in rec["cases"] if c["B"] == 1]
    bn = [c for c in rec["cases"] if c["B"] > 1]
    assert b1 and all(c["cudnn"] == "executed" and c["rel_err_vs_canonical_layernorm"] < 1e-5 for c in b1), rec
    assert bn and all(c["cudnn"].startswith("rejected") or c["rel_err_vs_stride_model"] < 1e-4 for c in bn), rec
