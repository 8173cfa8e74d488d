"""The oracle's VAE-decoder and CLIP text-encoder restatements (SURVEY.md §8f "next" rows) against golden vectors
produced by the reference's OWN Python (oracle/make_golden_vae.py -> tests/golden/reference_outputs_vae_clip.npz)."""
import os

import numpy as np
import torch

from conftest import rel_err

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_outputs_vae_clip.npz"))
TOL = 2e-5


def rnd(seed, *shape, scale=1.0, shift=0.0):
    g = np.random.Generator(np.random.Philox(seed))
    return torch.from_numpy(g.standard_normal(shape, dtype=np.float32) * np.float32(scale) + np.float32(shift))


def gold(name):
    return torch.from_numpy(GOLD[name])


def test_resnet_block(oracle):
    sd = {}
    oracle.add_resnet_block(sd, "rn", 64, 128, seed=701)
    oracle.add_resnet_block(sd, "rn2", 64, 64, seed=702)
    x = rnd(703, 1, 64, 8, 8, scale=1.3, shift=0.2)
    assert rel_err(oracle.resnet_block(sd, "rn", x), gold("resnet_block_64_128")) < TOL
    assert rel_err(oracle.resnet_block(sd, "rn2", x), gold("resnet_block_64_64")) < TOL


def test_attn_block_reference_reading(oracle):
    """4-D q/k/v into SDPA = per-channel attention over the rows of the H x W plane (parity note 3)."""
    sd = {}
    oracle.add_attn_block(sd, "ab", 64, seed=711)
    x = rnd(712, 1, 64, 6, 10)
    assert rel_err(oracle.attn_block(sd, "ab", x), gold("attn_block_6x10")) < TOL
    assert rel_err(oracle.attn_block(sd, "ab", rnd(713, 1, 64, 8, 8)), gold("attn_block_8x8")) < TOL
    # ... and it is NOT the canonical single-head block
    assert rel_err(oracle.attn_block(sd, "ab", x, quirks=False), gold("attn_block_6x10")) > 1e-2


def test_mid(oracle):
    sd = {}
    oracle.add_resnet_block(sd, "mid.block_1", 64, 64, seed=721)
    oracle.add_attn_block(sd, "mid.attn_1", 64, seed=721)
    oracle.add_resnet_block(sd, "mid.block_2", 64, 64, seed=721)
    assert rel_err(oracle.vae_mid(sd, "mid", rnd(722, 1, 64, 8, 8)), gold("mid_64")) < TOL


def test_decoder_and_decode(oracle):
    vsd = oracle.make_vae_decoder_state_dict()
    with torch.no_grad():
        y = oracle.vae_decoder(vsd, "first_stage_model.decoder", rnd(731, 1, 4, 4, 4))
    assert y.shape == (1, 3, 32, 32)
    assert rel_err(y, gold("decoder_4x4")) < 1e-4      # 30 convs deep, numpy vs torch reduction orders
    # decode post-processing: post_quant_conv(x / 0.18215), (x + 1) / 2, clip, * 255, uint8 truncation
    z = rnd(741, 1, 4, 64, 64)
    zq = oracle._conv(vsd, "first_stage_model.post_quant_conv", (1 / 0.18215) * z)
    assert rel_err(zq, gold("decode_postquant_in")) < TOL
    yy, xx = np.meshgrid(np.linspace(-1.6, 1.6, 512, dtype=np.float32), np.linspace(-1.2, 1.2, 512, dtype=np.float32), indexing="ij")
    field = torch.from_numpy(np.stack([yy * xx, yy + 0.3 * xx, np.sin(3 * yy) * np.cos(2 * xx)]).astype(np.float32)[None])
    field = field * gold("decode_field_scale")
    img = (torch.clamp(((field + 1.0) / 2.0).reshape(3, 512, 512).permute(1, 2, 0), 0, 1) * 255).to(torch.uint8)
    diff = (img.float() - gold("decode_uint8")).abs()
    assert diff.max().item() <= 1 and (diff > 0).float().mean().item() < 1e-3   # truncation at exact integers


def test_clip(oracle):
    csd = oracle.make_clip_state_dict()
    P = "cond_stage_model.transformer.text_model"
    h = rnd(751, 1, 77, 768)
    L0 = P + ".encoder.layers.0"
    mlp = oracle.linear(oracle.quick_gelu(oracle.linear(h, csd[L0 + ".mlp.fc1.weight"], csd[L0 + ".mlp.fc1.bias"])),
                        csd[L0 + ".mlp.fc2.weight"], csd[L0 + ".mlp.fc2.bias"])
    assert rel_err(mlp, gold("clip_mlp")) < TOL
    mask = torch.triu(torch.full((1, 1, 77, 77), float("-inf")), diagonal=1)
    assert rel_err(oracle.clip_attention(csd, L0 + ".self_attn", h, mask), gold("clip_attention")) < TOL
    ids = GOLD["clip_ids"].astype(np.int64)
    with torch.no_grad():
        out = oracle.clip_text_transformer(csd, ids)
    assert rel_err(out, gold("clip_text_transformer")) < 1e-4
