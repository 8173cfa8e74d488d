"""fp32 parity mode for the rows next to the hot path: VAE decoder and CLIP text encoder.

Held to the fp64 evaluation of the oracle at 1e-5 per block (deep chains: the bound at the assert) AND to the golden
vectors the reference's own Python produced in fp32 (tests/golden/reference_outputs_vae_clip.npz, made by
oracle/make_golden_vae.py) at the fp32 reduction-order noise of those vectors (2e-5 blocks / 1e-4 whole decoder, CLIP:
the same bounds tests/test_oracle_golden_vae.py gives the oracle)."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from test_fp32_gpu import _d, _sd64, fp32_mode  # noqa: F401  (autouse fixture: fp32 mode on for every test here)
from test_vae_clip_gpu import GOLD, _load, gold, rnd

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vae(oracle):
    from tinyfusers_b200.vae.vae import AutoencoderKL
    sd = oracle.make_vae_decoder_state_dict()
    return _load(AutoencoderKL(), sd, "first_stage_model"), sd


@pytest.fixture(scope="module")
def clip(oracle):
    from tinyfusers_b200.vae.encoder import CLIPTextTransformer
    sd = oracle.make_clip_state_dict()
    return _load(CLIPTextTransformer(), sd, "cond_stage_model.transformer.text_model"), sd


def test_resnet_block_fp32(oracle):
    from tinyfusers_b200.vision.resnet import ResnetBlock
    sd = {}
    oracle.add_resnet_block(sd, "rn", 64, 128, seed=701)
    x = rnd(703, 1, 64, 8, 8, scale=1.3, shift=0.2)
    y = _load(ResnetBlock(64, 128), sd, "rn")(x.cuda())
    assert rel_err(y, gold("resnet_block_64_128")) < 2e-5
    with torch.no_grad():
        assert rel_err(y, oracle.resnet_block(_sd64(sd), "rn", _d(x))) < 1e-5
    sd = {}
    oracle.add_resnet_block(sd, "r", 256, 128, seed=9)
    x = rnd(10, 2, 256, 64, 64)
    with torch.no_grad():
        ref = oracle.resnet_block(_sd64(sd), "r", _d(x))
    assert rel_err(_load(ResnetBlock(256, 128), sd, "r")(x.cuda()), ref) < 1e-5


@pytest.mark.parametrize("quirks", [True, False])
def test_attn_block_fp32(oracle, quirks):
    """quirks: the reference's per-plane reading (golden); canonical: one 512-wide head over H*W pixels - built in fp32 mode."""
    import tinyfusers_b200
    from tinyfusers_b200.attention.attention import AttnBlock
    tinyfusers_b200.set_quirks(quirks)
    try:
        sd = {}
        oracle.add_attn_block(sd, "ab", 64, seed=711)
        ab = _load(AttnBlock(64), sd, "ab")
        x = rnd(712, 1, 64, 6, 10)
        y = ab(x.cuda())
        if quirks:
            assert rel_err(y, gold("attn_block_6x10")) < 2e-5
        with torch.no_grad():
            assert rel_err(y, oracle.attn_block(_sd64(sd), "ab", _d(x), quirks)) < 1e-5
        sd = {}
        oracle.add_attn_block(sd, "a", 512, seed=3)         # the VAE's own geometry at a 32x32 plane, batch 2
        x = rnd(4, 2, 512, 32, 32)
        with torch.no_grad():
            ref = oracle.attn_block(_sd64(sd), "a", _d(x), quirks)
        assert rel_err(_load(AttnBlock(512), sd, "a")(x.cuda()), ref) < 1e-5
    finally:
        tinyfusers_b200.set_quirks(True)


def test_mid_and_decoder_fp32(vae, oracle):
    from tinyfusers_b200.vae.mid import Mid
    sd = {}
    oracle.add_resnet_block(sd, "mid.block_1", 64, 64, seed=721)
    oracle.add_attn_block(sd, "mid.attn_1", 64, seed=721)
    oracle.add_resnet_block(sd, "mid.block_2", 64, 64, seed=721)
    assert rel_err(_load(Mid(64), sd, "mid")(rnd(722, 1, 64, 8, 8).cuda()), gold("mid_64")) < 2e-5
    m, vsd = vae
    y = m.decoder(rnd(731, 1, 4, 4, 4).cuda())
    assert y.shape == (1, 3, 32, 32)
    assert rel_err(y, gold("decoder_4x4")) < 1e-4          # 30 convolutions deep; the golden itself is fp32
    z = rnd(56, 1, 4, 16, 16)
    with torch.no_grad():
        ref = oracle.vae_decoder(_sd64(vsd), "first_stage_model.decoder", _d(z))
    assert rel_err(m.decoder(z.cuda()), ref) < 5e-5


def test_decode_uint8_fp32(vae, oracle):
    from tinyfusers_b200.variants.sd import StableDiffusion
    m, sd = vae
    model = StableDiffusion.__new__(StableDiffusion)
    model.first_stage_model = m
    z = 0.18215 * rnd(51, 1, 4, 16, 16)
    img = model.decode(z.cuda())
    assert img.shape == (128, 128, 3) and img.dtype == torch.uint8
    with torch.no_grad():
        ref = oracle.vae_decode(sd, z)
    d = (img.cpu().int() - ref.int()).abs()
    # fp32 on both sides: only pixels whose float value sits within rounding noise of an integer level may differ, by 1
    assert d.max().item() <= 1 and (d > 0).float().mean().item() < 5e-3


def test_clip_fp32(clip, oracle):
    m, sd = clip
    h = rnd(751, 1, 77, 768)
    assert rel_err(m.encoder.layers[0].mlp(h.cuda()), gold("clip_mlp")) < 2e-5
    assert rel_err(m.encoder.layers[0].self_attn(h.cuda(), None), gold("clip_attention")) < 2e-5
    ids = GOLD["clip_ids"].astype(np.int64)
    out = m(ids)
    assert out.shape == (1, 77, 768) and out.dtype == torch.float32
    assert rel_err(out, gold("clip_text_transformer")) < 1e-4
    with torch.no_grad():
        ref = oracle.clip_text_transformer(_sd64(sd), ids)
    assert rel_err(out, ref) < 2e-5
