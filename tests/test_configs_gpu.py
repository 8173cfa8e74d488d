"""BASELINE.json configs beyond configs[1], at their real sizes:
  C1  down-block-0 (ResBlock 320 + SpatialTransformer 320/8x40) at the full 64x64 latent (4096-token self-attention),
      against the oracle;
  C4  8 images per GPU with CFG (effective batch 16) at 64x64: every image of the batch must equal the same image run
      alone (the path shards by image, so this is the property the data-parallel split relies on);
  C5  96x96 latent (9216-token self-attention, a 12x12 level whose 144 pixels are not a multiple of the 32-row statistics
      slots -> the 3-launch GroupNorm path): down-block-0 against the oracle, full UNet through CFG linearity."""
import contextlib
import io

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _load(obj, sd, prefix):
    from tinyfusers_b200.storage.state import update_state
    with contextlib.redirect_stdout(io.StringIO()):
        update_state(obj, sd, prefix)


@pytest.mark.parametrize("hw", [64, 96])
def test_down_block0_full_size(oracle, hw):
    from tinyfusers_b200.attention.attention import SpatialTransformer
    from tinyfusers_b200.vision.resnet import ResBlock
    sd = {}
    oracle.add_res_block(sd, "rb", 320, 320, seed=31)
    oracle.add_spatial_transformer(sd, "st", 320, 768, seed=31)
    g = torch.Generator().manual_seed(31)
    x = torch.randn(1, 320, hw, hw, generator=g)
    emb = torch.randn(1, 1280, generator=g)
    ctx = torch.randn(1, 77, 768, generator=g)
    rb, st = ResBlock(320, 1280, 320), SpatialTransformer(320, 768, 8, 40)
    _load(rb, sd, "rb")
    _load(st, sd, "st")
    y = st(rb(x.cuda(), emb.cuda()), ctx.cuda())
    with torch.no_grad():
        ref = oracle.spatial_transformer(sd, "st", oracle.res_block(sd, "rb", x, emb), ctx, 8, 40, True)
    assert rel_err(y, ref) < 1e-2


def test_batch8_images_equal_single_image_runs(sd_model, oracle):
    """C4 per-GPU batch: 8 images x CFG = 16 UNet samples in one step."""
    lat, unc, ctx = oracle.make_inputs(8, 64, seed=5, ctx_seed=6)
    ts, alphas, alphas_prev = oracle.sampler_schedule(50)
    i = 25
    args = (torch.tensor([ts[i]]).cuda(), alphas[[i]].cuda(), alphas_prev[[i]].cuda(), torch.tensor([7.5]))
    full = sd_model(unc.cuda(), ctx.cuda(), lat.cuda(), *args)
    assert full.shape == (8, 4, 64, 64) and torch.isfinite(full).all()
    for k in (0, 5):
        one = sd_model(unc[k:k + 1].cuda(), ctx[k:k + 1].cuda(), lat[k:k + 1].cuda(), *args)
        # different tile / split-K choices at M = 16x4096 vs 2x4096 rows: equal up to fp16 accumulation-order noise
        assert rel_err(full[k:k + 1], one) < 5e-3


def test_latent_96_cfg_linearity(sd_model, oracle):
    """C5 geometry through the whole UNet: e_t(g) affine in g, outputs finite (9216 / 2304 / 576 / 144 tokens)."""
    lat, unc, ctx = oracle.make_inputs(1, 96, seed=9, ctx_seed=10)
    t = torch.tensor([301]).cuda()
    e = [sd_model.get_model_output(unc.cuda(), ctx.cuda(), lat.cuda(), t, torch.tensor([g])) for g in (0.0, 1.0, 2.0)]
    assert e[1].shape == (1, 4, 96, 96) and torch.isfinite(e[1]).all()
    assert rel_err(e[0] + e[2], 2 * e[1]) < 1e-3
