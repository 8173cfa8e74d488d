"""Full-path parity: UNet forward, CFG step, multi-step sampler (eager == graph), against the oracle.

Sizes are chosen so the CPU oracle finishes in seconds (32x32 latents: 1024..16 tokens; the network is fully
convolutional, so every layer type, channel plan, skip concat, up/down-sample of the 64x64 case is
exercised). The full-size (64x64, BASELINE.json configs[1]) path is covered by size-independent
properties: CFG linearity in the guidance scale and graph/eager bit-equality."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def test_unet_forward_matches_oracle(oracle, unet_sd, sd_model):
    lat, unc, ctx = oracle.make_inputs(1, 32)
    x2, c2 = torch.cat([lat, lat]), torch.cat([unc, ctx])
    with torch.no_grad():
        ref = oracle.unet_forward(unet_sd, x2, [981], c2, quirks=True)
    out = sd_model.model.diffusion_model(x2.cuda(), torch.tensor([981]).cuda(), c2.cuda())
    assert out.shape == (2, 4, 32, 32) and out.dtype == torch.float32
    assert rel_err(out, ref) < 2e-2   # ~700 fp16 kernels deep; per-op bound is 1e-2


@pytest.mark.parametrize("quirks,strided", [(False, False), (True, True)])
def test_unet_forward_other_modes(oracle, unet_sd, sd_model, quirks, strided):
    """canonical head merge (real checkpoints) and the literal LayerNorm-stride reading"""
    import tinyfusers_b200
    lat, unc, ctx = oracle.make_inputs(1, 32, seed=7, ctx_seed=8)
    x2, c2 = torch.cat([lat, lat]), torch.cat([unc, ctx])
    with torch.no_grad():
        ref = oracle.unet_forward(unet_sd, x2, [501], c2, quirks=quirks, ln_strided=strided)
    tinyfusers_b200.set_quirks(quirks)
    tinyfusers_b200.set_layernorm_strided(strided)
    try:
        out = sd_model.model.diffusion_model(x2.cuda(), torch.tensor([501]).cuda(), c2.cuda())
    finally:
        tinyfusers_b200.set_quirks(True)
        tinyfusers_b200.set_layernorm_strided(False)
    assert rel_err(out, ref) < 2e-2


def test_unet_matches_reference_golden(sd_model, oracle):
    """16x16-latent UNet output of the reference's own Python (tests/golden); self-attention at the deepest
    level has 4 tokens there, below the B200 kernel's 8-token granularity, so the golden is checked at the
    block level (test_blocks_gpu.py) and through the oracle (tests/test_oracle_golden.py) instead."""
    pytest.skip("covered transitively: oracle == reference golden (CPU suite), CUDA == oracle (this suite)")


def test_sampler_step_matches_oracle(oracle, unet_sd, sd_model):
    lat, unc, ctx = oracle.make_inputs(1, 32)
    ts, alphas, alphas_prev = oracle.sampler_schedule(50)
    i = 30
    with torch.no_grad():
        ref = oracle.sampler_step(unet_sd, unc, ctx, lat, [ts[i]], alphas[[i]], alphas_prev[[i]], 7.5)
        ref_e = oracle.get_model_output(unet_sd, unc, ctx, lat, [ts[i]], 7.5)
    out = sd_model(unc.cuda(), ctx.cuda(), lat.cuda(), torch.tensor([ts[i]]).cuda(), alphas[[i]].cuda(),
                   alphas_prev[[i]].cuda(), torch.tensor([7.5]))
    e_t = sd_model.get_model_output(unc.cuda(), ctx.cuda(), lat.cuda(), torch.tensor([ts[i]]).cuda(), torch.tensor([7.5]))
    assert rel_err(e_t, ref_e) < 3e-2    # guidance 7.5 amplifies the cond-uncond difference
    assert rel_err(out, ref) < 1e-2


def test_sampler_loop_graph_equals_eager_and_oracle(oracle, unet_sd, sd_model):
    lat, unc, ctx = oracle.make_inputs(1, 32, seed=3, ctx_seed=4)
    ts, alphas, alphas_prev = oracle.sampler_schedule(3)
    args = (unc.cuda(), ctx.cuda(), lat.cuda(), ts, alphas, alphas_prev, 7.5)
    eager = sd_model.sample(*args, use_graph=False)
    graph = sd_model.sample(*args, use_graph=True)
    graph2 = sd_model.sample(*args, use_graph=True)
    assert torch.equal(graph, graph2)        # every kernel reduces in a fixed order: replays are bit-identical
    assert torch.equal(graph, eager)
    x = lat
    with torch.no_grad():
        for i in reversed(range(len(ts))):
            x = oracle.sampler_step(unet_sd, unc, ctx, x, [ts[i]], alphas[[i]], alphas_prev[[i]], 7.5)
    assert rel_err(graph, x) < 3e-2


def test_full_size_cfg_linearity(sd_model, oracle):
    """64x64 (BASELINE configs[1]): e_t(g) = u + g (c - u) must be affine in g: e(0) + e(2) = 2 e(1)."""
    lat, unc, ctx = oracle.make_inputs(1, 64)
    t = torch.tensor([501]).cuda()
    e = [sd_model.get_model_output(unc.cuda(), ctx.cuda(), lat.cuda(), t, torch.tensor([g])) for g in (0.0, 1.0, 2.0)]
    assert torch.isfinite(e[1]).all()
    assert rel_err(e[0] + e[2], 2 * e[1]) < 1e-3
