import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def rel_err(a, b):
    """max|a-b| / max|b| — the per-op metric of BASELINE.json's north_star (<= 1e-2 in fp16)."""
    import torch
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def rel_err_per_channel(a, b):
    """The same metric per (image, channel) plane of an NCHW tensor, worst plane: max_c [ max|a_c - b_c| / max|b_c| ]. The
    global metric normalises by the largest value of the whole tensor, so a wrong low-magnitude channel would pass it."""
    import torch
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert a.shape == b.shape and a.dim() == 4
    num = (a - b).abs().amax(dim=(2, 3))
    den = b.abs().amax(dim=(2, 3)).clamp_min(1e-12)
    return (num / den).max().item()


@pytest.fixture(scope="session")
def oracle():
    from oracle import ref_ops
    return ref_ops


@pytest.fixture(scope="session")
def unet_sd(oracle):
    """Seeded synthetic UNet weights with the reference's checkpoint keys (fp32, CPU)."""
    return oracle.make_unet_state_dict(seed=1234)


@pytest.fixture(scope="session")
def sd_model(unet_sd):
    """tinyfusers_b200 StableDiffusion with the synthetic weights injected through update_state."""
    import contextlib
    import io
    import torch
    from tinyfusers_b200.storage.state import update_state
    from tinyfusers_b200.variants.sd import StableDiffusion
    assert torch.cuda.is_available()
    m = StableDiffusion()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        update_state(m, unet_sd)
    skipped = [l for l in buf.getvalue().splitlines() if l.startswith("skipped")]
    # the only keys the synthetic dict does not carry are the bias-less attention projections
    # (and the VAE / CLIP weights, which have their own synthetic state dicts: the `vae_clip_model` fixture)
    unet_skipped = [k for k in skipped if "model.diffusion_model" in k]
    assert all(k.endswith((".to_q.bias", ".to_k.bias", ".to_v.bias")) for k in unet_skipped), unet_skipped[:5]
    return m
