"""Measured errors of the fp32 parity mode (and of the fp16 path against it) -> profiles/fp32_parity_r1.json.
Run on the GPU box:  python tools/fp32_report.py  (imports oracle/: a checker tool, like the tests)."""
import contextlib
import io
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import tinyfusers_b200  # noqa: E402
from conftest import rel_err  # noqa: E402
from oracle import ref_ops as oracle  # noqa: E402
from tinyfusers_b200.attention.attention import SpatialTransformer  # noqa: E402
from tinyfusers_b200.attention.sdpa import scaled_dot_product_attention  # noqa: E402
from tinyfusers_b200.ff.group_norm import group_norm  # noqa: E402
from tinyfusers_b200.ff.linear import Linear  # noqa: E402
from tinyfusers_b200.storage.state import update_state  # noqa: E402
from tinyfusers_b200.variants.sd import StableDiffusion  # noqa: E402
from tinyfusers_b200.vision.conv2d import conv_2d  # noqa: E402
from tinyfusers_b200.vision.resnet import ResBlock  # noqa: E402


def d(t):
    return t.double() if isinstance(t, torch.Tensor) and t.is_floating_point() else t


def load(obj, sd, prefix=None):
    with contextlib.redirect_stdout(io.StringIO()):
        update_state(obj, sd, prefix) if prefix else update_state(obj, sd)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return out, (time.perf_counter() - t0) / reps * 1e3


def main():
    rep = {}
    g = torch.Generator().manual_seed(1)
    for mode in ("fp32", "fp16"):
        tinyfusers_b200.set_precision(mode)
        r = {}
        x, w = torch.randn(1, 320, 64, 64, generator=g), torch.randn(320, 320, 3, 3, generator=g) / 53.7
        r["conv3x3_320_64x64"] = rel_err(conv_2d(x.cuda(), w.cuda(), [1, 1], [1, 1], [1, 1]), oracle.conv2d(d(x), d(w), None, (1, 1), (1, 1)))
        lin = Linear(320, 2560)
        lin.weight, lin.bias = (torch.randn(2560, 320, generator=g) / 17.9).cuda(), torch.randn(2560, generator=g).cuda()
        xl = torch.randn(1, 4096, 320, generator=g)
        r["linear_4096x320x2560"] = rel_err(lin(xl.cuda()), oracle.linear(d(xl), d(lin.weight.cpu()), d(lin.bias.cpu())))
        r["group_norm_320_64x64"] = rel_err(group_norm(x.cuda(), 32, 1e-5), oracle.group_norm(d(x), 32, 1e-5))
        q, k, v = (torch.randn(1, 8, 4096, 40, generator=g) for _ in range(3))
        r["sdpa_8x4096x4096x40"] = rel_err(scaled_dot_product_attention(q.cuda(), k.cuda(), v.cuda()),
                                           oracle.scaled_dot_product_attention(d(q), d(k), d(v)))
        sd = {}
        oracle.add_res_block(sd, "rb", 320, 320, seed=31)
        oracle.add_spatial_transformer(sd, "st", 320, 768, seed=31)
        gg = torch.Generator().manual_seed(31)
        x1, emb, ctx = torch.randn(1, 320, 64, 64, generator=gg), torch.randn(1, 1280, generator=gg), torch.randn(1, 77, 768, generator=gg)
        rb, st = ResBlock(320, 1280, 320), SpatialTransformer(320, 768, 8, 40)
        load(rb, sd, "rb")
        load(st, sd, "st")
        y, ms = timed(lambda: st(rb(x1.cuda(), emb.cuda()), ctx.cuda()))
        sd64 = {kk: d(vv) for kk, vv in sd.items()}
        with torch.no_grad():
            ref = oracle.spatial_transformer(sd64, "st", oracle.res_block(sd64, "rb", d(x1), d(emb)), d(ctx), 8, 40, True)
        r["C1_down_block0_64x64"] = rel_err(y, ref)
        r["C1_down_block0_64x64_ms_eager_incl_layout"] = ms
        rep[mode + "_vs_fp64_oracle"] = r
    # full step: fp16 path against the fp32 path, full sizes
    m = StableDiffusion()
    load(m, oracle.make_unet_state_dict(seed=1234))
    full = {}
    for hw in (64, 96):
        lat, unc, ctx = oracle.make_inputs(1, hw, seed=11, ctx_seed=12)
        args = (unc.cuda(), ctx.cuda(), lat.cuda(), torch.tensor([981]).cuda(), torch.tensor([7.5]))
        tinyfusers_b200.set_precision("fp32")
        e32, ms32 = timed(lambda: m.get_model_output(*args), reps=1)
        tinyfusers_b200.set_precision("fp16")
        e16, ms16 = timed(lambda: m.get_model_output(*args), reps=1)
        full[f"cfg_step_{hw}x{hw}"] = {"fp16_vs_fp32_mode": rel_err(e16, e32), "fp32_mode_ms": ms32, "fp16_eager_ms": ms16}
    rep["full_size_step"] = full
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "fp32_parity_r1.json"), "w") as fh:
        json.dump(rep, fh, indent=1)
    print(json.dumps(rep, indent=1))


if __name__ == "__main__":
    main()
