"""Generates tests/golden/sampler50_oracle.npz: the oracle's (fp32, CPU) 50-step CFG/DDIM trajectory end point on
the seeded synthetic weights and inputs, at 32x32 and 64x64 latents. Test infrastructure only.

    python oracle/make_sampler_golden.py            # ~10 min on 8 cores

tests/test_sampler50_gpu.py runs the same 50 steps through the CUDA path and reports PSNR against these latents
(north_star: >= 40 dB after 50 steps with a fixed seed; the VAE decode that follows is the same function of both)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_ops as R  # noqa: E402


def run(hw, steps=50, guidance=7.5):
    sd = R.make_unet_state_dict(seed=1234)
    lat, unc, ctx = R.make_inputs(1, hw)
    ts, alphas, alphas_prev = R.sampler_schedule(steps)
    x = lat
    t0 = time.time()
    with torch.no_grad():
        for i in reversed(range(len(ts))):
            x = R.sampler_step(sd, unc, ctx, x, [ts[i]], alphas[[i]], alphas_prev[[i]], guidance)
            print(f"hw={hw} step {len(ts) - i}/{len(ts)} {time.time() - t0:.0f}s |x|max={x.abs().max():.3f}", flush=True)
    return x.numpy()


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    out = {}
    for hw in (32, 64):
        out[f"latent_{hw}"] = run(hw)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sampler50_oracle.npz"), **out)
    print("saved")
