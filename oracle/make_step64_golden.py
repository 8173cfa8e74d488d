"""Generates tests/golden/step64_oracle.npz: ONE UNet forward and ONE CFG + DDIM step of the oracle (fp32, CPU) at the benched
size - BASELINE.json configs[1]: 64x64 latent, batch 2 ([uncond ; cond]), 77-token context - on the seeded synthetic weights
and inputs. Test infrastructure only.

    python oracle/make_step64_golden.py            # ~1 min on 8 cores

tests/test_unet_gpu.py holds the CUDA path to these tensors element-wise (global AND per-channel relative error), so the
headline configuration is pinned by an oracle tensor and not only by the 50-step PSNR and the CFG-linearity property."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_ops as R  # noqa: E402

if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    sd = R.make_unet_state_dict(seed=1234)
    lat, unc, ctx = R.make_inputs(1, 64)
    ts, alphas, alphas_prev = R.sampler_schedule(50)
    i = 30
    with torch.no_grad():
        x2, c2 = torch.cat([lat, lat]), torch.cat([unc, ctx])
        unet_out = R.unet_forward(sd, x2, [ts[i]], c2, quirks=True)                       # (2,4,64,64)
        e_t = unet_out[:1] + 7.5 * (unet_out[1:] - unet_out[:1])                           # variants/sd.py:44-45
        x_prev, _ = R.get_x_prev_and_pred_x0(lat, e_t, alphas[[i]], alphas_prev[[i]])
        # the same forward with the canonical head merge (what real checkpoints need)
        unet_canon = R.unet_forward(sd, x2, [ts[i]], c2, quirks=False)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "step64_oracle.npz"), step_index=np.int64(i), timestep=np.int64(ts[i]),
                        unet_out=unet_out.numpy(), e_t=e_t.numpy(), x_prev=x_prev.numpy(), unet_out_canonical=unet_canon.numpy())
    print("saved", float(unet_out.abs().max()), float(x_prev.abs().max()))
