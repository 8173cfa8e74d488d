"""GPU box: the reference's arithmetic through the vendor libraries it uses (cuDNN conv / cuBLAS GEMM / elementwise), i.e.
the oracle's torch ops executed on CUDA, for BASELINE.json configs[1] (one CFG UNet step, 64x64 latent, batch 2).

The reference itself (CuPy + cudnn-frontend python + ctypes cuBLAS) does not install offline, so this is the closest
measurable stand-in for "the reference's cuDNN/cuBLAS path on the same B200" (SURVEY.md §8d): the same operator sequence,
(i) fp32 with TF32 off = the reference's dtype, (ii) fp32 with TF32 on, (iii) fp16 weights/activations; each eager and
replayed from a CUDA graph (no per-call graph builds, no host syncs: kinder than the reference's literal path).
Prints one JSON object; used for DESIGN.md, never by the product path."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_ops as R  # noqa: E402


def bench(sd, lat2, ctx2, steps=10):
    # the timestep embedding is formed on the host in the oracle (numpy fp64): precompute it so a graph can be captured
    temb = R.timestep_embedding([501], 320).to(device=lat2.device)
    R.timestep_embedding = lambda timesteps, dim, max_period=10000: temb

    def step():
        with torch.no_grad():
            return R.unet_forward(sd, lat2, [501], ctx2, quirks=True)
    for _ in range(3):
        out = step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        out = step()
    b.record()
    torch.cuda.synchronize()
    eager = a.elapsed_time(b) / steps
    graph_ms = None
    try:
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            step()
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                out = step()
        for _ in range(3):
            g.replay()
        a.record()
        for _ in range(steps):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        graph_ms = a.elapsed_time(b) / steps
    except Exception as exc:  # capture may be refused by an op; eager number still stands
        graph_ms = f"capture failed: {type(exc).__name__}"
    return eager, graph_ms, bool(torch.isfinite(out.float()).all())


def main():
    dev = torch.device("cuda:0")
    sd32 = {k: v.to(dev) for k, v in R.make_unet_state_dict(seed=1234).items()}
    lat, unc, ctx = R.make_inputs(1, 64)
    lat2, ctx2 = torch.cat([lat, lat]).to(dev), torch.cat([unc, ctx]).to(dev)
    res = {"config": "one UNet forward at batch 2 (CFG pair), 64x64 latent; CFG combine + DDIM excluded (negligible)",
           "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(), "gpu": torch.cuda.get_device_name(0)}
    for name, tf32, half in (("fp32_tf32_off", False, False), ("fp32_tf32_on", True, False), ("fp16", True, True)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = True
        if half:
            sd = {k: v.half() for k, v in sd32.items()}
            l2, c2 = lat2.half(), ctx2.half()
        else:
            sd, l2, c2 = sd32, lat2, ctx2
        try:
            eager, graph, finite = bench(sd, l2, c2)
            res[name] = {"eager_ms": eager, "graph_ms": graph, "finite": finite,
                         "steps_per_s_best": 1000.0 / (graph if isinstance(graph, float) else eager)}
        except Exception as exc:
            res[name] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
