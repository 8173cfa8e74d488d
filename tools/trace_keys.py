"""Dev (GPU box): ordered list of kernel-call keys of one UNet CFG step (to join with an ncu launch list)."""
import contextlib, io, json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinyfusers_b200.variants.sd import StableDiffusion
from tinyfusers_b200 import runtime
m = StableDiffusion()
s = m._sampler((1, 4, 64, 64), 77)
eng = s.unet_engine
keys = []
eng.ctx._timed = lambda key, fn: (keys.append(key), fn())[1]
s.enqueue_step()
torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(keys, open("gpurun_out/keys.json", "w"))
print(len(keys), "keys")
