"""Dev (GPU box): A/B of the next-layer weight prefetch (tf_weight_prefetch_mode) on the captured CFG step: ms per step of
50-step trajectories with the hints baked into the graph vs without, interleaved, and the final latents compared (bit-equal:
a prefetch only warms L2). argv: [images latent]"""
import contextlib, io, json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tinyfusers_b200
from tinyfusers_b200 import synthetic as SY
from tinyfusers_b200.native.b200.ops import b200
from tinyfusers_b200.storage.state import update_state
from tinyfusers_b200.variants.sd import StableDiffusion
images = int(sys.argv[1]) if len(sys.argv) > 1 else 1
hw = int(sys.argv[2]) if len(sys.argv) > 2 else 64
b200.init(0)
m = StableDiffusion()
with contextlib.redirect_stdout(io.StringIO()):
    update_state(m, SY.make_unet_state_dict(seed=1234))
lat, unc, ctx = SY.make_inputs(images, hw)
lat, unc, ctx = lat.cuda(), unc.cuda(), ctx.cuda()
ts, al, ap = SY.sampler_schedule(50)
s = m._sampler(lat.shape, 77)
def loop():
    s.load(unc, ctx, lat); s.set_tables(ts, al, ap, 7.5); s.run(50)
def timed():
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); loop(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 50
res, outs = {True: [], False: []}, {}
for rep in range(4):
    for on in (False, True):
        tinyfusers_b200.set_weight_prefetch(on)
        if rep == 0:
            loop(); torch.cuda.synchronize()       # capture
        res[on].append(timed())
        outs[on] = s.latent.clone()
import ctypes
rec, hin, byt = ctypes.c_int(), ctypes.c_int(), ctypes.c_longlong()
b200.tf_weight_prefetch_stats(ctypes.byref(rec), ctypes.byref(hin), ctypes.byref(byt))
print("recorded launches", rec.value, "hinted", hin.value, "MB", byt.value / 1e6)
print(json.dumps({"images": images, "latent": hw, "ms_per_step_off": sorted(res[False])[len(res[False]) // 2], "ms_per_step_on": sorted(res[True])[len(res[True]) // 2],
                  "all_off": res[False], "all_on": res[True], "bit_equal": bool(torch.equal(outs[True], outs[False]))}))
