"""Dev (GPU box): per-call device time of one eager UNet CFG step, aggregated by kernel shape (CUDA events)."""
import collections, contextlib, io, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinyfusers_b200 import synthetic as R
from tinyfusers_b200.storage.state import update_state
from tinyfusers_b200.variants.sd import StableDiffusion
sd = R.make_unet_state_dict()
m = StableDiffusion()
with contextlib.redirect_stdout(io.StringIO()):
    update_state(m, sd)
lat, unc, ctx = R.make_inputs(1, 64)
s = m._sampler(lat.shape, 77)
s.load(unc.cuda(), ctx.cuda(), lat.cuda()); s.set_scalars(501, 0.5, 0.6, 7.5)
for _ in range(2): s.enqueue_step()
torch.cuda.synchronize()
eng = s.unet_engine
eng.ctx.prof = []
s.enqueue_step()
torch.cuda.synchronize()
agg = collections.OrderedDict()
for key, e0, e1, _fn in eng.ctx.prof:
    a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += e0.elapsed_time(e1) * 1000
eng.ctx.prof = None
tot = sum(v[1] for v in agg.values())
flop = lambda k: (2.0 * k[1] * k[2] * k[3] if k[0] == "gemm" else 2.0 * k[1] * (k[2] // k[6]) * (k[3] // k[6]) * k[5] * 9 * k[4] if k[0] == "conv3x3" else 4.0 * k[1] * k[2] * k[3] * k[4] * k[5] if k[0] == "attention" else 0)
print(f"total {tot:.0f} us over {sum(v[0] for v in agg.values())} calls (eager, event-bracketed: includes ~2-3 us event overhead per call)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    f = flop(k)
    print(f"{str(k):62s} n={v[0]:3d} total={v[1]:8.1f} us avg={v[1]/v[0]:7.1f} us" + (f"  {f*v[0]/v[1]/1e6:7.1f} TFLOP/s" if f else ""))
