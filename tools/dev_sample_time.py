"""Dev (GPU box): device and wall time of StableDiffusion.sample (50 steps, 64x64) vs the bare graph-replay loop."""
import contextlib, io, os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinyfusers_b200 import synthetic as SY
from tinyfusers_b200.native.b200.ops import b200
from tinyfusers_b200.storage.state import update_state
from tinyfusers_b200.variants.sd import StableDiffusion
b200.init(0)
m = StableDiffusion()
with contextlib.redirect_stdout(io.StringIO()):
    update_state(m, SY.make_unet_state_dict(seed=1234))
lat, unc, ctx = SY.make_inputs(1, 64)
lat, unc, ctx = lat.cuda(), unc.cuda(), ctx.cuda()
ts, al, ap = SY.sampler_schedule(50)
for i in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); x = m.sample(unc, ctx, lat, ts, al, ap, 7.5); e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"sample(): device {e0.elapsed_time(e1):.2f} ms, host enqueue {1000*(t1-t0):.2f} ms, wall {1000*(t2-t0):.2f} ms", flush=True)
s = m._sampler(lat.shape, 77)
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s.load(unc, ctx, lat); s.set_tables(ts, al, ap, 7.5); s.run(50); e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"load+tables+run: device {e0.elapsed_time(e1):.2f} ms, host enqueue {1000*(t1-t0):.2f} ms, wall {1000*(t2-t0):.2f} ms", flush=True)
