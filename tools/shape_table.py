"""Dev (GPU box): in-graph cost of every unique kernel call of one UNet CFG step.

The calls of one eager step are recorded (Context.prof keeps each call's closure; the arena addresses stay valid),
then every unique (kind, shape) is re-issued as a captured chain of 32 dependent launches and timed with CUDA events
around graph replays - the same conditions as the captured step (PDL overlap of prologues, no event overhead).
Prints n x us per shape, the achieved TFLOP/s and the share of the summed time."""
import collections, contextlib, io, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinyfusers_b200 import synthetic as R
from tinyfusers_b200.storage.state import update_state
from tinyfusers_b200.variants.sd import StableDiffusion
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
HW = int(sys.argv[2]) if len(sys.argv) > 2 else 64
sd = R.make_unet_state_dict()
m = StableDiffusion()
with contextlib.redirect_stdout(io.StringIO()):
    update_state(m, sd)
lat, unc, ctx = R.make_inputs(B, HW)
s = m._sampler(lat.shape, 77)
s.load(unc.cuda(), ctx.cuda(), lat.cuda()); s.set_scalars(501, 0.5, 0.6, 7.5)
for _ in range(2): s.enqueue_step()
torch.cuda.synchronize()
eng = s.unet_engine
eng.ctx.prof = []
s.enqueue_step()
torch.cuda.synchronize()
calls = eng.ctx.prof
eng.ctx.prof = None
uniq = collections.OrderedDict()
for key, _e0, _e1, fn in calls:
    u = uniq.setdefault(key, [0, fn]); u[0] += 1
N = 32
def chain(fn):
    g = torch.cuda.CUDAGraph(); side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(N): fn()
    for _ in range(2): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000 / (5 * N)
flop = lambda k: (2.0 * k[1] * k[2] * k[3] if k[0] == "gemm" else 2.0 * k[1] * (k[2] // k[6]) * (k[3] // k[6]) * k[5] * 9 * k[4] if k[0] == "conv3x3" else 4.0 * k[1] * k[2] * k[3] * k[4] * k[5] if k[0] == "attention" else 2.0 * k[1] * 4 * k[2] * k[3] * k[5] * 9 * k[4] if k[0] == "conv_up2x" else 0)   # conv_up2x: the reference's multiply-adds (3x3 over the 4x image)
rows = []
for key, (n, fn) in uniq.items():
    rows.append((key, n, chain(fn)))
tot = sum(n * us for _, n, us in rows)
print(f"batch {B} latent {HW}: sum of n x in-graph chain cost = {tot:.0f} us over {sum(n for _, n, _ in rows)} calls, {len(rows)} unique shapes")
bykind = collections.Counter()
for key, n, us in sorted(rows, key=lambda r: -r[1] * r[2]):
    f = flop(key); bykind[key[0]] += n * us
    print(f"{str(key):62s} n={n:3d} us={us:7.2f} total={n*us:8.1f} ({100*n*us/tot:4.1f}%)" + (f"  {f/us/1e6:7.1f} TFLOP/s" if f else ""))
print(dict(bykind))
