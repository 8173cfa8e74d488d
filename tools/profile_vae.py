"""Dev (GPU box): per-call device time of one VAE decode (64x64 latent -> 512x512) and of one CLIP encode, by kernel shape."""
import collections, contextlib, io, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tinyfusers_b200 import synthetic as R
from tinyfusers_b200.runtime import standalone_context
from tinyfusers_b200.storage.state import update_state
from tinyfusers_b200.vae.encoder import CLIPTextTransformer
from tinyfusers_b200.vae.vae import AutoencoderKL
hw = int(sys.argv[1]) if len(sys.argv) > 1 else 64
vae = AutoencoderKL()
with contextlib.redirect_stdout(io.StringIO()):
    update_state(vae, R.make_vae_decoder_state_dict(), "first_stage_model")
z = torch.randn(1, 4, hw, hw, device="cuda")
for _ in range(2): vae.decoder(z)
torch.cuda.synchronize()
eng = vae.decoder._engine(tuple(z.shape))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): eng.decode_nhwc_f32(z)
e1.record(); torch.cuda.synchronize()
print(f"decode {hw}x{hw} latent: {e0.elapsed_time(e1) / 5:.3f} ms per image (eager launches), arena {eng.arena_bytes / 2**20:.0f} MiB")

def table(prof, title):
    agg = collections.OrderedDict()
    for key, a, b, _fn in prof:
        r = agg.setdefault(key, [0, 0.0]); r[0] += 1; r[1] += a.elapsed_time(b) * 1000
    tot = sum(v[1] for v in agg.values())
    flop = lambda k: (2.0 * k[1] * k[2] * k[3] if k[0] == "gemm" else 2.0 * k[1] * (k[2] // k[6]) * (k[3] // k[6]) * k[5] * 9 * k[4] if k[0] == "conv3x3" else 0)
    print(f"{title}: {tot:.0f} us over {sum(v[0] for v in agg.values())} timed calls (event-bracketed, +2-3 us each)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f = flop(k)
        print(f"  {str(k):58s} n={v[0]:3d} total={v[1]:9.1f} us avg={v[1]/v[0]:8.1f}" + (f"  {f*v[0]/v[1]/1e6:7.1f} TFLOP/s" if f else ""))

eng.ctx.prof = []
eng.decode_nhwc_f32(z)
torch.cuda.synchronize()
table(eng.ctx.prof, "VAE decode"); eng.ctx.prof = None

clip = CLIPTextTransformer()
with contextlib.redirect_stdout(io.StringIO()):
    update_state(clip, R.make_clip_state_dict(), "cond_stage_model.transformer.text_model")
ids = np.array([[49406, 320, 1125, 539] + [49407] * 73])
for _ in range(2): clip(ids)
torch.cuda.synchronize()
e0.record()
for _ in range(5): clip(ids)
e1.record(); torch.cuda.synchronize()
print(f"CLIP encode (77 tokens, 12 layers): {e0.elapsed_time(e1) / 5:.3f} ms (eager launches)")
ctx = standalone_context(); ctx.prof = []
clip(ids); torch.cuda.synchronize()
table(ctx.prof, "CLIP encode"); ctx.prof = None
