import os, sys, io, contextlib, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_ops as R
from tinyfusers_b200.storage.state import update_state
from tinyfusers_b200.variants.sd import StableDiffusion
hw = int(sys.argv[1]) if len(sys.argv) > 1 else 16
sd = R.make_unet_state_dict()
m = StableDiffusion()
with contextlib.redirect_stdout(io.StringIO()):
    update_state(m, sd)
lat, unc, ctx = R.make_inputs(1, hw)
x2, c2 = torch.cat([lat, lat]), torch.cat([unc, ctx])
out = m.model.diffusion_model(x2.cuda(), torch.tensor([981]).cuda(), c2.cuda())
torch.cuda.synchronize()
with torch.no_grad():
    ref = R.unet_forward(sd, x2, [981], c2, quirks=True)
print("rel_err", ((out.cpu() - ref).abs().max() / ref.abs().max()).item())
