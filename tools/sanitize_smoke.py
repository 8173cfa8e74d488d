"""Dev (GPU box): every kernel family of the step once, at small sizes, for `compute-sanitizer --tool memcheck` (one tool per
gpurun call, /opt/skills/guides/B200_PROFILING.md): ResBlock (GroupNorm statistics from producers, conv3x3 with and without
split-K, skip conv appended along K), SpatialTransformer (LayerNorm, fused QKV GEMM, self- and cross-attention, GEGLU),
Up/Downsample (upsample folded into the conv), the three self-attention kernels, CFG + DDIM. Small sizes: the sanitizer runs
the kernels 10-100x slower. Prints the summed outputs so nothing is optimised away."""
import contextlib, io, math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinyfusers_b200 import synthetic as SY
from tinyfusers_b200.attention.attention import SpatialTransformer, _pad64
from tinyfusers_b200.native.b200.ops import b200
from tinyfusers_b200.storage.state import update_state
from tinyfusers_b200.vision.resnet import ResBlock
from tinyfusers_b200.vision.unet import Downsample, Upsample
b200.init(0)
g = torch.Generator().manual_seed(1)
tot = 0.0
sd = {}
SY.add_res_block(sd, "rb", 320, 320, seed=3)
SY.add_res_block(sd, "rb2", 640, 320, seed=4)
SY.add_spatial_transformer(sd, "st", 320, 768, seed=5)
SY._add_conv(sd, "up.conv", 320, 320, 3, 6)
SY._add_conv(sd, "down.op", 320, 320, 3, 7)
rb, rb2 = ResBlock(320, 1280, 320), ResBlock(640, 1280, 320)
st = SpatialTransformer(320, 768, 8, 40)
up, down = Upsample(320), Downsample(320)
with contextlib.redirect_stdout(io.StringIO()):
    for mod, name in ((rb, "rb"), (rb2, "rb2"), (st, "st"), (up, "up"), (down, "down")):
        update_state(mod, sd, name)
x = torch.randn(2, 320, 16, 16, generator=g).cuda()
x2 = torch.randn(2, 640, 16, 16, generator=g).cuda()
emb = torch.randn(1, 1280, generator=g).cuda()
ctx = torch.randn(2, 77, 768, generator=g).cuda()
h = rb(x, emb); tot += float(h.float().abs().sum()); print('rb', tot, flush=True)
h2 = rb2(x2, emb); tot += float(h2.float().abs().sum()); print('rb2', tot, flush=True)
y = st(h, ctx); tot += float(y.float().abs().sum()); print('st', tot, flush=True)
tot += float(up(y).float().abs().sum()) + float(down(y).float().abs().sum()); print('updown', tot, flush=True)
for ver, (B, NH, T, d) in ((1, (1, 2, 256, 40)), (2, (1, 2, 512, 40)), (3, (1, 2, 256, 40)), (1, (1, 2, 128, 160))):
    dp, dvp = (d + 15) // 16 * 16, _pad64(d)
    Q = torch.zeros(B, T, NH, dp, dtype=torch.half, device="cuda"); Q[..., :d] = torch.randn(B, T, NH, d, generator=g).cuda()
    K = torch.zeros(B, T, NH, dp, dtype=torch.half, device="cuda"); K[..., :d] = torch.randn(B, T, NH, d, generator=g).cuda()
    V = torch.zeros(B, T, NH, dvp, dtype=torch.half, device="cuda"); V[..., :d] = torch.randn(B, T, NH, d, generator=g).cuda()
    out = torch.zeros(B, T, NH, d, dtype=torch.half, device="cuda")
    b200.check(b200.tf_attention_set_variant(ver, -1), "variant")
    b200.check(b200.tf_attention_v_f16(Q.data_ptr(), NH * dp, K.data_ptr(), NH * dp, V.data_ptr(), NH * dvp, out.data_ptr(), T * NH * d, d, NH * d,
                                       B, NH, T, T, T, d, dp, dvp, 1 / math.sqrt(d), 0, torch.cuda.current_stream().cuda_stream), "attn")
    tot += float(out.float().abs().sum()); print('attn', ver, tot, flush=True)
b200.tf_attention_set_variant(0, -1)
eps = torch.randn(2, 256, 16, generator=g).cuda(); lat = torch.randn(1, 4, 16, 16, generator=g).cuda()
a_tab, ap_tab, idx = torch.tensor([0.5]).cuda(), torch.tensor([0.6]).cuda(), torch.zeros(1, dtype=torch.int32).cuda()
lo, eo = torch.empty_like(lat), torch.empty_like(lat)
b200.check(b200.tf_cfg_ddim_step_f32(eps.data_ptr(), 16, lat.data_ptr(), lo.data_ptr(), eo.data_ptr(), a_tab.data_ptr(), ap_tab.data_ptr(), idx.data_ptr(),
                                     7.5, 1, 4, 256, torch.cuda.current_stream().cuda_stream), "cfg")
torch.cuda.synchronize()
tot += float(lo.abs().sum())
assert math.isfinite(tot), tot
print("sanitize_smoke ok", tot)
