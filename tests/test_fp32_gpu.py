"""fp32 parity mode (tinyfusers_b200.set_precision("fp32"); csrc/tf_fp32.cu) against the oracle.

North star: per-op max relative error <= 1e-5 in fp32 mode; BASELINE.json configs[0] = down-block-0 (ResBlock 320 +
SpatialTransformer 320 / 8 x 40) at the 64x64 latent, batch 1, fp32. Every operator of SURVEY.md §8a is held to 1e-5
at the UNet's real layer sizes; the chained blocks (10-40 operators deep) to the tolerances written at each assert.
The reference value is the oracle evaluated in fp64 on the same fp32 inputs and weights (the exact value of the
reference's arithmetic); the fp32 oracle itself is asserted to sit inside the same bound, so the two readings agree.
"""
import contextlib
import io

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-5


@pytest.fixture(autouse=True)
def fp32_mode():
    import tinyfusers_b200
    tinyfusers_b200.set_precision("fp32")
    yield
    tinyfusers_b200.set_precision("fp16")


def _load(obj, sd, prefix):
    from tinyfusers_b200.storage.state import update_state
    with contextlib.redirect_stdout(io.StringIO()):
        update_state(obj, sd, prefix)


def _d(t):
    return t.double() if isinstance(t, torch.Tensor) and t.is_floating_point() else t


def _sd64(sd):
    return {k: _d(v) for k, v in sd.items()}


def _check(out, ref64, ref32=None, tol=TOL):
    assert out.dtype == torch.float32 and out.is_cuda
    assert tuple(out.shape) == tuple(ref64.shape)
    e = rel_err(out, ref64)
    assert e < tol, f"rel err {e:.3e} >= {tol:g}"
    if ref32 is not None:
        assert rel_err(ref32, ref64) < tol


def _randn(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


# ---- operators (§8a rows 10-12, 16-20) ------------------------------------------------------------------------------

@pytest.mark.parametrize("cin,cout,hw,k,stride", [(320, 320, 64, 3, 1), (320, 320, 64, 3, 2), (960, 640, 32, 3, 1),
                                                   (640, 320, 64, 1, 1), (4, 320, 64, 3, 1), (320, 4, 64, 3, 1),
                                                   (24, 40, 19, 3, 1)])
def test_conv_2d(oracle, cin, cout, hw, k, stride):
    from tinyfusers_b200.vision.conv2d import conv_2d
    x = _randn(2 if hw < 64 else 1, cin, hw, hw, seed=1)
    w = _randn(cout, cin, k, k, seed=2, scale=(cin * k * k) ** -0.5)
    p = k // 2
    out = conv_2d(x.cuda(), w.cuda(), [p, p], [stride, stride], [1, 1])
    ref64 = oracle.conv2d(_d(x), _d(w), None, (stride, stride), (p, p))
    _check(out, ref64, oracle.conv2d(x, w, None, (stride, stride), (p, p)))


def test_conv2d_module_bias(oracle):
    from tinyfusers_b200.vision.conv2d import Conv2d
    m = Conv2d(320, 640, kernel_size=[3, 3], stride=[2, 2], padding=[1, 1])
    m.weight, m.bias = _randn(640, 320, 3, 3, seed=3, scale=0.02).cuda(), _randn(640, seed=4).cuda()
    x = _randn(2, 320, 32, 32, seed=5)
    _check(m(x.cuda()), oracle.conv2d(_d(x), _d(m.weight.cpu()), _d(m.bias.cpu()), (2, 2), (1, 1)))


@pytest.mark.parametrize("M,K,N", [(4096, 320, 2560), (77, 768, 320), (1, 1280, 1280), (1000, 1283, 331)])
def test_linear(oracle, M, K, N):
    from tinyfusers_b200.ff.linear import Linear
    m = Linear(K, N)
    m.weight, m.bias = _randn(N, K, seed=6, scale=K ** -0.5).cuda(), _randn(N, seed=7).cuda()
    x = _randn(1, M, K, seed=8)
    _check(m(x.cuda()), oracle.linear(_d(x), _d(m.weight.cpu()), _d(m.bias.cpu())),
           oracle.linear(x, m.weight.cpu(), m.bias.cpu()))


@pytest.mark.parametrize("c,hw", [(320, 64), (1280, 8), (960, 32)])
def test_group_norm(oracle, c, hw):
    from tinyfusers_b200.ff.group_norm import GroupNorm, group_norm
    x = _randn(2, c, hw, hw, seed=9) * 3 + 0.5
    _check(group_norm(x.cuda(), 32, 1e-5), oracle.group_norm(_d(x), 32, 1e-5), oracle.group_norm(x, 32, 1e-5))
    m = GroupNorm(32, c)
    m.weight, m.bias = (1 + 0.1 * _randn(c, seed=10)).cuda(), _randn(c, seed=11).cuda()
    _check(m(x.cuda()), oracle.group_norm_affine(_d(x), 32, _d(m.weight.cpu()), _d(m.bias.cpu()), 1e-5))


@pytest.mark.parametrize("B,T,C", [(1, 4096, 320), (2, 256, 1280), (1, 77, 768)])
def test_layer_norm(oracle, B, T, C):
    from tinyfusers_b200.ff.layer_norm import LayerNorm
    m = LayerNorm(C)
    m.weight, m.bias = (1 + 0.1 * _randn(C, seed=12)).cuda(), _randn(C, seed=13).cuda()
    x = _randn(B, T, C, seed=14) * 2 + 1
    _check(m(x.cuda()), oracle.layer_norm(_d(x), _d(m.weight.cpu()), _d(m.bias.cpu()), 1e-5),
           oracle.layer_norm(x, m.weight.cpu(), m.bias.cpu(), 1e-5))


@pytest.mark.parametrize("B,NH,Tq,Tk,d", [(1, 8, 4096, 4096, 40), (2, 8, 1024, 77, 80), (2, 8, 64, 64, 160), (1, 3, 50, 37, 24)])
def test_sdpa(oracle, B, NH, Tq, Tk, d):
    from tinyfusers_b200.attention.sdpa import scaled_dot_product_attention
    q, k, v = _randn(B, NH, Tq, d, seed=15), _randn(B, NH, Tk, d, seed=16), _randn(B, NH, Tk, d, seed=17)
    out = scaled_dot_product_attention(q.cuda(), k.cuda(), v.cuda())
    with torch.no_grad():
        ref64 = oracle.scaled_dot_product_attention(_d(q), _d(k), _d(v))
    _check(out, ref64)


@pytest.mark.parametrize("op", ["sigmoid", "silu", "swish", "gelu", "quick_gelu"])
def test_activations(oracle, op):
    from tinyfusers_b200.storage.tensor import Tensor
    x = _randn(3, 1000, seed=18) * 4
    ref = {"swish": oracle.silu}.get(op) or getattr(oracle, op)
    _check(getattr(Tensor, op)(x.cuda()), ref(_d(x)), ref(x))


def test_geglu_feed_forward(oracle):
    from tinyfusers_b200.ff.nn import FeedForward
    sd = {}
    oracle._add_linear(sd, "ff.net.0.proj", 320, 2560, seed=19)
    oracle._add_linear(sd, "ff.net.2", 1280, 320, seed=19)
    m = FeedForward(320)
    _load(m, sd, "ff")
    x = _randn(1, 1024, 320, seed=20)
    _check(m.net[0](x.cuda()), oracle.geglu(_sd64(sd), "ff.net.0", _d(x)), oracle.geglu(sd, "ff.net.0", x))
    _check(m(x.cuda()), oracle.feed_forward(_sd64(sd), "ff", _d(x)), oracle.feed_forward(sd, "ff", x))


# ---- blocks (§8a rows 8, 9, 13-15) ----------------------------------------------------------------------------------

@pytest.mark.parametrize("quirks", [True, False])
@pytest.mark.parametrize("cross", [False, True])
def test_cross_attention(oracle, quirks, cross):
    import tinyfusers_b200
    from tinyfusers_b200.attention.attention import SpatialTransformer
    sd = {}
    oracle.add_spatial_transformer(sd, "st", 640, 768, seed=21)
    st = SpatialTransformer(640, 768, 8, 80)
    _load(st, sd, "st")
    name = "attn2" if cross else "attn1"
    m = getattr(st.transformer_blocks[0], name)
    x = _randn(2, 256, 640, seed=22)
    ctx = _randn(2, 77, 768, seed=23) if cross else None
    tinyfusers_b200.set_quirks(quirks)
    try:
        out = m(x.cuda(), ctx.cuda() if cross else None)
    finally:
        tinyfusers_b200.set_quirks(True)
    with torch.no_grad():
        ref64 = oracle.cross_attention(_sd64(sd), f"st.transformer_blocks.0.{name}", _d(x), _d(ctx), 8, 80, quirks)
    _check(out, ref64)


def test_basic_transformer_block_and_spatial_transformer(oracle):
    from tinyfusers_b200.attention.attention import SpatialTransformer
    sd = {}
    oracle.add_spatial_transformer(sd, "st", 320, 768, seed=24)
    st = SpatialTransformer(320, 768, 8, 40)
    _load(st, sd, "st")
    ctx = _randn(2, 77, 768, seed=25)
    t = _randn(2, 1024, 320, seed=26)
    with torch.no_grad():
        ref64 = oracle.basic_transformer_block(_sd64(sd), "st.transformer_blocks.0", _d(t), _d(ctx), 8, 40)
    _check(st.transformer_blocks[0](t.cuda(), ctx.cuda()), ref64, tol=2e-5)        # 13 operators deep
    x = _randn(2, 320, 32, 32, seed=27)
    with torch.no_grad():
        ref64 = oracle.spatial_transformer(_sd64(sd), "st", _d(x), _d(ctx), 8, 40)
    _check(st(x.cuda(), ctx.cuda()), ref64, tol=2e-5)                             # 16 operators deep


@pytest.mark.parametrize("cin,cout", [(320, 320), (960, 640)])
def test_res_block(oracle, cin, cout):
    from tinyfusers_b200.vision.resnet import ResBlock
    sd = {}
    oracle.add_res_block(sd, "rb", cin, cout, seed=28)
    rb = ResBlock(cin, 1280, cout)
    _load(rb, sd, "rb")
    x, emb = _randn(2, cin, 32, 32, seed=29), _randn(1, 1280, seed=30)
    with torch.no_grad():
        ref64 = oracle.res_block(_sd64(sd), "rb", _d(x), _d(emb))
        ref32 = oracle.res_block(sd, "rb", x, emb)
    _check(rb(x.cuda(), emb.cuda()), ref64, ref32, tol=2e-5)                       # 8 operators deep


def test_up_down_sample(oracle):
    from tinyfusers_b200.vision.unet import Downsample, Upsample
    sd = {}
    oracle._add_conv(sd, "up.conv", 640, 640, 3, seed=31)
    oracle._add_conv(sd, "down.op", 640, 640, 3, seed=31)
    up, down = Upsample(640), Downsample(640)
    _load(up, sd, "up")
    _load(down, sd, "down")
    x = _randn(2, 640, 16, 16, seed=32)
    _check(up(x.cuda()), oracle.upsample(_sd64(sd), "up", _d(x)), oracle.upsample(sd, "up", x))
    _check(down(x.cuda()), oracle.downsample(_sd64(sd), "down", _d(x)), oracle.downsample(sd, "down", x))


# ---- BASELINE.json configs[0] at its full size -----------------------------------------------------------------------

def test_c1_down_block0_fp32_full_size(oracle):
    """ResBlock(320,1280,320) + SpatialTransformer(320,768,8,40), 64x64 latent, batch 1, fp32: 4096-token self-attention,
    77-token cross-attention; 24 operators chained."""
    from tinyfusers_b200.attention.attention import SpatialTransformer
    from tinyfusers_b200.vision.resnet import ResBlock
    sd = {}
    oracle.add_res_block(sd, "rb", 320, 320, seed=31)
    oracle.add_spatial_transformer(sd, "st", 320, 768, seed=31)
    g = torch.Generator().manual_seed(31)
    x = torch.randn(1, 320, 64, 64, generator=g)
    emb = torch.randn(1, 1280, generator=g)
    ctx = torch.randn(1, 77, 768, generator=g)
    rb, st = ResBlock(320, 1280, 320), SpatialTransformer(320, 768, 8, 40)
    _load(rb, sd, "rb")
    _load(st, sd, "st")
    y = st(rb(x.cuda(), emb.cuda()), ctx.cuda())
    with torch.no_grad():
        ref32 = oracle.spatial_transformer(sd, "st", oracle.res_block(sd, "rb", x, emb), ctx, 8, 40, True)
        sd64 = _sd64(sd)
        ref64 = oracle.spatial_transformer(sd64, "st", oracle.res_block(sd64, "rb", _d(x), _d(emb)), _d(ctx), 8, 40, True)
    _check(y, ref64, ref32, tol=2e-5)


# ---- the whole step in fp32 (configs[1]'s function set, small latent so the CPU oracle takes seconds) ---------------

def test_unet_and_sampler_step_fp32(oracle, unet_sd, sd_model):
    lat, unc, ctx = oracle.make_inputs(1, 16)
    x2, c2 = torch.cat([lat, lat]), torch.cat([unc, ctx])
    sd64 = _sd64(unet_sd)
    with torch.no_grad():
        ref64 = oracle.unet_forward(sd64, _d(x2), [981], _d(c2), quirks=True)
    out = sd_model.model.diffusion_model(x2.cuda(), torch.tensor([981]).cuda(), c2.cuda())
    _check(out, ref64, tol=1e-4)        # ~700 fp32 operators deep
    ts, alphas, alphas_prev = oracle.sampler_schedule(50)
    i = 30
    with torch.no_grad():
        ref = oracle.sampler_step(sd64, _d(unc), _d(ctx), _d(lat), [ts[i]], _d(alphas[[i]]), _d(alphas_prev[[i]]), 7.5)
    got = sd_model(unc.cuda(), ctx.cuda(), lat.cuda(), torch.tensor([ts[i]]).cuda(), alphas[[i]].cuda(),
                   alphas_prev[[i]].cuda(), torch.tensor([7.5]))
    _check(got, ref, tol=1e-4)


# ---- the production fp16 path against the fp32 path ON THE GPU at BASELINE.json's full sizes ------------------------
# (the CPU oracle needs minutes for a 64x64 / 96x96 UNet; the fp32 mode above is oracle-checked operator by operator and
# at C1's full size, so it carries the oracle to the sizes the bench runs at)

@pytest.mark.parametrize("hw,t", [(64, 981), (96, 301)])
def test_fp16_step_matches_fp32_mode_full_size(oracle, sd_model, hw, t):
    import tinyfusers_b200
    lat, unc, ctx = oracle.make_inputs(1, hw, seed=11, ctx_seed=12)
    ts = torch.tensor([t]).cuda()
    g = torch.tensor([7.5])
    e32 = sd_model.get_model_output(unc.cuda(), ctx.cuda(), lat.cuda(), ts, g)
    assert e32.shape == (1, 4, hw, hw) and torch.isfinite(e32).all()
    tinyfusers_b200.set_precision("fp16")
    e16 = sd_model.get_model_output(unc.cuda(), ctx.cuda(), lat.cuda(), ts, g)
    assert rel_err(e16, e32) < 3e-2     # same bound as test_unet_gpu.py's CFG output against the oracle at 32x32


def test_sample_loop_fp32(oracle, unet_sd, sd_model):
    """StableDiffusion.sample in fp32 mode = the reference's host loop over __call__ (example/sd1.py:68-73)."""
    lat, unc, ctx = oracle.make_inputs(1, 16, seed=21, ctx_seed=22)
    ts, alphas, alphas_prev = oracle.sampler_schedule(3)
    out = sd_model.sample(unc.cuda(), ctx.cuda(), lat.cuda(), ts, alphas, alphas_prev, 7.5)
    sd64 = _sd64(unet_sd)
    x = _d(lat)
    with torch.no_grad():
        for i in reversed(range(len(ts))):
            x = oracle.sampler_step(sd64, _d(unc), _d(ctx), x, [ts[i]], _d(alphas[[i]]), _d(alphas_prev[[i]]), 7.5)
    _check(out, x, tol=2e-4)      # three CFG steps, each ~700 fp32 operators, guidance 7.5
