"""Block-level parity (the reference has no such tests — SURVEY.md §4): ResBlock, CrossAttention with the
head-major reshape quirk, BasicTransformerBlock with the LayerNorm stride quirk, SpatialTransformer,
Up/Downsample, all against the oracle on seeded weights/inputs. fp16 tolerance 1e-2 (max-normalised)."""
import contextlib
import io

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-2


def _load(obj, sd, prefix):
    from tinyfusers_b200.storage.state import update_state
    with contextlib.redirect_stdout(io.StringIO()):
        update_state(obj, sd, prefix)


@pytest.mark.parametrize("cin,cout,hw", [(320, 320, 16), (320, 640, 16), (2560, 1280, 8), (960, 640, 12)])
def test_resblock(oracle, cin, cout, hw):
    from tinyfusers_b200.vision.resnet import ResBlock
    sd = {}
    oracle.add_res_block(sd, "rb", cin, cout, seed=11)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, cin, hw, hw, generator=g)
    emb = torch.randn(1, 1280, generator=g)
    rb = ResBlock(cin, 1280, cout)
    _load(rb, sd, "rb")
    out = rb(x.cuda(), emb.cuda())
    assert rel_err(out, oracle.res_block(sd, "rb", x, emb)) < TOL


@pytest.mark.parametrize("cin,cout,hw", [(1280, 1280, 8), (2560, 1280, 8), (1280, 1280, 16)])
def test_resblock_cluster_splitk(oracle, cin, cout, hw):
    """The optional split-K fold inside a thread-block cluster (tf_gemm_set_cluster_splitk: partial tiles reduced through
    distributed shared memory in the conv launch itself, GroupNorm statistics included; off by default because it measured
    slower) on the small-M ResBlocks whose convolutions split K: against the oracle, and against the default workspace + fold
    path (same fp32 products, different summation order)."""
    from tinyfusers_b200.native.b200.ops import b200
    from tinyfusers_b200.vision.resnet import ResBlock
    sd = {}
    oracle.add_res_block(sd, "rb", cin, cout, seed=21)
    g = torch.Generator().manual_seed(21)
    x = torch.randn(2, cin, hw, hw, generator=g)
    emb = torch.randn(1, 1280, generator=g)
    rb = ResBlock(cin, 1280, cout)
    _load(rb, sd, "rb")
    ref = oracle.res_block(sd, "rb", x, emb)
    plain = rb(x.cuda(), emb.cuda())
    try:
        b200.check(b200.tf_gemm_set_cluster_splitk(1), "cluster split-K on")
        clus = rb(x.cuda(), emb.cuda())
    finally:
        b200.tf_gemm_set_cluster_splitk(-1)
    assert rel_err(clus, ref) < TOL
    assert rel_err(clus, plain) < 3e-3


@pytest.mark.parametrize("quirks", [True, False])
@pytest.mark.parametrize("c,d,T,B", [(320, 40, 256, 2), (640, 80, 64, 2), (1280, 160, 64, 1)])
def test_cross_attention_self_and_cross(oracle, quirks, c, d, T, B):
    import tinyfusers_b200
    from tinyfusers_b200.attention.attention import CrossAttention
    sd = {}
    oracle.add_spatial_transformer(sd, "st", c, 768, seed=12)
    p = "st.transformer_blocks.0"
    g = torch.Generator().manual_seed(12)
    x = torch.randn(B, T, c, generator=g)
    ctx = torch.randn(B, 77, 768, generator=g)
    tinyfusers_b200.set_quirks(quirks)
    try:
        a1 = CrossAttention(c, c, 8, d)
        _load(a1, sd, p + ".attn1")
        assert rel_err(a1(x.cuda()), oracle.cross_attention(sd, p + ".attn1", x, None, 8, d, quirks)) < TOL
        a2 = CrossAttention(c, 768, 8, d)
        _load(a2, sd, p + ".attn2")
        assert rel_err(a2(x.cuda(), ctx.cuda()), oracle.cross_attention(sd, p + ".attn2", x, ctx, 8, d, quirks)) < TOL
    finally:
        tinyfusers_b200.set_quirks(True)


@pytest.mark.parametrize("quirks,strided", [(True, False), (False, False), (True, True)])
def test_basic_transformer_block(oracle, quirks, strided):
    import tinyfusers_b200
    from tinyfusers_b200.attention.attention import BasicTransformerBlock
    c, d = 320, 40
    sd = {}
    oracle.add_spatial_transformer(sd, "st", c, 768, seed=13)
    p = "st.transformer_blocks.0"
    g = torch.Generator().manual_seed(13)
    x = torch.randn(2, 256, c, generator=g)
    ctx = torch.randn(2, 77, 768, generator=g)
    tinyfusers_b200.set_quirks(quirks)
    tinyfusers_b200.set_layernorm_strided(strided)
    try:
        blk = BasicTransformerBlock(c, 768, 8, d)
        _load(blk, sd, p)
        out = blk(x.cuda(), ctx.cuda())
    finally:
        tinyfusers_b200.set_quirks(True)
        tinyfusers_b200.set_layernorm_strided(False)
    assert rel_err(out, oracle.basic_transformer_block(sd, p, x, ctx, 8, d, quirks, strided)) < TOL


@pytest.mark.parametrize("c,d,hw", [(320, 40, 16), (640, 80, 8), (1280, 160, 8)])
def test_spatial_transformer(oracle, c, d, hw):
    from tinyfusers_b200.attention.attention import SpatialTransformer
    sd = {}
    oracle.add_spatial_transformer(sd, "st", c, 768, seed=14)
    g = torch.Generator().manual_seed(14)
    x = torch.randn(2, c, hw, hw, generator=g)
    ctx = torch.randn(2, 77, 768, generator=g)
    st = SpatialTransformer(c, 768, 8, d)
    _load(st, sd, "st")
    assert rel_err(st(x.cuda(), ctx.cuda()), oracle.spatial_transformer(sd, "st", x, ctx, 8, d, True)) < TOL


def test_up_down_sample(oracle):
    from tinyfusers_b200.vision.unet import Downsample, Upsample
    sd = {}
    oracle._add_conv(sd, "u.conv", 640, 640, 3, 15)
    oracle._add_conv(sd, "d.op", 320, 320, 3, 15)
    g = torch.Generator().manual_seed(15)
    xu, xd = torch.randn(2, 640, 8, 8, generator=g), torch.randn(2, 320, 16, 16, generator=g)
    up, down = Upsample(640), Downsample(320)
    _load(up, sd, "u")
    _load(down, sd, "d")
    assert rel_err(up(xu.cuda()), oracle.upsample(sd, "u", xu)) < TOL
    assert rel_err(down(xd.cuda()), oracle.downsample(sd, "d", xd)) < TOL


@pytest.mark.parametrize("n,c,h,w", [(2, 1280, 8, 8), (2, 1280, 16, 16), (2, 640, 32, 32), (1, 64, 4, 8), (3, 128, 8, 16), (1, 128, 12, 20)])
def test_upsample_folded_into_conv(oracle, n, c, h, w):
    """tf_conv2d_up2x_nhwc_f16 (four 2 x 2 phase convolutions of the original image with pre-summed taps) against the oracle's
    upsample + 3x3 convolution (reference: vision/unet.py:78-84) and against the two-launch path (upsample2x kernel + conv) -
    the two differ only by the fp16 rounding of the summed taps. The last shape (240 pixels) has no exact tiling: it must
    take the two-launch path and still match."""
    from tinyfusers_b200.vision import unet as U
    sd = {}
    oracle._add_conv(sd, "u.conv", c, c, 3, 16)
    g = torch.Generator().manual_seed(16 + h)
    x = torch.randn(n, c, h, w, generator=g)
    up = U.Upsample(c)
    _load(up, sd, "u")
    ref = oracle.upsample(sd, "u", x)
    fused = up(x.cuda())
    assert fused.shape == ref.shape == (n, c, 2 * h, 2 * w)
    assert rel_err(fused, ref) < TOL
    old = U.FUSED_UPSAMPLE
    try:
        U.FUSED_UPSAMPLE = False
        plain = up(x.cuda())
    finally:
        U.FUSED_UPSAMPLE = old
    assert rel_err(fused, plain) < 3e-3


def test_blocks_match_reference_goldens():
    """CUDA path against the outputs of the reference's OWN Python (tests/golden/reference_outputs.npz)."""
    import os
    import numpy as np
    from tinyfusers_b200.attention.attention import BasicTransformerBlock, SpatialTransformer
    from tinyfusers_b200.vision.resnet import ResBlock
    from oracle import ref_ops as R
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_outputs.npz"))

    def rnd(seed, *shape):
        g = np.random.Generator(np.random.Philox(seed))
        return torch.from_numpy(g.standard_normal(shape, dtype=np.float32))

    sd = {}
    R.add_spatial_transformer(sd, "st", 320, 768, seed=501)
    xt, ctx = rnd(502, 2, 64, 320), rnd(503, 2, 77, 768)
    blk = BasicTransformerBlock(320, 768, 8, 40)
    _load(blk, sd, "st.transformer_blocks.0")
    assert rel_err(blk(xt.cuda(), ctx.cuda()), torch.from_numpy(gold["transformer_block_b2"])) < TOL
    st = SpatialTransformer(320, 768, 8, 40)
    _load(st, sd, "st")
    assert rel_err(st(rnd(504, 2, 320, 8, 8).cuda(), ctx.cuda()), torch.from_numpy(gold["spatial_transformer"])) < TOL
    sd = {}
    R.add_res_block(sd, "rb", 320, 640, seed=601)
    rb = ResBlock(320, 1280, 640)
    _load(rb, sd, "rb")
    assert rel_err(rb(rnd(602, 2, 320, 8, 8).cuda(), rnd(603, 1, 1280).cuda()), torch.from_numpy(gold["res_block"])) < TOL


def test_spatial_transformer_with_layernorm_folded_into_gemms(oracle):
    """Opt-in path (TINYFUSERS_B200_FUSE_LN=1 | 2): norm1 / norm2 (/ norm3) folded into the projections that consume them
    (tf_gemm_ex_f16: row statistics from the producing GEMM, gamma in the weights, mean / rstd applied in the epilogue)."""
    from tinyfusers_b200.attention.attention import SpatialTransformer
    from tinyfusers_b200.runtime import standalone_context
    from tinyfusers_b200.storage.state import update_state
    import contextlib, io
    sd = {}
    oracle.add_spatial_transformer(sd, "st", 640, 768, seed=77)
    g = torch.Generator().manual_seed(77)
    x = torch.randn(2, 640, 16, 16, generator=g)
    c = torch.randn(2, 77, 768, generator=g)
    st = SpatialTransformer(640, 768, 8, 80)
    with contextlib.redirect_stdout(io.StringIO()):
        update_state(st, sd, "st")
    with torch.no_grad():
        ref = oracle.spatial_transformer(sd, "st", x, c, 8, 80, True)
    ctx = standalone_context()
    old = (ctx.fuse_ln, ctx.fuse_ln_all)
    try:
        ctx.fuse_ln, ctx.fuse_ln_all = False, False
        plain = st(x.cuda(), c.cuda())
        ctx.fuse_ln, ctx.fuse_ln_all = True, True      # all three norms folded (norm3 into the GEGLU projection too)
        folded = st(x.cuda(), c.cuda())
    finally:
        ctx.fuse_ln, ctx.fuse_ln_all = old
    assert rel_err(plain, ref) < 1e-2 and rel_err(folded, ref) < 1e-2
    assert rel_err(folded, plain) < 5e-3 and not torch.equal(folded, plain)     # a different path really ran
