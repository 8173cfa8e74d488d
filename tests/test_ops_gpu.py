"""Per-operator parity: the CUDA path (called through the drop-in wrappers -> ctypes C-ABI) against the
oracle restatement of the reference, on the same seeded inputs. Tolerance: max|a-b|/max|b| <= 1e-2
(fp16 operands, fp32 accumulate — BASELINE.json north_star; the reference's own tests use atol=rtol=1e-2,
tests/conv2d.py:33, tests/sdpa.py:79-81)."""
import math

import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-2


def _g(seed):
    return torch.Generator().manual_seed(seed)


@pytest.mark.parametrize("shape,cout,k,stride", [
    ((2, 320, 16, 16), 320, 3, 1), ((1, 64, 12, 20), 32, 3, 1), ((2, 128, 16, 16), 128, 3, 2),
    ((1, 2, 10, 10), 1, 3, 1),      # tiny channel counts (reference test shape family, tests/conv2d.py:14-18)
    ((2, 320, 8, 8), 640, 1, 1), ((3, 24, 5, 7), 40, 1, 1), ((1, 640, 9, 9), 640, 3, 2),
])
def test_conv2d(oracle, shape, cout, k, stride):
    from tinyfusers_b200.vision.conv2d import Conv2d, conv_2d
    g = _g(1)
    x = torch.randn(shape, generator=g)
    cin = shape[1]
    w = torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)
    b = torch.randn(cout, generator=g) * 0.1
    pad = [k // 2, k // 2]
    ref = oracle.conv2d(x, w, b, stride=(stride, stride), padding=pad)
    conv = Conv2d(cin, cout, kernel_size=[k, k], stride=[stride, stride], padding=pad)
    conv.weight, conv.bias = w.cuda(), b.cuda()
    out = conv(x.cuda())
    assert out.shape == ref.shape and out.dtype == torch.float32
    assert rel_err(out, ref) < TOL
    out2 = conv_2d(x.cuda(), w.cuda(), pad, [stride, stride], [1, 1])
    assert rel_err(out2, oracle.conv2d(x, w, None, stride=(stride, stride), padding=pad)) < TOL


def test_conv2d_unsupported_geometry_fails_loudly():
    from tinyfusers_b200.vision.conv2d import Conv2d
    conv = Conv2d(8, 8, kernel_size=[5, 5], padding=[2, 2])
    with pytest.raises(RuntimeError):
        conv(torch.zeros(1, 8, 8, 8, device="cuda"))


def test_conv_in_cin4(oracle):
    from tinyfusers_b200.vision.conv2d import Conv2d
    g = _g(2)
    x = torch.randn(2, 4, 32, 32, generator=g)
    w = torch.randn(320, 4, 3, 3, generator=g) / 6
    b = torch.randn(320, generator=g) * 0.1
    conv = Conv2d(4, 320, kernel_size=[3, 3], padding=[1, 1])
    conv.weight, conv.bias = w.cuda(), b.cuda()
    assert rel_err(conv(x.cuda()), oracle.conv2d(x, w, b, padding=(1, 1))) < TOL


@pytest.mark.parametrize("m,k,n,bias", [(1, 320, 1280, True), (77, 768, 320, False), (512, 1280, 1280, True),
                                         (4096, 320, 320, True), (5, 12, 20, True)])
def test_linear(oracle, m, k, n, bias):
    from tinyfusers_b200.ff.linear import Linear
    g = _g(3)
    x = torch.randn(2, m, k, generator=g)
    w = torch.randn(n, k, generator=g) / math.sqrt(k)
    b = torch.randn(n, generator=g) * 0.1 if bias else None
    lin = Linear(k, n, bias=bias)
    lin.weight, lin.bias = w.cuda(), (b.cuda() if bias else None)
    out = lin(x.cuda())
    assert out.shape == (2, m, n)
    assert rel_err(out, oracle.linear(x, w, b)) < TOL


@pytest.mark.parametrize("shape", [(2, 320, 16, 16), (1, 640, 8, 8), (2, 2560, 8, 8), (2, 960, 4, 4), (2048, 64, 2, 2)])
def test_group_norm(oracle, shape):
    from tinyfusers_b200.ff.group_norm import GroupNorm, group_norm
    g = _g(4)
    x = torch.randn(shape, generator=g) * 1.7 + 0.6
    C = shape[1]
    gamma, beta = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    gn = GroupNorm(32, C)
    gn.weight, gn.bias = gamma.cuda(), beta.cuda()
    assert rel_err(gn(x.cuda()), oracle.group_norm_affine(x, 32, gamma, beta, 1e-5)) < TOL
    assert rel_err(group_norm(x.cuda(), 32, 1e-5), oracle.group_norm(x, 32, 1e-5)) < TOL


@pytest.mark.parametrize("shape", [(1, 256, 320), (2, 256, 320), (2, 64, 1280), (4, 32, 640), (3, 16, 320)])
@pytest.mark.parametrize("strided", [False, True])
def test_layer_norm(oracle, shape, strided):
    import tinyfusers_b200
    from tinyfusers_b200.ff.layer_norm import LayerNorm
    g = _g(5)
    x = torch.randn(shape, generator=g) * 2 + 0.3
    C = shape[-1]
    gamma, beta = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    ln = LayerNorm(C)
    ln.weight, ln.bias = gamma.cuda(), beta.cuda()
    tinyfusers_b200.set_layernorm_strided(strided)
    try:
        out = ln(x.cuda())
    finally:
        tinyfusers_b200.set_layernorm_strided(False)
    assert rel_err(out, oracle.layer_norm(x, gamma, beta, 1e-5, ln_strided=strided)) < TOL


@pytest.mark.parametrize("B,NH,Tq,Tk,HS", [(2, 8, 256, 256, 40), (2, 8, 256, 77, 40), (1, 8, 64, 64, 160),
                                            (2, 8, 100, 77, 80), (1, 12, 130, 200, 64), (2, 2, 1024, 1024, 40)])
def test_sdpa(oracle, B, NH, Tq, Tk, HS):
    from tinyfusers_b200.attention.sdpa import scaled_dot_product_attention
    g = _g(6)
    q, k, v = (torch.randn(B, NH, T, HS, generator=g) for T in (Tq, Tk, Tk))
    out = scaled_dot_product_attention(q.cuda(), k.cuda(), v.cuda())
    assert out.shape == (B, NH, Tq, HS)
    assert rel_err(out, oracle.scaled_dot_product_attention(q, k, v)) < TOL


def test_sdpa_peaked_rows(oracle):
    """large score range: online softmax rescaling across key blocks must stay exact"""
    from tinyfusers_b200.attention.sdpa import scaled_dot_product_attention
    g = _g(7)
    q = torch.randn(1, 2, 128, 40, generator=g) * 4
    k = torch.randn(1, 2, 512, 40, generator=g) * 4
    v = torch.randn(1, 2, 512, 40, generator=g)
    out = scaled_dot_product_attention(q.cuda(), k.cuda(), v.cuda())
    assert rel_err(out, oracle.scaled_dot_product_attention(q.half().float(), k.half().float(), v.half().float())) < TOL


def test_geglu_feedforward(oracle):
    from tinyfusers_b200.ff.nn import FeedForward
    g = _g(8)
    dim = 320
    x = torch.randn(2, 128, dim, generator=g)
    sd = {}
    oracle._add_linear(sd, "ff.net.0.proj", dim, 8 * dim, 7)
    oracle._add_linear(sd, "ff.net.2", 4 * dim, dim, 7)
    ff = FeedForward(dim)
    ff.net[0].proj.weight, ff.net[0].proj.bias = sd["ff.net.0.proj.weight"].cuda(), sd["ff.net.0.proj.bias"].cuda()
    ff.net[2].weight, ff.net[2].bias = sd["ff.net.2.weight"].cuda(), sd["ff.net.2.bias"].cuda()
    assert rel_err(ff.net[0](x.cuda()), oracle.geglu(sd, "ff.net.0", x)) < TOL
    assert rel_err(ff(x.cuda()), oracle.feed_forward(sd, "ff", x)) < TOL


@pytest.mark.parametrize("t", [1, 21, 501, 981])
def test_timestep_embedding(oracle, t):
    from tinyfusers_b200.vision.unet import timestep_embedding
    out = timestep_embedding(torch.tensor([t]), 320)
    ref = oracle.timestep_embedding([t], 320)
    assert out.shape == (1, 320)
    assert (out.cpu() - ref).abs().max().item() < 2e-6   # fp64 angles on both sides


def test_activations(oracle):
    from tinyfusers_b200.storage.tensor import Tensor
    x = torch.linspace(-12, 12, 4001)
    for name in ("sigmoid", "silu", "swish", "gelu", "quick_gelu"):
        ref = getattr(oracle, "silu" if name == "swish" else name)(x)
        out = getattr(Tensor, name)(x.cuda())
        assert (out.cpu() - ref).abs().max().item() < 2e-5, name
    assert Tensor.sequential([lambda v: v + 1, lambda v: v * 2], 3) == 8


def test_ddim_and_alphas(oracle):
    from tinyfusers_b200.variants.sd import StableDiffusion, get_alphas_cumprod
    ac = get_alphas_cumprod()
    ref = oracle.get_alphas_cumprod()
    assert rel_err(ac, ref) < 1e-6
    g = _g(9)
    x, e = torch.randn(1, 4, 64, 64, generator=g), torch.randn(1, 4, 64, 64, generator=g)
    a_t, a_prev = ref[[501]], ref[[481]]
    sd = StableDiffusion.__new__(StableDiffusion)
    xp, p0 = StableDiffusion.get_x_prev_and_pred_x0(sd, x.cuda(), e.cuda(), a_t.cuda(), a_prev.cuda())
    rxp, rp0 = oracle.get_x_prev_and_pred_x0(x, e, a_t, a_prev)
    assert rel_err(xp, rxp) < 1e-5 and rel_err(p0, rp0) < 1e-5


@pytest.mark.parametrize("kind,n,h,w,cin,cout,splits", [
    ("conv3", 2, 32, 32, 320, 320, 0), ("conv3", 2, 16, 16, 640, 1280, 0), ("conv3", 2, 8, 8, 1280, 1280, 4),
    ("conv1", 2, 32, 32, 320, 640, 0), ("conv3s2", 2, 32, 32, 320, 320, 0), ("conv1", 2, 8, 8, 2560, 1280, 3),
])
def test_producer_statistics_and_fused_groupnorm(oracle, kind, n, h, w, cin, cout, splits):
    """tf_conv2d_nhwc_gn_f16 / tf_gemm_gn_f16 leave per-slot {sum, sumsq}; tf_groupnorm_fused_nhwc_f16 consumes them.
    Checked against (a) the sums of the produced fp16 tensor and (b) the oracle's GroupNorm + SiLU of the oracle conv."""
    import torch.nn.functional as F
    from tinyfusers_b200 import packing
    from tinyfusers_b200.native.b200.ops import b200
    from tinyfusers_b200.runtime import gn_unit, standalone_context, stream_ptr
    ctx = standalone_context()
    g = _g(11)
    k = 1 if kind == "conv1" else 3
    stride = 2 if kind == "conv3s2" else 1
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)
    bias = torch.randn(cout, generator=g) * 0.1
    gamma, beta = 1 + 0.1 * torch.randn(cout, generator=g), 0.1 * torch.randn(cout, generator=g)
    ho, wo = (h + 2 * (k // 2) - k) // stride + 1, (w + 2 * (k // 2) - k) // stride + 1
    xa = x.permute(0, 2, 3, 1).contiguous().half().cuda()
    wp = (packing.conv3x3_weight(wt.cuda(), 64, 8) if k == 3 else packing.conv1x1_weight(wt.cuda(), 8))
    out = torch.empty(n, ho, wo, cout, dtype=torch.half, device="cuda")
    unit = gn_unit(cout)
    assert b200.tf_gn_stats_supported(n, ho, wo, cout, unit, 1 if k == 3 else 0)
    slots = ho * wo // 32
    stats = torch.full((n, slots, cout // unit, 2), float("nan"), device="cuda")
    bc = bias.cuda()
    b200.tf_gemm_set_tuning(0, splits)
    try:
        st = b200.tf_conv2d_nhwc_gn_f16(xa.data_ptr(), n, h, w, cin, cin, wp.data_ptr(), cout, k, stride, out.data_ptr(), cout,
                                        bc.data_ptr(), None, 0, 0, ctx.ws.data_ptr(), ctx.ws_bytes, stats.data_ptr(), unit,
                                        stream_ptr())
    finally:
        b200.tf_gemm_set_tuning(0, 0)
    b200.check(st, "tf_conv2d_nhwc_gn_f16")
    torch.cuda.synchronize()
    o32 = out.float()
    got = stats.sum(dim=1)                                                   # (n, units, 2)
    want_s = o32.reshape(n, ho * wo, cout // unit, unit).sum(dim=(1, 3))
    want_q = (o32 * o32).reshape(n, ho * wo, cout // unit, unit).sum(dim=(1, 3))
    assert torch.isfinite(stats).all()
    assert rel_err(got[..., 0], want_s) < 1e-4 and rel_err(got[..., 1], want_q) < 1e-4
    y = torch.empty_like(out)
    gc, bt = gamma.cuda(), beta.cuda()
    st = b200.tf_groupnorm_fused_nhwc_f16(out.data_ptr(), cout, cout, stats.data_ptr(), unit, None, 0, 0, None, 1,
                                          y.data_ptr(), cout, n, ho * wo, 32, gc.data_ptr(), bt.data_ptr(), 1e-5, 1, stream_ptr())
    b200.check(st, "tf_groupnorm_fused_nhwc_f16")
    ref = F.silu(F.group_norm(oracle.conv2d(x, wt, bias, stride=(stride, stride), padding=[k // 2, k // 2]), 32, gamma, beta, 1e-5))
    assert rel_err(y.permute(0, 3, 1, 2), ref) < TOL


def test_fused_groupnorm_two_sources(oracle):
    """Channel concatenation of two producers (the UNet's skip connections): 640 + 320 channels, groups of 30."""
    import torch.nn.functional as F
    from tinyfusers_b200.native.b200.ops import b200
    from tinyfusers_b200.runtime import standalone_context, stream_ptr
    standalone_context()
    g = _g(12)
    n, hw, c1, c2 = 2, 256, 640, 320
    cat = torch.randn(n, hw, c1 + c2, generator=g).half().cuda()
    gamma, beta = (1 + 0.1 * torch.randn(c1 + c2, generator=g)).cuda(), (0.1 * torch.randn(c1 + c2, generator=g)).cuda()
    def slot_stats(t, unit):
        f = t.float().reshape(n, hw // 32, 32, t.shape[-1] // unit, unit)
        return torch.stack((f.sum(dim=(2, 4)), (f * f).sum(dim=(2, 4))), dim=-1).contiguous()
    s1, s2 = slot_stats(cat[..., :c1], 10), slot_stats(cat[..., c1:], 10)
    y = torch.empty_like(cat)
    st = b200.tf_groupnorm_fused_nhwc_f16(cat.data_ptr(), c1 + c2, c1, s1.data_ptr(), 10, cat.data_ptr() + 2 * c1, c1 + c2, c2,
                                          s2.data_ptr(), 10, y.data_ptr(), c1 + c2, n, hw, 32, gamma.data_ptr(), beta.data_ptr(),
                                          1e-5, 0, stream_ptr())
    b200.check(st, "tf_groupnorm_fused_nhwc_f16")
    ref = F.group_norm(cat.float().cpu().permute(0, 2, 1).reshape(n, c1 + c2, 16, 16), 32, gamma.cpu(), beta.cpu(), 1e-5)
    assert rel_err(y.permute(0, 2, 1).reshape(n, c1 + c2, 16, 16), ref) < TOL


@pytest.mark.parametrize("ctas", [1, 2])
@pytest.mark.parametrize("case", [
    ("gemm", 8192, 320, 320), ("gemm", 300, 72, 200), ("gemm", 2048, 640, 2560), ("gemm", 384, 1280, 640),
    ("gemm_geglu", 2048, 5120, 640), ("gemm_split", 512, 1280, 5120),
    ("conv", 2, 64, 64, 320, 320, 1), ("conv", 2, 32, 32, 640, 640, 1), ("conv", 1, 24, 40, 64, 96, 1),
    ("conv", 2, 32, 32, 320, 320, 2), ("conv_split", 2, 16, 16, 1280, 1280, 1), ("conv", 3, 8, 8, 128, 64, 1),
])
def test_gemm_single_and_pair_cta_tiles(oracle, case, ctas):
    """The same GEMM / implicit-GEMM conv through the single-CTA (128-row) and the CTA-pair (cta_group::2, 256-row)
    kernels, each against the fp32 oracle; odd m-tile counts, ragged N, split-K and the GEGLU epilogue included."""
    import torch.nn.functional as F
    from tinyfusers_b200 import packing
    from tinyfusers_b200.native.b200.ops import b200
    from tinyfusers_b200.runtime import standalone_context, stream_ptr
    ctx = standalone_context()
    g = _g(21)
    kind = case[0]
    b200.tf_gemm_set_ctas(ctas)
    b200.tf_gemm_set_tuning(0, 3 if kind.endswith("_split") else 0)
    try:
        if kind.startswith("gemm"):
            _, M, N, K = case
            A = torch.randn(M, K, generator=g)
            W = torch.randn(N, K, generator=g) / math.sqrt(K)
            bias = torch.randn(N, generator=g) * 0.1
            geglu = kind == "gemm_geglu"
            res = None if geglu else torch.randn(M, N, generator=g)
            ref = A.half().float() @ W.half().float().t() + bias
            if geglu:
                val, gate = ref[:, :N // 2], ref[:, N // 2:]
                ref = val * F.gelu(gate, approximate="tanh")
                Wp, bp = packing.geglu_pack(W.cuda(), bias.cuda())
            else:
                ref = ref + res.half().float()
                Wp, bp = W.half().cuda(), bias.cuda()
            No = N // 2 if geglu else N
            Ad = A.half().cuda()
            out = torch.empty(M, No, dtype=torch.half, device="cuda")
            rd = res.half().cuda() if res is not None else None
            st = b200.tf_gemm_f16(Ad.data_ptr(), K, Wp.data_ptr(), K, out.data_ptr(), No, M, N, K, bp.data_ptr(),
                                  rd.data_ptr() if rd is not None else None, No, b200.TF_EPI_GEGLU if geglu else 0,
                                  ctx.ws.data_ptr(), ctx.ws_bytes, stream_ptr())
            b200.check(st, "tf_gemm_f16")
            assert rel_err(out, ref) < 3e-3
        else:
            _, n, h, w, cin, cout, stride = case
            x = torch.randn(n, cin, h, w, generator=g)
            wt = torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)
            bias = torch.randn(cout, generator=g) * 0.1
            ref = oracle.conv2d(x.half().float(), wt.half().float(), bias, stride=(stride, stride), padding=[1, 1])
            xa = x.permute(0, 2, 3, 1).contiguous().half().cuda()
            wp = packing.conv3x3_weight(wt.cuda(), 64, 8)
            ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
            out = torch.empty(n, ho, wo, cout, dtype=torch.half, device="cuda")
            bc = bias.cuda()
            st = b200.tf_conv2d_nhwc_f16(xa.data_ptr(), n, h, w, cin, cin, wp.data_ptr(), cout, 3, stride, out.data_ptr(), cout,
                                         bc.data_ptr(), None, 0, 0, ctx.ws.data_ptr(), ctx.ws_bytes, stream_ptr())
            b200.check(st, "tf_conv2d_nhwc_f16")
            assert rel_err(out.permute(0, 3, 1, 2), ref) < 3e-3
    finally:
        b200.tf_gemm_set_ctas(0)
        b200.tf_gemm_set_tuning(0, 0)


def test_c_abi_graph_capture_without_torch_graphs():
    """tf_graph_begin_capture / end_capture / launch / destroy: a GEMM -> LayerNorm -> GEMM chain captured through the library's
    own C-ABI (what a CuPy / ctypes host of the reference would use; the package itself captures with torch.cuda.CUDAGraph)
    replays to the bit-identical result of the eager chain, and follows its inputs on the next launch."""
    import ctypes
    from tinyfusers_b200.native.b200.ops import b200
    b200.init(0)
    g = _g(31)
    M, K, N = 256, 320, 640
    A = torch.randn(M, K, generator=g).half().cuda()
    W1 = (torch.randn(N, K, generator=g) / math.sqrt(K)).half().cuda()
    W2 = (torch.randn(K, N, generator=g) / math.sqrt(N)).half().cuda()
    gamma, beta = torch.ones(N).cuda(), torch.zeros(N).cuda()
    h = torch.empty(M, N, dtype=torch.half, device="cuda")
    hn = torch.empty_like(h)
    out = torch.empty(M, K, dtype=torch.half, device="cuda")
    ws = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    side = torch.cuda.Stream()
    S = side.cuda_stream

    def chain():
        b200.check(b200.tf_gemm_f16(A.data_ptr(), K, W1.data_ptr(), K, h.data_ptr(), N, M, N, K, None, None, 0, 0, ws.data_ptr(), ws.numel(), S), "gemm1")
        b200.check(b200.tf_layernorm_f16(h.data_ptr(), hn.data_ptr(), M, N, gamma.data_ptr(), beta.data_ptr(), 1e-5, 1, S), "ln")
        b200.check(b200.tf_gemm_f16(hn.data_ptr(), N, W2.data_ptr(), N, out.data_ptr(), K, M, K, N, None, None, 0, 0, ws.data_ptr(), ws.numel(), S), "gemm2")
    torch.cuda.synchronize()
    chain()
    side.synchronize()
    eager = out.clone()
    exec_ = ctypes.c_void_p()
    b200.check(b200.tf_graph_begin_capture(S), "begin")
    chain()
    b200.check(b200.tf_graph_end_capture(S, ctypes.byref(exec_)), "end")
    try:
        out.zero_()
        torch.cuda.synchronize()
        b200.check(b200.tf_graph_launch(exec_, S), "launch")
        side.synchronize()
        assert torch.equal(out, eager)
        ref = torch.nn.functional.layer_norm((A.float() @ W1.float().T).half().float(), (N,)).half().float() @ W2.float().T
        assert rel_err(out, ref) < TOL
        A.mul_(2.0)                       # same graph, new input contents
        torch.cuda.synchronize()
        b200.check(b200.tf_graph_launch(exec_, S), "launch")
        side.synchronize()
        assert not torch.equal(out, eager)
        assert rel_err(out, ref) < TOL    # LayerNorm makes the chain scale-invariant
    finally:
        b200.check(b200.tf_graph_destroy(exec_), "destroy")
