"""north_star: "correctness is checked against the reference's own cuDNN/cuBLAS path on identical random-init weights and
inputs". The reference reaches cuDNN (conv_fprop, layernorm) and cuBLAS (SGEMM) through Python bindings that do not install
offline; torch on the GPU box calls the same libraries, so these tests hold the B200 kernels to cuDNN / cuBLAS results
computed ON THE SAME GPU in fp32 (TF32 off, as the reference's dtype) at the UNet's real layer sizes - the oracle
(oracle/ref_ops.py) stays the primary checker, this is the library-side cross-check. Tolerance: max|a-b| / max|b| <= 1e-2."""
import math

import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fp32_libraries():
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


@pytest.mark.parametrize("n,cin,cout,hw,stride", [(2, 320, 320, 64, 1), (2, 640, 1280, 16, 1), (2, 1280, 1280, 8, 1),
                                                  (2, 320, 320, 64, 2), (1, 128, 128, 256, 1)])
def test_conv3x3_vs_cudnn(n, cin, cout, hw, stride):
    from tinyfusers_b200.vision.conv2d import Conv2d
    g = torch.Generator().manual_seed(cin + hw)
    x = torch.randn(n, cin, hw, hw, generator=g).cuda()
    conv = Conv2d(cin, cout, kernel_size=[3, 3], stride=[stride, stride], padding=[1, 1])
    conv.weight = (torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(9 * cin)).cuda()
    conv.bias = (0.02 * torch.randn(cout, generator=g)).cuda()
    ref = F.conv2d(x, conv.weight, conv.bias, stride=stride, padding=1)       # cuDNN, fp32
    assert rel_err(conv(x), ref) < 1e-2


@pytest.mark.parametrize("M,N,K", [(8192, 320, 320), (2048, 640, 2560), (512, 1280, 5120), (77, 768, 768)])
def test_linear_vs_cublas(M, N, K):
    from tinyfusers_b200.ff.linear import Linear
    g = torch.Generator().manual_seed(M + N)
    x = torch.randn(M, K, generator=g).cuda()
    lin = Linear(K, N)
    lin.weight = (torch.randn(N, K, generator=g) / math.sqrt(K)).cuda()
    lin.bias = (0.02 * torch.randn(N, generator=g)).cuda()
    ref = torch.addmm(lin.bias, x, lin.weight.t())                              # cuBLAS SGEMM, fp32
    assert rel_err(lin(x), ref) < 1e-2


@pytest.mark.parametrize("B,T,C", [(2, 4096, 320), (2, 256, 1280), (1, 77, 768)])
def test_layernorm_vs_cudnn_semantics(B, T, C):
    from tinyfusers_b200.ff.layer_norm import LayerNorm
    g = torch.Generator().manual_seed(C)
    x = (1.5 * torch.randn(B, T, C, generator=g) + 0.3).cuda()
    ln = LayerNorm(C)
    ln.weight = (1 + 0.1 * torch.randn(C, generator=g)).cuda()
    ln.bias = (0.1 * torch.randn(C, generator=g)).cuda()
    ref = F.layer_norm(x, (C,), ln.weight, ln.bias, 1e-5)      # what real cuDNN executes for the reference's graph (B = 1 probe)
    assert rel_err(ln(x), ref) < 1e-2


@pytest.mark.parametrize("B,T,Tk,d", [(2, 4096, 4096, 40), (2, 1024, 77, 80), (2, 256, 256, 160)])
def test_sdpa_vs_library(B, T, Tk, d):
    from tinyfusers_b200.attention.sdpa import scaled_dot_product_attention
    g = torch.Generator().manual_seed(T + d)
    q = torch.randn(B, 8, T, d, generator=g).cuda()
    k = torch.randn(B, 8, Tk, d, generator=g).cuda()
    v = torch.randn(B, 8, Tk, d, generator=g).cuda()
    ref = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(d), dim=-1) @ v     # two cuBLAS batched SGEMMs + softmax
    assert rel_err(scaled_dot_product_attention(q, k, v), ref) < 1e-2


def test_groupnorm_vs_library():
    from tinyfusers_b200.ff.group_norm import GroupNorm
    g = torch.Generator().manual_seed(11)
    x = (1.7 * torch.randn(2, 640, 32, 32, generator=g) + 0.6).cuda()
    gn = GroupNorm(32, 640)
    gn.weight = (1 + 0.1 * torch.randn(640, generator=g)).cuda()
    gn.bias = (0.1 * torch.randn(640, generator=g)).cuda()
    assert rel_err(gn(x), F.group_norm(x, 32, gn.weight, gn.bias, 1e-5)) < 1e-2


def test_sdpa_causal_mask_both_forms():
    """The reference's CLIP call: sdpa(q, k, v, attn_mask = triu(full(-inf), k=1)) (vae/encoder.py:79, sdpa.py:67-68)."""
    from tinyfusers_b200.attention.sdpa import scaled_dot_product_attention
    g = torch.Generator().manual_seed(5)
    q, k, v = (torch.randn(2, 12, 77, 64, generator=g).cuda() for _ in range(3))
    add = torch.triu(torch.full((1, 1, 77, 77), float("-inf")), diagonal=1).cuda()
    ref = torch.softmax((q @ k.transpose(-1, -2)) / 8.0 + add, dim=-1) @ v
    assert rel_err(scaled_dot_product_attention(q, k, v, attn_mask=add), ref) < 1e-2
    assert rel_err(scaled_dot_product_attention(q, k, v, attn_mask=(add == 0)), ref) < 1e-2
    with pytest.raises(RuntimeError, match="only the causal mask"):
        scaled_dot_product_attention(q, k, v, attn_mask=torch.zeros(1, 1, 77, 77).cuda().bool())
