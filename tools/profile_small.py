"""Focused ncu target (GPU box): the small memory-bound kernels + two GEMM shapes, 3 launches each."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinyfusers_b200.native.b200.ops import b200
from tinyfusers_b200.runtime import gn_unit
dev = torch.device("cuda:0"); b200.init(0)
S = lambda: torch.cuda.current_stream().cuda_stream
ws = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
def ln(rows, C):
    x = torch.randn(rows, C, device=dev).half(); o = torch.empty_like(x); g = torch.ones(C, device=dev); b = torch.zeros(C, device=dev)
    for _ in range(3): b200.check(b200.tf_layernorm_f16(x.data_ptr(), o.data_ptr(), rows, C, g.data_ptr(), b.data_ptr(), 1e-5, 1, S()), "ln")
def gn(n, hw, C):
    x = torch.randn(n, hw, C, device=dev).half(); o = torch.empty_like(x); g = torch.ones(C, device=dev); b = torch.zeros(C, device=dev)
    u = gn_unit(C); st = torch.randn(n, hw // 32, C // u, 2, device=dev).abs()
    for _ in range(3): b200.check(b200.tf_groupnorm_fused_nhwc_f16(x.data_ptr(), C, C, st.data_ptr(), u, None, 0, 0, None, 1, o.data_ptr(), C, n, hw, 32, g.data_ptr(), b.data_ptr(), 1e-5, 1, S()), "gn")
def gemm(M, Nn, K):
    A = torch.randn(M, K, device=dev).half(); W = (torch.randn(Nn, K, device=dev) / 30).half(); bias = torch.randn(Nn, device=dev)
    out = torch.empty(M, Nn, dtype=torch.half, device=dev); r = torch.randn(M, Nn, device=dev).half()
    for _ in range(3): b200.check(b200.tf_gemm_f16(A.data_ptr(), K, W.data_ptr(), K, out.data_ptr(), Nn, M, Nn, K, bias.data_ptr(), r.data_ptr(), Nn, 0, ws.data_ptr(), ws.numel(), S()), "gemm")
ln(8192, 320); ln(512, 1280); gn(2, 4096, 320); gn(2, 256, 1280); gemm(8192, 320, 320); gemm(512, 1280, 1280)
torch.cuda.synchronize(); print("done")
