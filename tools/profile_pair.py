"""ncu target (GPU box): the same GEMM / conv through single-CTA and CTA-pair tiles."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinyfusers_b200.native.b200.ops import b200
dev = torch.device("cuda:0"); b200.init(0)
S = lambda: torch.cuda.current_stream().cuda_stream
ws = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
def conv(NI, H, W, Cin, Cout, reps=2):
    x = torch.randn(NI, H, W, Cin, device=dev).half(); w = (torch.randn(Cout, 3, 3, Cin, device=dev) / 50).half()
    b = torch.randn(Cout, device=dev); out = torch.empty(NI, H, W, Cout, dtype=torch.half, device=dev)
    for _ in range(reps):
        b200.check(b200.tf_conv2d_nhwc_f16(x.data_ptr(), NI, H, W, Cin, Cin, w.data_ptr(), Cout, 3, 1, out.data_ptr(), Cout, b.data_ptr(), None, 0, 0, ws.data_ptr(), ws.numel(), S()), "conv")
def gemm(M, N, K, reps=2):
    A = torch.randn(M, K, device=dev).half(); W = (torch.randn(N, K, device=dev) / 30).half(); b = torch.randn(N, device=dev)
    out = torch.empty(M, N, dtype=torch.half, device=dev)
    for _ in range(reps):
        b200.check(b200.tf_gemm_f16(A.data_ptr(), K, W.data_ptr(), K, out.data_ptr(), N, M, N, K, b.data_ptr(), None, 0, 0, ws.data_ptr(), ws.numel(), S()), "gemm")
for ctas in (1, 2):
    b200.tf_gemm_set_ctas(ctas)
    gemm(8192, 320, 1280)
    conv(2, 64, 64, 320, 320)
torch.cuda.synchronize(); print("done")
