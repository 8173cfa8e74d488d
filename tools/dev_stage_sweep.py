"""Dev (GPU box): is the GEMM / conv main loop bound by ring depth x TMA latency, or by operand feed?
In-graph time (chain of 16 launches, cold weight copies) of a few shapes with the smem ring capped at 2..8 stages."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinyfusers_b200.native.b200.ops import b200
dev = torch.device("cuda:0"); b200.init(0)
S = lambda: torch.cuda.current_stream().cuda_stream
ws = torch.empty(192 << 20, dtype=torch.uint8, device=dev)
def chain(fns, N=16):
    g = torch.cuda.CUDAGraph(); side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for i in range(N): fns[i % len(fns)]()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000 / (3 * N)
def conv(NI, H, W, Cin, Cout):
    x = torch.randn(NI, H, W, Cin, device=dev).half()
    copies = max(2, min(16, math.ceil(300e6 / (Cout * 9 * Cin * 2))))
    Wt = [(torch.randn(Cout, 3, 3, Cin, device=dev) / 50).half() for _ in range(copies)]
    b = torch.randn(Cout, device=dev); out = torch.empty(NI, H, W, Cout, dtype=torch.half, device=dev)
    fns = [(lambda w: (lambda: b200.check(b200.tf_conv2d_nhwc_f16(x.data_ptr(), NI, H, W, Cin, Cin, w.data_ptr(), Cout, 3, 1, out.data_ptr(), Cout, b.data_ptr(), None, 0, 4, ws.data_ptr(), ws.numel(), S()), "conv")))(w) for w in Wt]
    res = []
    for cap in (2, 3, 4, 5, 6, 7, 8):
        b200.tf_gemm_set_max_stages(cap); res.append(f"{cap}: {chain(fns):6.2f}")
    b200.tf_gemm_set_max_stages(0)
    print(f"conv3x3 {NI}x{H}x{W} {Cin}->{Cout}  us by stage cap  " + "  ".join(res), flush=True)
def gemm(M, N, K):
    A = torch.randn(M, K, device=dev).half()
    copies = max(2, min(16, math.ceil(300e6 / (N * K * 2))))
    Wt = [(torch.randn(N, K, device=dev) / 30).half() for _ in range(copies)]
    b = torch.randn(N, device=dev); out = torch.empty(M, N, dtype=torch.half, device=dev)
    fns = [(lambda w: (lambda: b200.check(b200.tf_gemm_f16(A.data_ptr(), K, w.data_ptr(), K, out.data_ptr(), N, M, N, K, b.data_ptr(), None, 0, 4, ws.data_ptr(), ws.numel(), S()), "gemm")))(w) for w in Wt]
    res = []
    for cap in (2, 3, 4, 5, 6, 7, 8):
        b200.tf_gemm_set_max_stages(cap); res.append(f"{cap}: {chain(fns):6.2f}")
    b200.tf_gemm_set_max_stages(0)
    print(f"gemm {M}x{N}x{K}  us by stage cap  " + "  ".join(res), flush=True)
conv(2, 64, 64, 320, 320); conv(2, 64, 64, 640, 640); conv(2, 32, 32, 640, 640); conv(2, 16, 16, 1280, 1280); conv(2, 8, 8, 1280, 1280)
conv(16, 64, 64, 320, 320); conv(1, 512, 512, 128, 128)
gemm(8192, 320, 1280); gemm(512, 1280, 5120); gemm(2048, 640, 2560)
