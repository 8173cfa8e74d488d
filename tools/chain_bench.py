"""Dev (GPU box): in-graph cost of ONE launch of the small kernels, measured as a captured chain of 64 identical
dependent launches (CUDA events around graph replays). Tells the launch floor from the kernel bodies."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinyfusers_b200.native.b200.ops import b200
from tinyfusers_b200.runtime import gn_unit
dev = torch.device("cuda:0"); b200.init(0)
S = lambda: torch.cuda.current_stream().cuda_stream
ws = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
N = 64

def chain(name, fn):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(N):
                fn()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:60s} {e0.elapsed_time(e1) * 1000 / (10 * N):7.2f} us / launch", flush=True)

cnt = torch.zeros(1, dtype=torch.int32, device=dev)
chain("tf_add_int (1 thread)", lambda: b200.check(b200.tf_add_int(cnt.data_ptr(), 1, S()), "add"))

def ln(rows, C):
    x = torch.randn(rows, C, device=dev).half(); o = torch.empty_like(x); g = torch.ones(C, device=dev); b = torch.zeros(C, device=dev)
    chain(f"layernorm rows={rows} C={C}", lambda: b200.check(b200.tf_layernorm_f16(x.data_ptr(), o.data_ptr(), rows, C, g.data_ptr(), b.data_ptr(), 1e-5, 1, S()), "ln"))
for r, c in ((8192, 320), (2048, 640), (512, 1280), (128, 1280)): ln(r, c)

def gn(n, hw, C, fused=True):
    x = torch.randn(n, hw, C, device=dev).half(); o = torch.empty_like(x); g = torch.ones(C, device=dev); b = torch.zeros(C, device=dev)
    u = gn_unit(C)
    st = torch.randn(n, hw // 32, C // u, 2, device=dev).abs()
    wsb = torch.empty(b200.tf_groupnorm_workspace_bytes(n, 32), dtype=torch.uint8, device=dev)
    if fused:
        chain(f"groupnorm fused n={n} hw={hw} C={C}", lambda: b200.check(b200.tf_groupnorm_fused_nhwc_f16(x.data_ptr(), C, C, st.data_ptr(), u, None, 0, 0, None, 1, o.data_ptr(), C, n, hw, 32, g.data_ptr(), b.data_ptr(), 1e-5, 1, S()), "gn"))
    else:
        chain(f"groupnorm 3-launch n={n} hw={hw} C={C} (per call)", lambda: b200.check(b200.tf_groupnorm_nhwc_f16(x.data_ptr(), C, C, None, 0, 0, o.data_ptr(), C, n, hw, 32, g.data_ptr(), b.data_ptr(), 1e-5, 1, wsb.data_ptr(), S()), "gn"))
for n, hw, c in ((2, 4096, 320), (2, 1024, 640), (2, 256, 1280), (2, 64, 1280), (2, 64, 2560), (2, 4096, 960)): gn(n, hw, c)
gn(2, 4096, 320, fused=False)

def gemm(M, Nn, K, res=False, gnst=False):
    A = torch.randn(M, K, device=dev).half(); W = (torch.randn(Nn, K, device=dev) / 30).half(); bias = torch.randn(Nn, device=dev)
    out = torch.empty(M, Nn, dtype=torch.half, device=dev); r = torch.randn(M, Nn, device=dev).half()
    chain(f"gemm M={M} N={Nn} K={K} res={res}", lambda: b200.check(b200.tf_gemm_f16(A.data_ptr(), K, W.data_ptr(), K, out.data_ptr(), Nn, M, Nn, K, bias.data_ptr(), r.data_ptr() if res else None, Nn, 0, ws.data_ptr(), ws.numel(), S()), "gemm"))
for m, n, k in ((8192, 320, 320), (2048, 640, 640), (512, 1280, 1280), (128, 1280, 1280), (8192, 768, 320), (8192, 320, 1280), (512, 1280, 5120)): gemm(m, n, k, res=True)

def conv(NI, H, W_, Cin, Cout, gnst=False):
    x = torch.randn(NI, H, W_, Cin, device=dev).half(); w = (torch.randn(Cout, 3, 3, Cin, device=dev) / 50).half()
    b = torch.randn(Cout, device=dev); out = torch.empty(NI, H, W_, Cout, dtype=torch.half, device=dev)
    u = gn_unit(Cout); st = torch.empty(NI, H * W_ // 32, Cout // u, 2, device=dev)
    if gnst:
        chain(f"conv3x3+stats {NI}x{H}x{W_} {Cin}->{Cout}", lambda: b200.check(b200.tf_conv2d_nhwc_gn_f16(x.data_ptr(), NI, H, W_, Cin, Cin, w.data_ptr(), Cout, 3, 1, out.data_ptr(), Cout, b.data_ptr(), None, 0, 0, ws.data_ptr(), ws.numel(), st.data_ptr(), u, S()), "conv"))
    else:
        chain(f"conv3x3 {NI}x{H}x{W_} {Cin}->{Cout}", lambda: b200.check(b200.tf_conv2d_nhwc_f16(x.data_ptr(), NI, H, W_, Cin, Cin, w.data_ptr(), Cout, 3, 1, out.data_ptr(), Cout, b.data_ptr(), None, 0, 0, ws.data_ptr(), ws.numel(), S()), "conv"))
for a in ((2, 64, 64, 320, 320), (2, 32, 32, 640, 640), (2, 16, 16, 1280, 1280), (2, 8, 8, 1280, 1280)):
    conv(*a); conv(*a, gnst=True)

def attn(B, NH, T, Tk, d):
    dp = (d + 15) // 16 * 16; Tkp = (Tk + 7) // 8 * 8
    Q = torch.randn(B * T, NH * dp, device=dev).half(); K = torch.randn(B * Tkp, NH * dp, device=dev).half(); Vt = torch.randn(NH * dp, B * Tkp, device=dev).half()
    out = torch.empty(B, NH, T, d, dtype=torch.half, device=dev)
    chain(f"attention B={B} NH={NH} T={T} Tk={Tk} d={d}", lambda: b200.check(b200.tf_attention_f16(Q.data_ptr(), NH * dp, K.data_ptr(), NH * dp, Vt.data_ptr(), B * Tkp, out.data_ptr(), NH * T * d, T * d, d, B, NH, T, Tk, Tkp, d, dp, 1 / math.sqrt(d), S()), "attn"))
for a in ((2, 8, 4096, 4096, 40), (2, 8, 1024, 1024, 80), (2, 8, 256, 256, 160), (2, 8, 4096, 77, 40), (2, 8, 1024, 77, 80), (2, 8, 256, 77, 160), (2, 8, 64, 64, 160)): attn(*a)
