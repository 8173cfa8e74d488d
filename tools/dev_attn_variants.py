"""Dev (GPU box): in-graph time of the attention kernel variants (key-block width x CTAs per SM) on the UNet's shapes."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinyfusers_b200.native.b200.ops import b200
dev = torch.device("cuda:0"); b200.init(0)
S = lambda: torch.cuda.current_stream().cuda_stream
def chain(fn, N=16):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(N): fn()
    for _ in range(2): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000 / (5 * N)
def run(B, NH, T, d, dp):
    q = torch.randn(B * T, NH * dp, device=dev).half(); k = torch.randn(B * T, NH * dp, device=dev).half()
    vt = torch.randn(NH * dp, B * T, device=dev).half(); out = torch.empty(B * T, NH * d, device=dev).half()
    ref = None
    for tune in (0, 2064, 3064, 2128):
        b200.tf_attention_set_tuning(tune)
        fn = lambda: b200.check(b200.tf_attention_f16(q.data_ptr(), NH * dp, k.data_ptr(), NH * dp, vt.data_ptr(), B * T, out.data_ptr(), T * NH * d, d, NH * d, B, NH, T, T, T, d, dp, 1 / math.sqrt(d), S()), "attn")
        try:
            us = chain(fn)
        except RuntimeError as e:
            print(f"  B={B} NH={NH} T={T} d={d} tune={tune}: {str(e)[:80]}"); continue
        o = out.float().clone()
        if ref is None: ref = o
        print(f"  B={B} NH={NH} T={T} d={d} tune={tune:4d}: {us:7.2f} us   {4.0 * B * NH * T * T * d / us / 1e6:6.1f} TFLOP/s  maxdiff vs auto {float((o - ref).abs().max()):.2e}")
    b200.tf_attention_set_tuning(0)
run(2, 8, 4096, 40, 48)
run(16, 8, 4096, 40, 48)
run(2, 8, 1024, 80, 80)
run(8, 8, 9216, 40, 48)
