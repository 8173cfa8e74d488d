"""Dev (GPU box): tf_attention_kernel (one query tile per CTA) vs tf_attention2_kernel (two tiles in ping-pong, P in tensor
memory, EMU of 8 exponentials on the FMA pipe): error against torch SDPA on the same inputs and in-graph time per launch."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinyfusers_b200.native.b200.ops import b200
from tinyfusers_b200.attention.attention import _pad64
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
def run(B, NH, T, d, causal=False, scale_in=1.0, variants=((1, 0, 0), (2, 0, 0), (2, 2, 0), (2, 2, 1))):
    dp = (d + 15) // 16 * 16; dvp = _pad64(d)
    g = torch.Generator(device="cuda").manual_seed(T + d)
    q = torch.randn(B, T, NH, d, generator=g, device=dev) * scale_in
    k = torch.randn(B, T, NH, d, generator=g, device=dev) * scale_in
    v = torch.randn(B, T, NH, d, generator=g, device=dev)
    Q = torch.zeros(B, T, NH, dp, dtype=torch.half, device=dev); Q[..., :d] = q
    K = torch.zeros(B, T, NH, dp, dtype=torch.half, device=dev); K[..., :d] = k
    V = torch.zeros(B, T, NH, dvp, dtype=torch.half, device=dev); V[..., :d] = v
    V1 = V.clone()
    if dvp > d: V1[..., d] = 1.0       # TF_ATTN_V_ONES_COLUMN: the row sums come out of P V
    out = torch.zeros(B, T, NH, d, dtype=torch.half, device=dev)
    heads = lambda t: t.half().float().permute(0, 2, 1, 3)
    ref = torch.nn.functional.scaled_dot_product_attention(heads(q), heads(k), heads(v), is_causal=causal).permute(0, 2, 1, 3)
    for var in variants:
        ver, emu, ones = (tuple(var) + (0,))[:3]
        if ones and dvp <= d: continue
        Vx = V1 if ones else V
        b200.check(b200.tf_attention_set_variant(ver, emu), "variant")
        out.zero_()
        fn = lambda: b200.check(b200.tf_attention_v_f16(Q.data_ptr(), NH * dp, K.data_ptr(), NH * dp, Vx.data_ptr(), NH * dvp, out.data_ptr(),
                                 T * NH * d, d, NH * d, B, NH, T, T, T, d, dp, dvp, 1.0 / math.sqrt(d), (1 if causal else 0) | (2 if ones else 0), S()), "attn")
        try:
            us = chain(fn)
        except RuntimeError as e:
            print(f"  B={B} NH={NH} T={T} d={d} v{ver} emu{emu} ones{ones}: FAILED {str(e)[:120]}", flush=True); continue
        err = float((out.float() - ref).abs().max() / ref.abs().max())
        print(f"  B={B} NH={NH} T={T} d={d} causal={int(causal)} in_scale={scale_in} v{ver} emu{emu} ones{ones}: {us:8.2f} us  {4.0 * B * NH * T * T * d / us / 1e6 / (2 if causal else 1):6.1f} TFLOP/s  rel_err {err:.2e}", flush=True)
    b200.tf_attention_set_variant(0, -1)
if __name__ == "__main__":
    # argv: "ver,emu ver,emu ..." then optional "quick"
    var = tuple(tuple(int(x) for x in a.split(",")) for a in sys.argv[1:] if "," in a) or ((1, 0, 0), (1, 0, 1), (2, 0, 0), (2, 2, 0), (2, 0, 1), (2, 2, 1), (2, 4, 1), (2, 12, 1))
    quick = "quick" in sys.argv
    run(2, 8, 4096, 40, variants=var)
    if not quick:
        run(2, 8, 4096, 40, scale_in=3.0, variants=var)      # peaked rows: exercises the lazy rescale
        run(2, 8, 1024, 80, variants=var)
        run(16, 8, 4096, 40, variants=var)
        run(16, 8, 1024, 80, variants=var)
        run(1, 8, 512, 64, causal=True, variants=var)
        run(8, 8, 9216, 40, variants=var[:1] + var[2:3])
