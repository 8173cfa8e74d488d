"""Shared by tools/autotune_gemm.py and tools/dev_clusterk.py (GPU box): build a GEMM / conv launch from a shape key of the
step's profile (runtime.Context._timed keys) with rotating cold weight copies, and time it as a captured chain."""
import ctypes, math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinyfusers_b200.native.b200.ops import b200

dev = torch.device("cuda:0"); b200.init(0)
S = lambda: torch.cuda.current_stream().cuda_stream
ws = torch.empty(192 << 20, dtype=torch.uint8, device=dev)
N_CHAIN = 16


def chain_us(fns):
    g = torch.cuda.CUDAGraph(); side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for i in range(N_CHAIN): fns[i % len(fns)]()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000 / (3 * N_CHAIN)


def last_choice():
    a, b, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    b200.tf_gemm_last_choice(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
    return a.value, b.value, c.value


def make_case(key):
    """-> (launchers over weight copies, out tensor, table key (is_conv, M, N, K, klass), m_tiles, k_blocks, bn_mult, allow_split)"""
    if key[0] == "gemm":
        _, M, N, K, flags, has_res, gn_unit, gn_hw = key[:8]
        wbytes = N * K * 2
        copies = max(2, min(16, math.ceil(300e6 / wbytes)))
        A = torch.randn(M, K, device=dev).half()
        Ws = [(torch.randn(N, K, device=dev) / math.sqrt(K)).half() for _ in range(copies)]
        bias = torch.randn(N, device=dev)
        No = N // 2 if flags & 2 else N
        out = torch.empty(M, No, dtype=torch.float32 if flags & 1 else torch.half, device=dev)
        res = torch.randn(M, No, device=dev).half() if has_res else None
        st = torch.zeros(max(1, (M // 32) * (N // gn_unit) * 2), device=dev) if gn_unit else None
        def mk(W):
            if gn_unit:
                return lambda: b200.check(b200.tf_gemm_gn_f16(A.data_ptr(), K, W.data_ptr(), K, out.data_ptr(), No, M, N, K, bias.data_ptr(), res.data_ptr() if has_res else None, No, flags, ws.data_ptr(), ws.numel(), st.data_ptr(), gn_unit, gn_hw, S()), "gemm")
            return lambda: b200.check(b200.tf_gemm_f16(A.data_ptr(), K, W.data_ptr(), K, out.data_ptr(), No, M, N, K, bias.data_ptr(), res.data_ptr() if has_res else None, No, flags, ws.data_ptr(), ws.numel(), S()), "gemm")
        klass = (flags & 3) | (4 if gn_unit else 0) | (8 if has_res else 0)
        keep = (A, Ws, bias, out, res, st)
        return [mk(W) for W in Ws], out, (0, M, N, K, klass), (M + 127) // 128, (K + 63) // 64, (math.lcm(32, gn_unit) if gn_unit else 32), not (flags & 2), keep
    _, n, h, w, cin, cout, stride, has_res, gn_unit = key[:9]
    c2 = key[9] if len(key) > 9 else 0     # channels of the appended 1x1 skip convolution's source
    ho, wo = (h + 2 - 3) // stride + 1, (w + 2 - 3) // stride + 1
    M, K = n * ho * wo, 9 * cin + c2
    wbytes = cout * K * 2
    copies = max(2, min(16, math.ceil(300e6 / wbytes)))
    x = torch.randn(n, h, w, cin, device=dev).half()
    Ws = [(torch.randn(cout, K, device=dev) / math.sqrt(K)).half() for _ in range(copies)]
    x2 = torch.randn(n, h, w, c2, device=dev).half() if c2 else None
    bias = torch.randn(cout, device=dev)
    f32 = cout <= 16
    ldc = 16 if cout == 8 and cin == 320 else cout
    out = torch.empty(n, ho, wo, ldc, dtype=torch.float32 if f32 else torch.half, device=dev)
    res = torch.randn(n, ho, wo, cout, device=dev).half() if has_res else None
    st = torch.zeros(max(1, n * (ho * wo // 32) * (cout // gn_unit) * 2), device=dev) if gn_unit else None
    flags = 1 if f32 else 0
    def mk(W):
        if c2:
            return lambda: b200.check(b200.tf_conv2d_nhwc_skip_f16(x.data_ptr(), n, h, w, cin, cin, x2.data_ptr(), c2, c2, W.data_ptr(), cout, out.data_ptr(), ldc, bias.data_ptr(), flags, ws.data_ptr(), ws.numel(), st.data_ptr() if gn_unit else None, gn_unit, S()), "conv+skip")
        if gn_unit:
            return lambda: b200.check(b200.tf_conv2d_nhwc_gn_f16(x.data_ptr(), n, h, w, cin, cin, W.data_ptr(), cout, 3, stride, out.data_ptr(), ldc, bias.data_ptr(), res.data_ptr() if has_res else None, cout, flags, ws.data_ptr(), ws.numel(), st.data_ptr(), gn_unit, S()), "conv")
        return lambda: b200.check(b200.tf_conv2d_nhwc_f16(x.data_ptr(), n, h, w, cin, cin, W.data_ptr(), cout, 3, stride, out.data_ptr(), ldc, bias.data_ptr(), res.data_ptr() if has_res else None, cout, flags, ws.data_ptr(), ws.numel(), S()), "conv")
    klass = flags | (4 if gn_unit else 0) | (8 if has_res else 0) | (16 if stride == 2 else 0)
    keep = (x, Ws, bias, out, res, st, x2)
    return [mk(W) for W in Ws], out, (1, M, cout, K, klass), (M + 127) // 128, K // 64, (math.lcm(32, gn_unit) if gn_unit else 32), True, keep


