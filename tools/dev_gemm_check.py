"""Dev script (GPU box): first-light check of the tcgen05 GEMM / implicit-GEMM conv against torch.
Not part of the test-suite (tests/ use the oracle); prints a table and writes gpurun_out/dev_gemm.json."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinyfusers_b200.native.b200.ops import b200

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
b200.init(0)
stream = lambda: torch.cuda.current_stream().cuda_stream
ws = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
results = []

def gelu_tanh(x):
    return 0.5 * x * (1 + torch.tanh(x * 0.7978845608 * (1 + 0.044715 * x * x)))

def run_gemm(M, N, K, bias=True, residual=False, geglu=False, out_f32=False, bn=0, splits=0, time_it=False):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    A = (torch.randn(M, K, generator=g) ).half().to(dev)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).half().to(dev)
    b = torch.randn(N, generator=g).float().to(dev) if bias else None
    No = N // 2 if geglu else N
    R = torch.randn(M, No, generator=g).half().to(dev) if residual else None
    out = torch.full((M, No), float("nan"), dtype=torch.float32 if out_f32 else torch.float16, device=dev)
    flags = (1 if out_f32 else 0) | (2 if geglu else 0)
    b200.tf_gemm_set_tuning(bn, splits)
    def call():
        st = b200.tf_gemm_f16(A.data_ptr(), K, W.data_ptr(), K, out.data_ptr(), No, M, N, K,
                              b.data_ptr() if bias else None, R.data_ptr() if residual else None, No, flags,
                              ws.data_ptr(), ws.numel(), stream())
        b200.check(st, "tf_gemm_f16")
    call()
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t()
    if bias: ref = ref + b
    if geglu:
        r = ref.view(M, N // 32, 2, 16)
        ref = (r[:, :, 0] * gelu_tanh(r[:, :, 1])).reshape(M, No)
    if residual: ref = ref + R.float()
    err = (out.float() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-9)
    rec = dict(kind="gemm", M=M, N=N, K=K, bias=bias, residual=residual, geglu=geglu, out_f32=out_f32, bn=bn, splits=splits,
               rel_err=err, nan=bool(torch.isnan(out).any().item()))
    if time_it:
        for _ in range(3): call()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): call()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        rec["ms"] = ms; rec["tflops"] = 2.0 * M * N * K / ms / 1e9
    b200.tf_gemm_set_tuning(0, 0)
    results.append(rec); print(rec, flush=True)

def run_conv(NI, H, W, Cin, Cout, stride=1, bias=True, residual=False, bn=0, splits=0, time_it=False):
    g = torch.Generator(device="cpu").manual_seed(NI + H * 5 + Cin + Cout)
    x = torch.randn(NI, H, W, Cin, generator=g).half().to(dev)            # NHWC
    w = (torch.randn(Cout, 3, 3, Cin, generator=g) / (9 * Cin) ** 0.5).half().to(dev)  # OHWI
    b = torch.randn(Cout, generator=g).float().to(dev) if bias else None
    Ho, Wo = (H + 2 - 3) // stride + 1, (W + 2 - 3) // stride + 1
    R = torch.randn(NI, Ho, Wo, Cout, generator=g).half().to(dev) if residual else None
    out = torch.full((NI, Ho, Wo, Cout), float("nan"), dtype=torch.float16, device=dev)
    b200.tf_gemm_set_tuning(bn, splits)
    def call():
        st = b200.tf_conv2d_nhwc_f16(x.data_ptr(), NI, H, W, Cin, Cin, w.data_ptr(), Cout, 3, stride, out.data_ptr(), Cout,
                                     b.data_ptr() if bias else None, R.data_ptr() if residual else None, Cout, 0,
                                     ws.data_ptr(), ws.numel(), stream())
        b200.check(st, "tf_conv2d_nhwc_f16")
    call()
    torch.cuda.synchronize()
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), b, stride=stride, padding=1)
    ref = ref.permute(0, 2, 3, 1)
    if residual: ref = ref + R.float()
    err = (out.float() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-9)
    rec = dict(kind="conv3x3", NI=NI, H=H, W=W, Cin=Cin, Cout=Cout, stride=stride, residual=residual, bn=bn, splits=splits,
               rel_err=err, nan=bool(torch.isnan(out).any().item()))
    if time_it:
        for _ in range(3): call()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): call()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        rec["ms"] = ms; rec["tflops"] = 2.0 * NI * Ho * Wo * Cout * 9 * Cin / ms / 1e9
    b200.tf_gemm_set_tuning(0, 0)
    results.append(rec); print(rec, flush=True)

try:
    run_gemm(128, 128, 64, bias=False)
    run_gemm(128, 128, 256, bias=False)
    run_gemm(256, 320, 320)
    run_gemm(100, 72, 200)
    run_gemm(8192, 320, 320, residual=True, time_it=True)
    run_gemm(8192, 2560, 320, geglu=True, time_it=True)
    run_gemm(8192, 2560, 320, time_it=True)
    run_gemm(8192, 320, 1280, residual=True, time_it=True)
    run_gemm(2048, 640, 2560, residual=True, time_it=True)
    run_gemm(160, 320, 768, bias=False)
    run_gemm(128, 1280, 11520, splits=4)
    run_gemm(128, 1280, 11520, time_it=True)
    run_gemm(512, 16, 2880, out_f32=True)
    run_gemm(4096, 4096, 4096, bias=False, bn=256, time_it=True)
    run_gemm(4096, 4096, 4096, bias=False, bn=128, time_it=True)
    run_conv(1, 12, 12, 64, 32)
    run_conv(3, 20, 28, 128, 48)
    run_conv(2, 64, 64, 320, 320, time_it=True)
    run_conv(2, 64, 64, 320, 320, residual=True)
    run_conv(2, 32, 32, 640, 640, time_it=True)
    run_conv(2, 16, 16, 1280, 1280, time_it=True)
    run_conv(2, 8, 8, 1280, 1280, time_it=True)
    run_conv(2, 8, 8, 1280, 1280, splits=1, time_it=True)
    run_conv(2, 64, 64, 960, 320, time_it=True)
    run_conv(2, 64, 64, 320, 320, stride=2)
    run_conv(1, 10, 14, 64, 64, stride=2)
    run_conv(2, 16, 16, 1280, 1280, stride=2, time_it=True)
finally:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(results, open("gpurun_out/dev_gemm.json", "w"), indent=1)
bad = [r for r in results if r["rel_err"] > 5e-3 or r["nan"]]
print("FAILED CASES:", len(bad))
for r in bad: print("  ", r)
