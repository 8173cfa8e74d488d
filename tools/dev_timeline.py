"""Dev: per-CTA clock stamps of the GEMM kernel for a few UNet shapes (where does a short kernel spend its time?)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinyfusers_b200.native.b200.ops import b200
dev = torch.device("cuda:0"); b200.init(0)
S = lambda: torch.cuda.current_stream().cuda_stream
ws = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
tl = torch.zeros(148 * 16 + 512, dtype=torch.int64, device=dev)
def report(name, grid):
    torch.cuda.synchronize()
    t = tl[:148 * 16].view(148, 16)[:grid:2].cpu().double()   # even CTAs (the leaders in pair mode)
    e = t[:, 7:8]
    names = ["setup_done", "mma_start", "first_operands", "mma_issued", "acc_ready", "stored", "exit"]
    d = (t[:, :7] - e)
    ex = {k: int((t[:, i] - e[:, 0]).median().item()) for k, i in (("c0_done", 12), ("c1_tmem_loaded", 8), ("c1_staged", 9), ("c1_lds_done", 10), ("c1_done", 11))}
    tr = tl[148 * 16:].cpu()
    nk = int((tr[256:] > 0).sum())
    print("   issue:", [int(v) for v in tr[:nk][:48]])
    print("   full :", [int(v) for v in tr[256:256 + nk][:48]])
    tl.zero_()
    print(name, "grid", grid, " median cycles since kernel entry:", {n: int(d[:, i].median().item()) for i, n in enumerate(names)}, ex)
def gemm(M, N, K, residual=False, geglu=False, dbg=0):
    A = torch.randn(M, K, device=dev).half(); W = (torch.randn(N, K, device=dev) / 30).half(); b = torch.randn(N, device=dev)
    No = N // 2 if geglu else N
    R = torch.randn(M, No, device=dev).half() if residual else None
    out = torch.empty(M, No, dtype=torch.half, device=dev)
    for i in range(3):
        b200.tf_gemm_set_timeline(tl.data_ptr() if i == 2 else None)
        b200.check(b200.tf_gemm_f16(A.data_ptr(), K, W.data_ptr(), K, out.data_ptr(), No, M, N, K, b.data_ptr(), R.data_ptr() if residual else None, No, (2 if geglu else 0) | dbg, ws.data_ptr(), ws.numel(), S()), "gemm")
    b200.tf_gemm_set_timeline(None)
    report(f"gemm {M}x{N}x{K} res={residual} geglu={geglu} dbg={hex(dbg)}", min(148, (M + 127) // 128 * max(1, N // 160)))
def conv(NI, H, W, Cin, Cout):
    x = torch.randn(NI, H, W, Cin, device=dev).half(); w = (torch.randn(Cout, 3, 3, Cin, device=dev) / 50).half()
    b = torch.randn(Cout, device=dev); out = torch.empty(NI, H, W, Cout, dtype=torch.half, device=dev)
    for i in range(3):
        b200.tf_gemm_set_timeline(tl.data_ptr() if i == 2 else None)
        b200.check(b200.tf_conv2d_nhwc_f16(x.data_ptr(), NI, H, W, Cin, Cin, w.data_ptr(), Cout, 3, 1, out.data_ptr(), Cout, b.data_ptr(), None, 0, 0, ws.data_ptr(), ws.numel(), S()), "conv")
    b200.tf_gemm_set_timeline(None)
    report(f"conv {NI}x{H}x{W} {Cin}->{Cout}", 128)
def cadence(name):
    torch.cuda.synchronize()
    tr = tl[148 * 16:].cpu()
    nk = int((tr[256:] > 0).sum())
    full = [int(v) for v in tr[256:256 + nk]]
    iss = [int(v) for v in tr[:nk]]
    if nk > 8:
        print(f"{name}: kbs={nk} full cadence {(full[-1] - full[4]) / (nk - 5):.0f} cyc/kb, issue cadence {(iss[-1] - iss[4]) / (nk - 5):.0f}, first full {full[0]}, issue->full latency {full[0] - iss[0]}")
    tl.zero_()
def gemm_c(M, N, K, bn, ctas):
    b200.tf_gemm_set_ctas(ctas); b200.tf_gemm_set_tuning(bn, 1)
    A = torch.randn(M, K, device=dev).half(); W = (torch.randn(N, K, device=dev) / 30).half(); b = torch.randn(N, device=dev)
    out = torch.empty(M, N, dtype=torch.half, device=dev)
    for i in range(3):
        b200.tf_gemm_set_timeline(tl.data_ptr() if i == 2 else None)
        b200.check(b200.tf_gemm_f16(A.data_ptr(), K, W.data_ptr(), K, out.data_ptr(), N, M, N, K, b.data_ptr(), None, 0, 0, ws.data_ptr(), ws.numel(), S()), "gemm")
    b200.tf_gemm_set_timeline(None)
    cadence(f"gemm {M}x{N}x{K} bn={bn} ctas={ctas}")
def conv_c(NI, H, W, Cin, Cout, bn, ctas):
    b200.tf_gemm_set_ctas(ctas); b200.tf_gemm_set_tuning(bn, 1)
    x = torch.randn(NI, H, W, Cin, device=dev).half(); w = (torch.randn(Cout, 3, 3, Cin, device=dev) / 50).half()
    b = torch.randn(Cout, device=dev); out = torch.empty(NI, H, W, Cout, dtype=torch.half, device=dev)
    for i in range(3):
        b200.tf_gemm_set_timeline(tl.data_ptr() if i == 2 else None)
        b200.check(b200.tf_conv2d_nhwc_f16(x.data_ptr(), NI, H, W, Cin, Cin, w.data_ptr(), Cout, 3, 1, out.data_ptr(), Cout, b.data_ptr(), None, 0, 0, ws.data_ptr(), ws.numel(), S()), "conv")
    b200.tf_gemm_set_timeline(None)
    cadence(f"conv {NI}x{H}x{W} {Cin}->{Cout} bn={bn} ctas={ctas}")
b200.tf_gemm_set_ctas(0); b200.tf_gemm_set_tuning(0, 0)
def epi(name, grid):
    torch.cuda.synchronize()
    t = tl[:148 * 16].view(148, 16)[:grid:2].cpu().double()
    e = t[:, 7]
    names = {0: "setup_done", 1: "mma_start", 2: "first_full", 3: "mma_issued", 4: "acc_ready", 8: "c0_tmem", 9: "c0_staged", 10: "c0_tma", 11: "cN_tmem", 12: "cN_tma", 5: "stored", 6: "exit"}
    print(name, {v: int((t[:, k] - e).median().item()) for k, v in names.items()})
    tl.zero_()
def gemm_e(M, N, K, residual):
    A = torch.randn(M, K, device=dev).half(); W = (torch.randn(N, K, device=dev) / 30).half(); b = torch.randn(N, device=dev)
    R = torch.randn(M, N, device=dev).half()
    out = torch.empty(M, N, dtype=torch.half, device=dev)
    for i in range(3):
        b200.tf_gemm_set_timeline(tl.data_ptr() if i == 2 else None)
        b200.check(b200.tf_gemm_f16(A.data_ptr(), K, W.data_ptr(), K, out.data_ptr(), N, M, N, K, b.data_ptr(), R.data_ptr() if residual else None, N, 0, ws.data_ptr(), ws.numel(), S()), "gemm")
    b200.tf_gemm_set_timeline(None)
    epi(f"gemm {M}x{N}x{K} res={residual}", 128)
gemm_e(8192, 320, 320, True)
gemm_e(8192, 320, 320, False)
gemm_e(8192, 320, 1280, True)
