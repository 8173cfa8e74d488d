"""Dev script (GPU box): attention / norm / elementwise kernels against torch. Not part of tests/."""
import json, math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinyfusers_b200.native.b200.ops import b200

dev = torch.device("cuda:0")
b200.init(0)
S = lambda: torch.cuda.current_stream().cuda_stream
res = []
def relerr(a, b): return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-9)).item()
def timeit(fn, n=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def attn(B, NH, Tq, Tk, d, quirk=True, bn=0, time_it=False):
    dp = (d + 15) // 16 * 16
    Tk_pad = (Tk + 7) // 8 * 8
    g = torch.Generator().manual_seed(B + NH + Tq + Tk + d)
    q = torch.randn(B, NH, Tq, d, generator=g); k = torch.randn(B, NH, Tk, d, generator=g); v = torch.randn(B, NH, Tk, d, generator=g)
    Q = torch.zeros(B * Tq, NH * dp, dtype=torch.half); K = torch.zeros(B * Tk_pad, NH * dp, dtype=torch.half)
    Vt = torch.zeros(NH * dp, B * Tk_pad, dtype=torch.half)
    Q.view(B, Tq, NH, dp)[..., :d] = q.permute(0, 2, 1, 3).half()
    K.view(B, Tk_pad, NH, dp)[:, :Tk, :, :d] = k.permute(0, 2, 1, 3).half()
    Vt.view(NH, dp, B, Tk_pad)[:, :d, :, :Tk] = v.permute(1, 3, 0, 2).half()
    Q, K, Vt = Q.to(dev), K.to(dev), Vt.to(dev)
    out = torch.full((B, NH, Tq, d) if quirk else (B, Tq, NH, d), float("nan"), dtype=torch.half, device=dev)
    if quirk: osb, osh, ost = NH * Tq * d, Tq * d, d
    else: osb, osh, ost = Tq * NH * d, d, NH * d
    b200.tf_attention_set_tuning(bn)
    def call():
        b200.check(b200.tf_attention_f16(Q.data_ptr(), NH * dp, K.data_ptr(), NH * dp, Vt.data_ptr(), B * Tk_pad, out.data_ptr(),
                                         osb, osh, ost, B, NH, Tq, Tk, Tk_pad, d, dp, 1.0 / math.sqrt(d), S()), "tf_attention_f16")
    call(); torch.cuda.synchronize()
    qh, kh, vh = q.half().float().to(dev), k.half().float().to(dev), v.half().float().to(dev)
    ref = torch.softmax(qh @ kh.transpose(-1, -2) / math.sqrt(d), dim=-1) @ vh
    if not quirk: ref = ref.permute(0, 2, 1, 3)
    rec = dict(kind="attn", B=B, NH=NH, Tq=Tq, Tk=Tk, d=d, quirk=quirk, bn=bn, rel_err=relerr(out, ref), nan=bool(torch.isnan(out).any()))
    if time_it:
        rec["ms"] = timeit(call); rec["tflops"] = 4.0 * B * NH * Tq * Tk * d / rec["ms"] / 1e9
    b200.tf_attention_set_tuning(0)
    res.append(rec); print(rec, flush=True)

def groupnorm(NI, HW, C1, C2, silu=True, time_it=False):
    g = torch.Generator().manual_seed(NI + HW + C1 + C2)
    C = C1 + C2
    x1 = (torch.randn(NI, HW, C1, generator=g) * 1.5 + 0.7).half().to(dev)
    x2 = (torch.randn(NI, HW, C2, generator=g) * 0.5 - 0.3).half().to(dev) if C2 else None
    gamma = (1 + 0.1 * torch.randn(C, generator=g)).to(dev); beta = (0.1 * torch.randn(C, generator=g)).to(dev)
    out = torch.full((NI, HW, C), float("nan"), dtype=torch.half, device=dev)
    ws = torch.empty(2 * 32 * NI, dtype=torch.float32, device=dev)
    def call():
        b200.check(b200.tf_groupnorm_nhwc_f16(x1.data_ptr(), C1, C1, x2.data_ptr() if C2 else None, C2, C2, out.data_ptr(), C, NI, HW, 32,
                                              gamma.data_ptr(), beta.data_ptr(), 1e-5, 1 if silu else 0, ws.data_ptr(), S()), "tf_groupnorm_nhwc_f16")
    call(); torch.cuda.synchronize()
    x = torch.cat([x1, x2], dim=-1) if C2 else x1
    xr = x.float().permute(0, 2, 1).reshape(NI, C, HW, 1)
    ref = torch.nn.functional.group_norm(xr, 32, gamma, beta, 1e-5)
    if silu: ref = torch.nn.functional.silu(ref)
    ref = ref.reshape(NI, C, HW).permute(0, 2, 1)
    rec = dict(kind="groupnorm", NI=NI, HW=HW, C1=C1, C2=C2, silu=silu, rel_err=relerr(out, ref), nan=bool(torch.isnan(out).any()))
    if time_it:
        rec["ms"] = timeit(call); rec["GBps"] = 2 * 2.0 * NI * HW * C / rec["ms"] / 1e6
    res.append(rec); print(rec, flush=True)

def layernorm(B, T, C, il, time_it=False):
    g = torch.Generator().manual_seed(B + T + C)
    x = (torch.randn(B, T, C, generator=g) * 2 + 0.5).half().to(dev)
    gamma = (1 + 0.1 * torch.randn(C, generator=g)).to(dev); beta = (0.1 * torch.randn(C, generator=g)).to(dev)
    out = torch.full((B, T, C), float("nan"), dtype=torch.half, device=dev)
    def call():
        b200.check(b200.tf_layernorm_f16(x.data_ptr(), out.data_ptr(), B * T, C, gamma.data_ptr(), beta.data_ptr(), 1e-5, il, S()), "tf_layernorm_f16")
    call(); torch.cuda.synchronize()
    if il == 1:
        ref = torch.nn.functional.layer_norm(x.float(), (C,), gamma, beta, 1e-5)
    else:
        mem = x.float().reshape(T * B // il, C, il)
        ref = (torch.nn.functional.layer_norm(mem.permute(0, 2, 1), (C,), gamma, beta, 1e-5)).permute(0, 2, 1).reshape(B, T, C)
    rec = dict(kind="layernorm", B=B, T=T, C=C, il=il, rel_err=relerr(out, ref), nan=bool(torch.isnan(out).any()))
    if time_it:
        rec["ms"] = timeit(call); rec["GBps"] = 2 * 2.0 * B * T * C / rec["ms"] / 1e6
    res.append(rec); print(rec, flush=True)

def misc():
    # timestep embedding
    t = torch.tensor([981.0, 1.0], device=dev); idx = torch.tensor([0], dtype=torch.int32, device=dev)
    out = torch.empty(320, device=dev)
    b200.check(b200.tf_timestep_embedding_f32(t.data_ptr(), idx.data_ptr(), 320, 10000.0, out.data_ptr(), S()), "temb")
    half = 160
    fr = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float64) / half)
    a = 981.0 * fr
    ref = torch.cat([torch.cos(a), torch.sin(a)]).float().to(dev)
    res.append(dict(kind="temb", rel_err=relerr(out, ref), nan=False)); print(res[-1])
    # gemv
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1280, generator=g).to(dev); W = (torch.randn(2000, 1280, generator=g) / 36).half().to(dev)
    bb = torch.randn(2000, generator=g).to(dev); b2 = torch.randn(2000, generator=g).to(dev)
    o = torch.empty(2000, device=dev)
    b200.check(b200.tf_gemv_f16w(x.data_ptr(), W.data_ptr(), bb.data_ptr(), b2.data_ptr(), o.data_ptr(), 2000, 1280, 1, S()), "gemv")
    ref = torch.nn.functional.silu(x) @ W.float().t() + bb + b2
    res.append(dict(kind="gemv", rel_err=relerr(o, ref), nan=False)); print(res[-1])
    # conv_in
    x = torch.randn(2, 4, 64, 64, generator=g).to(dev); w = (torch.randn(320, 4, 3, 3, generator=g) / 6).to(dev); bi = torch.randn(320, generator=g).to(dev)
    o = torch.empty(2, 64, 64, 320, dtype=torch.half, device=dev)
    b200.check(b200.tf_conv3x3_smallcin_f32nchw(x.data_ptr(), w.data_ptr(), bi.data_ptr(), o.data_ptr(), 2, 4, 64, 64, 320, 320, S()), "conv_in")
    ref = torch.nn.functional.conv2d(x, w, bi, padding=1).permute(0, 2, 3, 1)
    res.append(dict(kind="conv_in", rel_err=relerr(o, ref), nan=False)); print(res[-1])
    # upsample
    x = torch.randn(2, 8, 8, 64, generator=g).half().to(dev); o = torch.empty(2, 16, 16, 64, dtype=torch.half, device=dev)
    b200.check(b200.tf_upsample_nearest2x_nhwc_f16(x.data_ptr(), 64, o.data_ptr(), 64, 2, 8, 8, 64, S()), "ups")
    ref = x.repeat_interleave(2, 1).repeat_interleave(2, 2)
    res.append(dict(kind="upsample", rel_err=relerr(o, ref), nan=False)); print(res[-1])
    # layout
    x = torch.randn(2, 20, 7, 9, generator=g).to(dev); o = torch.empty(2, 63, 24, dtype=torch.half, device=dev).fill_(0)
    b200.check(b200.tf_nchw_to_nhwc_f16(x.data_ptr(), 1, o.data_ptr(), 2, 20, 63, 24, S()), "n2h")
    ref = x.reshape(2, 20, 63).permute(0, 2, 1)
    res.append(dict(kind="nchw2nhwc", rel_err=relerr(o[..., :20], ref), nan=False)); print(res[-1])
    back = torch.empty(2, 20, 63, device=dev)
    b200.check(b200.tf_nhwc_to_nchw(o.data_ptr(), 24, back.data_ptr(), 1, 2, 20, 63, S()), "h2n")
    res.append(dict(kind="nhwc2nchw", rel_err=relerr(back, x.reshape(2, 20, 63).half()), nan=False)); print(res[-1])
    # pad tokens
    c = torch.randn(2, 77, 768, generator=g).to(dev); o = torch.empty(2, 80, 768, dtype=torch.half, device=dev)
    b200.check(b200.tf_pad_tokens_f32_to_f16(c.data_ptr(), o.data_ptr(), 2, 77, 80, 768, S()), "pad")
    ok = relerr(o[:, :77], c) ; z = o[:, 77:].abs().max().item()
    res.append(dict(kind="pad_tokens", rel_err=ok + z, nan=False)); print(res[-1])
    # cfg + ddim
    eps = torch.randn(2, 4096, 16, generator=g).to(dev); lat = torch.randn(1, 4, 4096, generator=g).to(dev)
    at = torch.tensor([0.5, 0.3], device=dev); ap = torch.tensor([0.7, 0.5], device=dev); idx = torch.tensor([1], dtype=torch.int32, device=dev)
    o = torch.empty_like(lat); e = torch.empty_like(lat)
    b200.check(b200.tf_cfg_ddim_step_f32(eps.data_ptr(), 16, lat.data_ptr(), o.data_ptr(), e.data_ptr(), at.data_ptr(), ap.data_ptr(), idx.data_ptr(), 7.5, 1, 4, 4096, S()), "cfg")
    u = eps[0, :, :4].t().reshape(1, 4, 4096); cc = eps[1, :, :4].t().reshape(1, 4, 4096)
    et = u + 7.5 * (cc - u)
    px0 = (lat - math.sqrt(1 - 0.3) * et) / math.sqrt(0.3)
    ref = math.sqrt(0.5) * px0 + math.sqrt(1 - 0.5) * et
    res.append(dict(kind="cfg_ddim", rel_err=relerr(o, ref) + relerr(e, et), nan=False)); print(res[-1])

try:
    misc()
    groupnorm(2, 4096, 320, 0, time_it=True)
    groupnorm(2, 4096, 640, 320, time_it=True)
    groupnorm(2, 64, 1280, 1280)
    groupnorm(1, 100, 64, 0, silu=False)
    layernorm(2, 4096, 320, 1, time_it=True)
    layernorm(2, 4096, 320, 2, time_it=True)
    layernorm(2, 256, 1280, 2)
    layernorm(2, 256, 1280, 1)
    layernorm(4, 64, 640, 4)
    attn(1, 1, 128, 64, 64)
    attn(1, 1, 128, 128, 64)
    attn(1, 2, 128, 256, 48)
    attn(2, 8, 256, 256, 160)
    attn(2, 8, 64, 64, 160)
    attn(2, 8, 1024, 1024, 80, time_it=True)
    attn(2, 8, 4096, 4096, 40, time_it=True)
    attn(2, 8, 4096, 4096, 40, bn=128, time_it=True)
    attn(2, 8, 4096, 77, 40, time_it=True)
    attn(2, 8, 1024, 77, 80, quirk=False)
    attn(2, 8, 200, 77, 160)
finally:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/dev_ops.json", "w"), indent=1)
bad = [r for r in res if r["rel_err"] > 5e-3 or r["nan"]]
print("FAILED CASES:", len(bad))
for r in bad: print("  ", r)
