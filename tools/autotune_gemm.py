"""GPU box: measure (BN, split-K, CTAs per tile) for every GEMM / conv shape of the UNet step (and the VAE decode) and
write the winners to gpurun_out/gemm_tuning.json (copy to tinyfusers_b200/native/b200/gemm_tuning.json to ship it).

    python tools/autotune_gemm.py [--configs 1x64,8x64,4x96] [--vae] [--min-gain 0.03]

Every candidate is timed as a captured chain of 16 dependent launches that rotate through enough copies of the weight
matrix to exceed the L2 (in the real step every layer's weights arrive cold from HBM; activations are L2-warm), CUDA
events around graph replays. A candidate only counts if the library honoured it (tf_gemm_last_choice) and its output
matches the default configuration's."""
import argparse, contextlib, io, json, math, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinyfusers_b200.native.b200.ops import b200

ap = argparse.ArgumentParser()
ap.add_argument("--configs", default="1x64")
ap.add_argument("--vae", action="store_true")
ap.add_argument("--min-gain", type=float, default=0.03)
ap.add_argument("--out", default="gpurun_out/gemm_tuning.json")
ap.add_argument("--only-gn", action="store_true", help="only shapes whose launch also emits GroupNorm statistics")
args = ap.parse_args()

os.environ["TINYFUSERS_B200_TUNING"] = "0"     # measure against the built-in model
dev = torch.device("cuda:0"); b200.init(0)
S = lambda: torch.cuda.current_stream().cuda_stream
ws = torch.empty(192 << 20, dtype=torch.uint8, device=dev)
import ctypes
N_CHAIN = 16


def record_shapes():
    from oracle import ref_ops as R
    from tinyfusers_b200.storage.state import update_state
    from tinyfusers_b200.variants.sd import StableDiffusion
    keys = {}
    m = StableDiffusion()
    with contextlib.redirect_stdout(io.StringIO()):
        update_state(m, R.make_unet_state_dict())
    for cfg in args.configs.split(","):
        B, HW = (int(v) for v in cfg.split("x"))
        lat, unc, ctx = R.make_inputs(B, HW)
        s = m._sampler(lat.shape, 77)
        s.load(unc.cuda(), ctx.cuda(), lat.cuda()); s.set_scalars(501, 0.5, 0.6, 7.5)
        s.enqueue_step(); torch.cuda.synchronize()
        s.unet_engine.ctx.prof = []
        s.enqueue_step(); torch.cuda.synchronize()
        for k, _a, _b, _fn in s.unet_engine.ctx.prof:
            if k[0] in ("gemm", "conv3x3"): keys[k] = keys.get(k, 0) + 1
        s.unet_engine.ctx.prof = None
        m._samplers.clear(); del s; torch.cuda.empty_cache()
    if args.vae:
        with contextlib.redirect_stdout(io.StringIO()):
            update_state(m.first_stage_model, R.make_vae_decoder_state_dict(), "first_stage_model")
        z = torch.randn(1, 4, 64, 64, device=dev)
        m.first_stage_model.decoder(z)
        eng = m.first_stage_model.decoder._engine(tuple(z.shape))
        eng.ctx.prof = []
        eng.decode_nhwc_f32(z); torch.cuda.synchronize()
        for k, _a, _b, _fn in eng.ctx.prof:
            if k[0] in ("gemm", "conv3x3"): keys[k] = keys.get(k, 0) + 1
        eng.ctx.prof = None
    del m; torch.cuda.empty_cache()
    return keys


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


def tune(key, count):
    fns, out, tkey, m_tiles, k_blocks, bn_mult, allow_split, keep = make_case(key)
    N = tkey[2]
    b200.tf_gemm_set_tuning(0, 0); b200.tf_gemm_set_ctas(0)
    fns[0](); torch.cuda.synchronize()
    default = last_choice()
    ref = out.float().clone()
    base_us = chain_us(fns)
    best, best_us, tried = default, base_us, 0
    for bn in range(32, 257, 32):
        unit_ok = bn % bn_mult == 0        # split-K launches leave the GroupNorm statistics to the fold: any multiple of 32
        if not unit_ok and not allow_split: continue
        n_tiles = (N + bn - 1) // bn
        if N / (n_tiles * bn) < 0.65 and bn > bn_mult: continue   # the narrowest tile is always a candidate (N = 8 conv_out)
        for ctas in (1, 2):
            if ctas == 2 and m_tiles < 2: continue
            mt = 2 * ((m_tiles + 1) // 2) if ctas == 2 else m_tiles
            for sp in (1, 2, 3, 4, 5, 6, 8, 10, 12, 14, 16):
                if sp == 1 and not unit_ok: continue
                if sp > 1 and (not allow_split or k_blocks // sp < 2): break
                tiles = mt * n_tiles * sp
                if tiles > 2.2 * 148 and sp > 1: break
                if sp > 1 and tiles < 32: continue
                if (bn, sp, ctas) == default: continue
                b200.tf_gemm_set_tuning(bn, sp); b200.tf_gemm_set_ctas(ctas)
                try:
                    fns[0](); torch.cuda.synchronize()
                except RuntimeError:
                    continue
                if last_choice() != (bn, sp, ctas): continue
                err = float((out.float() - ref).abs().max() / ref.abs().max().clamp_min(1e-9))
                if not err < 5e-3:
                    print(f"   !! {key} {(bn, sp, ctas)} differs from default by {err:.2e}; skipped", flush=True); continue
                us = chain_us(fns); tried += 1
                if us < best_us: best, best_us = (bn, sp, ctas), us
    b200.tf_gemm_set_tuning(0, 0); b200.tf_gemm_set_ctas(0)
    gain = 1 - best_us / base_us
    print(f"{str(key):66s} n={count:3d} default {default} {base_us:7.2f} us -> best {best} {best_us:7.2f} us ({100 * gain:4.1f}%) [{tried} tried]", flush=True)
    del keep, fns; torch.cuda.empty_cache()
    return tkey, default, base_us, best, best_us, count


t0 = time.time()
keys = record_shapes()
print(f"{len(keys)} unique GEMM / conv shapes", flush=True)
entries, saved = [], 0.0
for key, count in sorted(keys.items(), key=lambda kv: str(kv[0])):
    if args.only_gn and not (key[8] if key[0] == "conv3x3" else key[6]):
        continue
    tkey, default, base_us, best, best_us, count = tune(key, count)
    if best != default and best_us < (1 - args.min_gain) * base_us:
        entries.append(list(tkey) + list(best) + [round(base_us, 2), round(best_us, 2), count, str(key)])
        saved += count * (base_us - best_us)
os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
with open(args.out, "w") as fh:
    json.dump({"how": "tools/autotune_gemm.py on one B200: per shape, chain of 16 launches rotating cold weight copies, CUDA events; "
                      "entry = [is_conv, M, N, K, class, BN, splits, ctas, default_us, tuned_us, launches_per_step, shape]",
               "configs": args.configs, "vae": args.vae, "entries": entries}, fh, indent=0)
print(f"{len(entries)} tuned entries, {saved:.0f} us saved per recorded step set, {time.time() - t0:.0f} s -> {args.out}")
