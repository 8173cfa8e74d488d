"""Dev (GPU box): split-K folded inside a thread-block cluster (DSMEM) vs the fp32 workspace + fold kernel, on the split-K
shapes of one UNet step: output and GroupNorm statistics against the workspace path, in-graph time per call (cold weights)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _gemm_cases import b200, chain_us, last_choice, make_case

KEYS = [
    (("conv3x3", 2, 8, 8, 1280, 1280, 1, False, 10), 4),
    (("conv3x3", 2, 8, 8, 1280, 1280, 1, True, 10), 4),
    (("conv3x3", 2, 8, 8, 2560, 1280, 1, False, 10), 3),
    (("conv3x3", 2, 8, 8, 1280, 1280, 1, False, 10, 2560), 3),
    (("conv3x3", 2, 16, 16, 1280, 1280, 1, False, 10), 2),
    (("conv3x3", 2, 16, 16, 1280, 1280, 1, True, 10), 1),
    (("conv3x3", 2, 16, 16, 2560, 1280, 1, False, 10), 2),
    (("conv3x3", 2, 16, 16, 1920, 1280, 1, False, 10), 1),
    (("conv3x3", 2, 16, 16, 640, 1280, 1, False, 10), 1),
    (("conv3x3", 2, 16, 16, 1280, 1280, 1, False, 10, 2560), 2),
    (("conv3x3", 2, 16, 16, 1280, 1280, 2, False, 10), 1),
    (("conv3x3", 2, 32, 32, 640, 640, 2, False, 10), 1),
    (("conv3x3", 2, 64, 64, 320, 320, 2, False, 10), 1),
    (("conv3x3", 2, 32, 32, 1920, 640, 1, False, 10), 1),
    (("gemm", 128, 1280, 1280, 0, True, 0, 0), 2),
    (("gemm", 128, 1280, 1280, 0, False, 0, 0), 2),
    (("gemm", 128, 1280, 5120, 0, True, 0, 0), 1),
    (("gemm", 128, 1280, 2560, 0, False, 0, 0), 3),
    (("gemm", 512, 1280, 5120, 0, True, 0, 0), 5),
    (("gemm", 512, 1280, 2560, 0, False, 0, 0), 2),
    (("gemm", 128, 1280, 1280, 0, True, 10, 64), 1),
]
sweep = "--sweep" in sys.argv
saved = 0.0
for key, count in KEYS:
    fns, out, tkey, m_tiles, k_blocks, bn_mult, allow_split, keep = make_case(key)
    st = keep[5] if key[0] == "conv3x3" else keep[5]
    b200.tf_gemm_set_tuning(0, 0); b200.tf_gemm_set_ctas(0)
    b200.tf_gemm_set_cluster_splitk(0)
    out.zero_(); fns[0](); torch.cuda.synchronize()
    off_choice = last_choice(); ref = out.float().clone(); ref_st = st.clone() if st is not None else None
    us_off = chain_us(fns)
    b200.tf_gemm_set_cluster_splitk(1)
    out.zero_()
    if st is not None: st.zero_()
    fns[0](); torch.cuda.synchronize()
    on_choice = last_choice()
    err = float((out.float() - ref).abs().max() / ref.abs().max())
    serr = float((st - ref_st).abs().max() / ref_st.abs().max()) if st is not None else 0.0
    us_on = chain_us(fns)
    best, best_us = on_choice, us_on
    if sweep and off_choice[1] > 1:
        unit = key[8] if key[0] == "conv3x3" else key[6]
        unit = unit or 1
        for bn in range(16, 257, 16):
            for sp in range(2, 9):
                if bn % (2 * sp) or (bn // sp) % unit or bn % unit or (bn, sp, 1) == on_choice: continue
                n_tiles = -(-tkey[2] // bn)
                if tkey[2] / (n_tiles * bn) < 0.8 or m_tiles * n_tiles * sp > 300 or m_tiles * n_tiles * sp < 40: continue
                b200.tf_gemm_set_tuning(bn, sp)
                try:
                    fns[0](); torch.cuda.synchronize()
                except RuntimeError:
                    continue
                if last_choice() != (bn, sp, 1): continue
                e2 = float((out.float() - ref).abs().max() / ref.abs().max())
                if not e2 < 5e-3:
                    print(f"   !! {key} {(bn, sp)} err {e2:.2e}"); continue
                us = chain_us(fns)
                if us < best_us: best, best_us = (bn, sp, 1), us
        b200.tf_gemm_set_tuning(0, 0)
    saved += count * (us_off - best_us)
    print(f"{str(key):62s} n={count} workspace {off_choice} {us_off:6.2f} us | cluster {on_choice} {us_on:6.2f} us err {err:.1e} stats {serr:.1e}"
          + (f" | best {best} {best_us:6.2f}" if sweep else ""), flush=True)
    del keep, fns; torch.cuda.empty_cache()
b200.tf_gemm_set_cluster_splitk(-1)
print(f"saved per step (these shapes): {saved:.0f} us")
