"""The measured tile-choice table (tf_gemm_tuning_add / native/b200/gemm_tuning.json): an entry steers the launch,
changes nothing but the summation order, and an entry that does not fit the call is ignored."""
import ctypes
import json
import os

import pytest
import torch

from conftest import rel_err


def _choice(b200):
    a, b, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    b200.tf_gemm_last_choice(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
    return a.value, b.value, c.value


def test_shipped_table_is_well_formed():
    from tinyfusers_b200.native.b200.ops import b200
    path = os.path.join(os.path.dirname(b200.path), "gemm_tuning.json")
    entries = json.load(open(path))["entries"]
    assert len(entries) > 0
    for e in entries:
        is_conv, M, N, K, klass, bn, splits, ctas = (int(v) for v in e[:8])
        assert is_conv in (0, 1) and M > 0 and N > 0 and K > 0 and 0 <= klass < 32
        assert 32 <= bn <= 256 and bn % 32 == 0 and 1 <= splits <= 16 and ctas in (1, 2)
        assert float(e[9]) < float(e[8])      # the tuned time beat the model's choice when it was measured


@pytest.mark.gpu
def test_entry_steers_launch_and_keeps_result():
    from tinyfusers_b200.native.b200.ops import b200
    b200.init(0)
    S = torch.cuda.current_stream().cuda_stream
    M, N, K = 768, 640, 1280
    g = torch.Generator().manual_seed(3)
    A = torch.randn(M, K, generator=g).cuda().half()
    W = (torch.randn(N, K, generator=g) / 36).cuda().half()
    bias = torch.randn(N, generator=g).cuda()
    ws = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    out = torch.empty(M, N, dtype=torch.half, device="cuda")
    run = lambda: b200.check(b200.tf_gemm_f16(A.data_ptr(), K, W.data_ptr(), K, out.data_ptr(), N, M, N, K, bias.data_ptr(),
                                              None, 0, 0, ws.data_ptr(), ws.numel(), S), "gemm")
    try:
        b200.tf_gemm_tuning_clear()
        run()
        model_choice, ref = _choice(b200), out.clone()
        forced = (64, 4, 1) if model_choice != (64, 4, 1) else (128, 2, 1)
        b200.check(b200.tf_gemm_tuning_add(0, M, N, K, 0, *forced), "tf_gemm_tuning_add")
        run()
        assert _choice(b200) == forced
        assert rel_err(out, ref) < 2e-3
        assert rel_err(out, A.float() @ W.float().t() + bias) < 5e-3
        # an entry that cannot be honoured (split-K partials larger than the workspace) falls back to the model
        small = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
        b200.check(b200.tf_gemm_tuning_add(0, M, N, K, 0, 64, 16, 1), "tf_gemm_tuning_add")
        b200.check(b200.tf_gemm_f16(A.data_ptr(), K, W.data_ptr(), K, out.data_ptr(), N, M, N, K, bias.data_ptr(), None, 0, 0,
                                    small.data_ptr(), small.numel(), S), "gemm")
        sp = _choice(b200)[1]
        assert sp != 16 and (sp == 1 or sp * M * N * 4 <= small.numel())
        assert rel_err(out, ref) < 2e-3
    finally:
        b200.load_tuning()
