"""tf_attention_v_f16: attention with V in its natural (keys, head dim padded to 64) layout - the MN-major B operand of
O += P V - against torch on the same random inputs, at the UNet's head sizes (40 / 80 / 160), CLIP's (64, causal) and the
C4 / C5 grid sizes that take the 3-CTA-per-SM variant. Tolerance: max|a-b| / max|b| <= 3e-3 (fp16 P and V)."""
import math

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _run(B, NH, Tq, Tk, d, causal=False, seed=0, pad64=False, ones_col=False):
    from tinyfusers_b200.native.b200.ops import b200
    from tinyfusers_b200.attention.attention import _pad64
    b200.init(0)
    dp = (d + 15) // 16 * 16
    dvp = (d + 63) // 64 * 64 if pad64 else _pad64(d)     # 128-byte-atom tiles, or the layout the model uses
    Tkp = (Tk + 7) // 8 * 8
    g = torch.Generator().manual_seed(seed + Tq + d)
    q = torch.randn(B, Tq, NH, d, generator=g)
    k = torch.randn(B, Tk, NH, d, generator=g)
    v = torch.randn(B, Tk, NH, d, generator=g)
    Q = torch.zeros(B, Tq, NH, dp, dtype=torch.half, device="cuda"); Q[..., :d] = q.cuda()
    K = torch.zeros(B, Tkp, NH, dp, dtype=torch.half, device="cuda"); K[:, :Tk, :, :d] = k.cuda()
    # pad key rows / pad head columns of V must be finite (rows are masked, columns must be ZERO: they host the row sums)
    V = torch.zeros(B, Tkp, NH, dvp, dtype=torch.half, device="cuda"); V[:, :Tk, :, :d] = v.cuda()
    if ones_col:      # TF_ATTN_V_ONES_COLUMN: column d of every head is 1.0 in every key row, the row sums come out of P V
        assert dvp > d
        V[..., d] = 1.0
    out = torch.zeros(B, Tq, NH, d, dtype=torch.half, device="cuda")
    st = b200.tf_attention_v_f16(Q.data_ptr(), NH * dp, K.data_ptr(), NH * dp, V.data_ptr(), NH * dvp, out.data_ptr(),
                                 Tq * NH * d, d, NH * d, B, NH, Tq, Tk, Tkp, d, dp, dvp, 1.0 / math.sqrt(d),
                                 (1 if causal else 0) | (2 if ones_col else 0),
                                 torch.cuda.current_stream().cuda_stream)
    b200.check(st, "tf_attention_v_f16")
    heads = lambda t: t.cuda().half().float().permute(0, 2, 1, 3)
    ref = torch.nn.functional.scaled_dot_product_attention(heads(q), heads(k), heads(v), is_causal=causal)
    return out.float().permute(0, 2, 1, 3), ref


@pytest.mark.parametrize("B,NH,Tq,Tk,d", [(2, 8, 4096, 4096, 40), (2, 8, 1024, 1024, 80), (2, 8, 256, 256, 160),
                                          (2, 8, 64, 64, 160), (2, 8, 4096, 77, 40), (2, 8, 256, 77, 160),
                                          (1, 8, 576, 576, 80), (1, 3, 200, 333, 64)])
def test_natural_v_matches_torch(B, NH, Tq, Tk, d):
    for pad64 in (False, True):
        out, ref = _run(B, NH, Tq, Tk, d, pad64=pad64)
        assert rel_err(out, ref) < 3e-3, pad64


def test_natural_v_three_ctas_per_sm_variant():
    for pad64 in (False, True):               # 2048 CTAs -> 64-key blocks, 3 CTAs / SM (pad64: L overlaid on O's pad columns)
        out, ref = _run(8, 8, 4096, 4096, 40, pad64=pad64)
        assert rel_err(out, ref) < 3e-3, pad64


@pytest.mark.parametrize("T,NH,d", [(77, 12, 64), (200, 4, 64), (640, 2, 40)])
def test_natural_v_causal(T, NH, d):
    out, ref = _run(1, NH, T, T, d, causal=True)
    assert rel_err(out, ref) < 3e-3


@pytest.mark.parametrize("emu", [0, 2, 4])
@pytest.mark.parametrize("B,NH,T,d,causal", [(2, 8, 4096, 40, False), (1, 4, 1024, 80, False), (1, 2, 512, 64, True),
                                             (1, 1, 256, 40, False), (1, 3, 768, 96, False)])
def test_two_tile_pingpong_kernel(B, NH, T, d, causal, emu):
    """tf_attention2_kernel forced on (version 2) in every variant: exponentials all on MUFU / 2 / 4 of 8 on the FMA pipe
    (degree-3 polynomial, max relative error 7.6e-5 per value), row sums by
    the ones-tile MMA or out of a ones column of V. Same tolerance as the one-tile kernel."""
    from tinyfusers_b200.native.b200.ops import b200
    b200.init(0)
    try:
        b200.check(b200.tf_attention_set_variant(2, emu), "variant")
        for ones in (False, True):
            if ones and d % 64 == 0:
                continue          # no pad column to host the ones
            out, ref = _run(B, NH, T, T, d, causal=causal, pad64=ones, ones_col=ones)
            assert rel_err(out, ref) < 3e-3, (emu, ones)
    finally:
        b200.tf_attention_set_variant(0, -1)


def test_two_tile_kernel_peaked_rows_rescale():
    """Scores whose row maximum keeps growing along the keys (sorted-key ramp x large scale): exercises the lazy rescale of O
    against P V MMAs still in flight - the hazard the per-block pv_done wait closes."""
    from tinyfusers_b200.native.b200.ops import b200
    b200.init(0)
    B, NH, T, d = 1, 2, 2048, 40
    dp, dvp = 48, 64
    g = torch.Generator().manual_seed(5)
    q = torch.randn(B, T, NH, d, generator=g).abs() * 2.0
    k = torch.randn(B, T, NH, d, generator=g).abs() * torch.linspace(0.05, 3.0, T).reshape(1, T, 1, 1)   # later keys score higher
    v = torch.randn(B, T, NH, d, generator=g)
    Q = torch.zeros(B, T, NH, dp, dtype=torch.half, device="cuda"); Q[..., :d] = q.cuda()
    K = torch.zeros(B, T, NH, dp, dtype=torch.half, device="cuda"); K[..., :d] = k.cuda()
    V = torch.zeros(B, T, NH, dvp, dtype=torch.half, device="cuda"); V[..., :d] = v.cuda()
    out = torch.zeros(B, T, NH, d, dtype=torch.half, device="cuda")
    heads = lambda t: t.cuda().half().float().permute(0, 2, 1, 3)
    ref = torch.nn.functional.scaled_dot_product_attention(heads(q), heads(k), heads(v)).permute(0, 2, 1, 3)
    try:
        for ver in (1, 2, 3):
            b200.check(b200.tf_attention_set_variant(ver, 2), "variant")
            out.zero_()
            st = b200.tf_attention_v_f16(Q.data_ptr(), NH * dp, K.data_ptr(), NH * dp, V.data_ptr(), NH * dvp, out.data_ptr(),
                                         T * NH * d, d, NH * d, B, NH, T, T, T, d, dp, dvp, 1.0 / math.sqrt(d), 0,
                                         torch.cuda.current_stream().cuda_stream)
            b200.check(st, "tf_attention_v_f16")
            assert rel_err(out.float(), ref) < 3e-3, ver
    finally:
        b200.tf_attention_set_variant(0, -1)


@pytest.mark.parametrize("emu", [0, 2, 4])
@pytest.mark.parametrize("B,NH,T,Tk,d,causal", [(2, 8, 4096, 4096, 40, False), (1, 2, 512, 512, 40, True), (1, 3, 384, 200, 40, False),
                                                (1, 1, 256, 256, 48, False), (2, 2, 1024, 1024, 32, False), (1, 2, 256, 136, 40, False)])
def test_split_row_kernel(B, NH, T, Tk, d, causal, emu):
    """tf_attention3_kernel forced on (version 3): every row split over two threads with their own reference max and output
    accumulator, combined in the epilogue - also with a causal mask (half rows that see no key in a block, or in no block at
    all) and a ragged last key block (Tk % 64 != 0, incl. a second half that is masked entirely). Row sums by the ones-tile
    MMA into V's pad columns or out of a ones column of V. Same tolerance as the other kernels."""
    from tinyfusers_b200.native.b200.ops import b200
    b200.init(0)
    try:
        b200.check(b200.tf_attention_set_variant(3, emu), "variant")
        for ones in (False, True):
            out, ref = _run(B, NH, T, Tk, d, causal=causal, pad64=True, ones_col=ones)
            assert rel_err(out, ref) < 3e-3, (emu, ones)
    finally:
        b200.tf_attention_set_variant(0, -1)
