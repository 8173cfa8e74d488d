"""The reference's OWN softmax kernel (tinyfusers/native/cuda/softmax.cu:24-112, compiled from /root/reference by
oracle/build_ref.py into oracle/_ref/libref_softmax.so, launched with the geometry of attention/sdpa.py:59-73) run on the
GPU, as the yardstick for
  * the oracle's restatement of it (oracle.softmax_rows)                      <= 1e-5
  * the fp32 parity kernel tf_softmax_rows_f32                                 <= 1e-5
  * SDPA assembled the way the reference does (fp32 matmul, this kernel, fp32 matmul; attention/sdpa.py:62-76)
    against the fp32 parity SDPA (<= 1e-5) and the fused fp16 tcgen05 attention kernel (<= 1e-2).
The reference's Python names the kernel `softmax_forward_kernel` (sdpa.py:15), a symbol softmax.cu does not define; the
kernel it holds is `softmax_kernel`."""
import ctypes
import math
import os

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref_softmax.so")


@pytest.fixture(scope="module")
def ref_softmax():
    if not os.path.exists(LIB):
        pytest.skip("oracle/_ref/libref_softmax.so not built (python oracle/build_ref.py where /root/reference exists)")
    lib = ctypes.CDLL(LIB)
    lib.ref_softmax_forward.restype = ctypes.c_int
    lib.ref_softmax_forward.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]

    def run(x):
        x = x.contiguous()
        out = torch.empty_like(x)
        st = lib.ref_softmax_forward(out.data_ptr(), x.data_ptr(), x.numel() // x.shape[-1], x.shape[-1],
                                     torch.cuda.current_stream().cuda_stream)
        assert st == 0, f"reference softmax_kernel launch failed with status {st}"
        torch.cuda.synchronize()
        return out
    return run


@pytest.mark.parametrize("N,C", [(2048, 256), (4096, 77), (512, 4096), (37, 1000), (64, 9216)])
def test_reference_softmax_kernel_vs_oracle_and_fp32_kernel(ref_softmax, oracle, N, C):
    from tinyfusers_b200.native.b200.ops import b200
    b200.init(0)
    x = torch.randn(N, C, generator=torch.Generator().manual_seed(N + C)) * 3
    ref = ref_softmax(x.cuda())
    assert rel_err(oracle.softmax_rows(x), ref) < 1e-5
    assert rel_err(oracle.softmax_rows(x.double()), ref) < 1e-5
    mine = x.cuda().clone()
    b200.check(b200.tf_softmax_rows_f32(mine.data_ptr(), N, C, 0, torch.cuda.current_stream().cuda_stream),
               "tf_softmax_rows_f32")
    assert rel_err(mine, ref) < 1e-5
    assert abs(ref.sum(dim=-1) - 1).max().item() < 1e-5


@pytest.mark.parametrize("B,NH,Tq,Tk,d", [(2, 8, 1024, 1024, 80), (2, 8, 4096, 77, 40), (1, 8, 4096, 4096, 40)])
def test_sdpa_through_the_reference_kernel(ref_softmax, B, NH, Tq, Tk, d):
    import tinyfusers_b200
    from tinyfusers_b200.attention.sdpa import scaled_dot_product_attention
    g = torch.Generator().manual_seed(Tq + Tk)
    q, k, v = (torch.randn(B, NH, t, d, generator=g).cuda() for t in (Tq, Tk, Tk))
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False          # the reference's matmuls are true fp32 (cuBLAS SGEMM)
    try:
        scale = torch.tensor(1.0 / math.sqrt(d), dtype=torch.float32).item()
        preatt = scale * torch.matmul(q, k.transpose(-1, -2))                  # sdpa.py:66
        att = ref_softmax(preatt.reshape(B * NH * Tq, Tk)).reshape(B, NH, Tq, Tk)   # sdpa.py:72-74
        ref = torch.matmul(att, v)                                              # sdpa.py:76
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32
    assert rel_err(scaled_dot_product_attention(q, k, v), ref) < 1e-2           # fused tcgen05 kernel, fp16 operands
    tinyfusers_b200.set_precision("fp32")
    try:
        assert rel_err(scaled_dot_product_attention(q, k, v), ref) < 1e-5
    finally:
        tinyfusers_b200.set_precision("fp16")
