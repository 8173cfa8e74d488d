"""Linear with the reference's signature (reference: tinyfusers/ff/linear.py:112-121).

`cp.dot(x, W.T) + b` (cuBLAS SGEMM through CuPy, separate bias pass) becomes one tf_gemm_f16 call:
tcgen05 GEMM, fp16 operands, fp32 accumulate, bias fused in the epilogue."""
import torch

from .. import fp32, packing
from ..native.b200.ops import b200
from ..runtime import F16, F32, as_f16, require_cuda, standalone_context
from ..storage.state import _default_device


class Linear:
    def __init__(self, in_features, out_features, bias=True):
        dev = _default_device()
        # reference init: all ones (linear.py:114-115); overwritten by update_state in practice
        self.weight = torch.ones((out_features, in_features), dtype=F32, device=dev)
        self.bias = torch.ones((out_features,), dtype=F32, device=dev) if bias else None

    # packed fp16 weight (K padded to 8) + fp32 bias, rebuilt only when .weight/.bias change
    def _packed(self):
        return packing.cached(self, "lin", (self.weight, self.bias),
                              lambda: (packing.pad_rows(packing.linear_weight(self.weight), 8), packing.f32(self.bias)))

    def __call__(self, x):
        require_cuda(x, "x")
        if fp32.enabled():      # parity mode: plain fp32 kernels (csrc/tf_fp32.cu)
            return fp32.linear(x, self.weight, self.bias)
        ctx = standalone_context()
        w, b = self._packed()
        out_f, in_f = self.weight.shape
        Kp = w.shape[1]
        x2 = x.reshape(-1, in_f)
        M = x2.shape[0]
        if Kp == in_f:
            a = as_f16(x2)
        else:
            a = torch.zeros((M, Kp), dtype=F16, device=x.device)
            a[:, :in_f] = x2
        Np = w.shape[0]
        out = torch.empty((M, Np), dtype=F32, device=x.device)
        if b is not None and Np != out_f:
            b = torch.nn.functional.pad(b, (0, Np - out_f))
        ctx.gemm(a.data_ptr(), Kp, M, Kp, w.data_ptr(), Np, out.data_ptr(), Np,
                 bias=b.data_ptr() if b is not None else None, flags=b200.TF_EPI_OUT_F32)
        return out[:, :out_f].reshape(*x.shape[:-1], out_f)
