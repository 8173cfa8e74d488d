"""Static activation helpers with the reference's names (reference: tinyfusers/storage/tensor.py:64-86).

The reference's `Tensor` is also a raw cudaMalloc container (tensor.py:9-62) that the UNet path never
uses (out of scope); only the static API the model files call is mirrored.
Inside the UNet these activations are fused into GroupNorm / GEMM epilogues; called on their own they
run the stand-alone elementwise kernel."""
import functools

import torch

from .. import fp32
from ..native.b200.ops import b200
from ..runtime import require_cuda, stream_ptr

_OPS = {"sigmoid": 0, "silu": 1, "gelu": 2, "quick_gelu": 3}


def _unary(x, op):
    require_cuda(x, "x")
    if fp32.enabled():
        return fp32.unary(x, _OPS[op])
    if x.dtype not in (torch.float32, torch.float16):
        x = x.to(torch.float32)
    x = x.contiguous()
    out = torch.empty_like(x)
    st = b200.tf_unary(x.data_ptr(), out.data_ptr(), x.numel(), _OPS[op], 1 if x.dtype == torch.float32 else 0,
                       stream_ptr())
    b200.check(st, "tf_unary")
    return out


class Tensor:
    @staticmethod
    def sigmoid(x):
        return _unary(x, "sigmoid")

    @staticmethod
    def silu(x):
        return _unary(x, "silu")

    @staticmethod
    def swish(x):
        return _unary(x, "silu")

    @staticmethod
    def gelu(x):
        return _unary(x, "gelu")

    @staticmethod
    def quick_gelu(x):
        return _unary(x, "quick_gelu")

    @staticmethod
    def sequential(iterable, init):
        return functools.reduce(lambda x, f: f(x), iterable, init)
