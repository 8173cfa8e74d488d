"""LayerNorm with the reference's signatures (reference: tinyfusers/ff/layer_norm.py:8-49).

Canonical LayerNorm over the last dimension: that is what real cuDNN computes for the reference's graph at
batch 1, the only case it executes — at batch > 1 it rejects the reference's stride declaration
[B*T*C, 1, B*C, B] (layer_norm.py:10) with CUDNN_STATUS_NOT_SUPPORTED (oracle/cudnn_probe.py). The literal
reading of those strides (view the buffer as (T, C, B), normalise over C) is available through
`tinyfusers_b200.set_layernorm_strided(True)` and runs the same kernel with interleave = B."""
from typing import Tuple, Union

import numpy as np
import torch

from .. import fp32, get_layernorm_strided, packing
from ..runtime import F16, F32, as_f16, as_f32, require_cuda, standalone_context
from ..storage.state import _default_device


def layer_norm(x_gpu, scale_gpu, bias_gpu, epsilon_cpu):
    """x_gpu: (1, B, T, C) as in the reference; scale/bias: (1,1,1,C); epsilon_cpu: np.full((1,1,1,1))."""
    require_cuda(x_gpu, "x_gpu")
    if fp32.enabled():
        return fp32.layer_norm(x_gpu, scale_gpu.reshape(-1), bias_gpu.reshape(-1), float(np.asarray(epsilon_cpu).reshape(-1)[0]))
    ctx = standalone_context()
    _, B, T, C = x_gpu.shape
    xh = as_f16(x_gpu)
    out = torch.empty_like(xh)
    g = scale_gpu.reshape(-1).to(F32).contiguous()
    b = bias_gpu.reshape(-1).to(F32).contiguous()
    il = B if get_layernorm_strided() else 1
    ctx.layernorm(xh.data_ptr(), out.data_ptr(), B * T, C, g.data_ptr(), b.data_ptr(), float(np.asarray(epsilon_cpu).reshape(-1)[0]), il)
    return as_f32(out)


class LayerNorm:
    def __init__(self, normalized_shape: Union[int, Tuple[int, ...]], eps: float = 1e-5, elementwise_affine: bool = True):
        self.normalized_shape = (normalized_shape,) if isinstance(normalized_shape, int) else tuple(normalized_shape)
        self.axis, self.elementwise_affine = tuple(-1 - i for i in range(len(self.normalized_shape))), elementwise_affine
        dev = _default_device()
        self.weight = torch.ones(*self.normalized_shape, dtype=F32, device=dev) if elementwise_affine else None
        self.bias = torch.zeros(*self.normalized_shape, dtype=F32, device=dev) if elementwise_affine else None
        self.eps = np.full((1, 1, 1, 1), eps, dtype=np.float32)

    def _packed(self):
        return packing.cached(self, "ln", (self.weight, self.bias),
                              lambda: (packing.f32(self.weight), packing.f32(self.bias)))

    def __call__(self, x):
        assert self.normalized_shape == tuple(x.shape[-len(self.normalized_shape):]), \
            f"last dimensions of {x.shape} must match {self.normalized_shape}"
        out_shape = x.shape
        x4 = x.reshape(1, -1, x.shape[-2], x.shape[-1]) if x.dim() == 3 else x.reshape(1, 1, -1, x.shape[-1])
        g, b = self._packed()
        y = layer_norm(x4, g.reshape(1, 1, 1, -1), b.reshape(1, 1, 1, -1), self.eps)
        return y.reshape(out_shape)

    # fast path on a (rows = B*T, C) fp16 buffer
    def _run(self, ctx, x_ptr, out_ptr, B, T, C):
        g, b = self._packed()
        ctx.layernorm(x_ptr, out_ptr, B * T, C, g.data_ptr(), b.data_ptr(), float(self.eps.reshape(-1)[0]),
                      B if ctx.ln_strided else 1)
