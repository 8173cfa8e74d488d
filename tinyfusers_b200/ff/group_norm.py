"""GroupNorm with the reference's signatures (reference: tinyfusers/ff/group_norm.py:3-21)."""
import torch

from .. import fp32, packing
from ..runtime import F32, act_to_nchw, nchw_to_act, new_act_tensor, standalone_context
from ..storage.state import _default_device


def group_norm(x, num_groups, eps):
    """(x - mean) / sqrt(biased_var + eps) per (n, group); no affine (group_norm.py:3-11)."""
    if fp32.enabled():
        return fp32.group_norm(x, num_groups, eps)
    ctx = standalone_context()
    a = nchw_to_act(x, c_pad_to=8)
    C = x.shape[1]
    if a.c != C:
        raise RuntimeError(f"group_norm: channel count {C} must be a multiple of 8 for the B200 kernel")
    out = new_act_tensor(a.n, a.h, a.w, a.c, device=x.device)
    ctx.groupnorm(a, out, None, None, float(eps), silu=False, groups=num_groups)
    return act_to_nchw(out, C)


class GroupNorm:
    def __init__(self, num_groups: int, num_channels: int, eps: float = 1e-5, affine: bool = True):
        self.num_groups, self.num_channels, self.eps = num_groups, num_channels, eps
        dev = _default_device()
        self.weight = torch.ones(num_channels, dtype=F32, device=dev) if affine else None
        self.bias = torch.zeros(num_channels, dtype=F32, device=dev) if affine else None

    def _packed(self):
        return packing.cached(self, "gn", (self.weight, self.bias),
                              lambda: (packing.f32(self.weight), packing.f32(self.bias)))

    def __call__(self, x):
        if fp32.enabled():
            return fp32.group_norm(x, self.num_groups, self.eps, self.weight, self.bias)
        ctx = standalone_context()
        a = nchw_to_act(x, c_pad_to=8)
        if a.c != self.num_channels:
            raise RuntimeError(f"GroupNorm: got {x.shape[1]} channels, expected {self.num_channels} (multiple of 8)")
        out = new_act_tensor(a.n, a.h, a.w, a.c, device=x.device)
        self._run(ctx, a, out, silu=False)
        return act_to_nchw(out, self.num_channels)

    # fast path: NHWC fp16 in -> NHWC fp16 out, optional fused SiLU
    def _run(self, ctx, x, out, silu):
        g, b = self._packed()
        ctx.groupnorm(x, out, g.data_ptr() if g is not None else None, b.data_ptr() if b is not None else None,
                      float(self.eps), silu, self.num_groups)
        return out
