"""Conv2d with the reference's signatures (reference: tinyfusers/vision/conv2d.py:9-59).

The reference builds a cuDNN conv_fprop graph on every call, zero-fills the output, transposes
NHWC->NCHW and adds the bias in a separate pass. Here a convolution is one tf_conv2d_nhwc_f16 launch:
NHWC fp16 implicit GEMM on tcgen05 with bias / residual fused in the epilogue."""
import functools
import math
import operator

import torch

from .. import fp32, packing
from ..native.b200.ops import b200
from ..runtime import F16, F32, Act, act_to_nchw, nchw_to_act, new_act_tensor, require_cuda, standalone_context, stream_ptr
from ..storage.state import _default_device


def _pair(v):
    return [int(v), int(v)] if isinstance(v, int) else [int(v[0]), int(v[1])]


def _check_supported(ksize, stride, padding, dilation):
    k, s, p, d = _pair(ksize), _pair(stride), _pair(padding), _pair(dilation)
    ok = d == [1, 1] and ((k == [3, 3] and p == [1, 1] and s in ([1, 1], [2, 2])) or
                          (k == [1, 1] and p == [0, 0] and s == [1, 1]))
    if not ok:
        raise RuntimeError(f"tinyfusers_b200 conv2d: kernel {k} stride {s} padding {p} dilation {d} has no B200 kernel "
                           "(built: 3x3 pad 1 stride 1|2, 1x1 pad 0 stride 1)")
    return k[0], s[0]


def _conv_act(ctx, x, w_packed, cout_p, k, stride, out, bias_ptr=None, residual=None, flags=0, w_static=True, row_stats=None):
    # the output's GroupNorm statistics are produced by this launch when `out` carries a (single-part) buffer
    gn = None
    if out.gn is not None and len(out.gn) == 1 and out.gn[0][0] == 0 and out.gn[0][1] == cout_p and flags == 0:
        gn = (out.gn[0][2], out.gn[0][3])
    if k == 3:
        ctx.conv3x3(x, w_packed.data_ptr(), cout_p, out, bias=bias_ptr, residual=residual, stride=stride, flags=flags, gn=gn,
                    w_static=w_static)
    else:
        ctx.gemm(x.ptr, x.stride, x.rows, x.c, w_packed.data_ptr(), cout_p, out.ptr, out.stride, bias=bias_ptr,
                 residual_ptr=residual.ptr if residual is not None else None,
                 ldr=residual.stride if residual is not None else 0, flags=flags,
                 gn=(gn[0], gn[1], out.h * out.w) if gn else None, w_static=w_static, row_stats=row_stats)


def conv_2d(X_gpu, W_gpu, padding, stride, dilation):
    """NCHW cross-correlation without bias -> NCHW fp32 (reference: conv2d.py:9-28)."""
    require_cuda(X_gpu, "X_gpu")
    if fp32.enabled():
        if fp32._sq(dilation, "dilation") != 1:
            raise RuntimeError("tinyfusers_b200 conv_2d (fp32): dilation has no kernel")
        return fp32.conv2d(X_gpu, W_gpu, None, fp32._sq(stride, "stride"), fp32._sq(padding, "padding"))
    k, s = _check_supported(W_gpu.shape[2:], stride, padding, dilation)
    ctx = standalone_context()
    O = W_gpu.shape[0]
    if k == 3:
        w = packing.conv3x3_weight(W_gpu, 64, 8)
        a = nchw_to_act(X_gpu, c_pad_to=64)
    else:
        w = packing.conv1x1_weight(W_gpu, 8)
        a = nchw_to_act(X_gpu, c_pad_to=8)
    Op = w.shape[0]
    Ho, Wo = (a.h + 2 * (k // 2) - k) // s + 1, (a.w + 2 * (k // 2) - k) // s + 1
    out = new_act_tensor(a.n, Ho, Wo, Op, device=X_gpu.device)
    _conv_act(ctx, a, w, Op, k, s, out, w_static=False)    # w was packed a moment ago on this stream: not a static weight
    return act_to_nchw(out, O)


def conv_2d_16(X_gpu, W_gpu, padding, stride, dilation):
    """fp16-I/O variant of the reference (conv2d.py:31-46); the B200 kernel is fp16-I/O already."""
    return conv_2d(X_gpu, W_gpu, padding, stride, dilation).to(F16)


class Conv2d:
    def __init__(self, in_channels, out_channels, kernel_size, stride=[1, 1], padding=[0, 0], dilation=[1, 1], groups=1, bias=True):
        self.kernel_size = kernel_size
        self.stride, self.padding, self.dilation, self.groups = stride, padding, dilation, groups
        dev = _default_device()
        shape = (out_channels, in_channels // self.groups, *self.kernel_size)
        if dev.type == "cuda":  # reference init U(-sqrt3, sqrt3) / U(+-1/sqrt(fan_in)) (conv2d.py:52-54)
            self.weight = (torch.rand(shape, dtype=F32, device=dev) * 2 - 1) * math.sqrt(3.0)
            bound = 1 / math.sqrt(functools.reduce(operator.mul, shape[1:], 1))
            self.bias = (torch.rand((out_channels,), dtype=F32, device=dev) * 2 - 1) * bound if bias else None
        else:  # CPU: weights are containers only (filled by update_state); nothing computes here
            self.weight = torch.zeros(shape, dtype=F32)
            self.bias = torch.zeros((out_channels,), dtype=F32) if bias else None

    def _geometry(self):
        return _check_supported(self.kernel_size, self.stride, self.padding, self.dilation)

    def _packed(self):
        k, _ = self._geometry()
        def build():
            w = packing.conv3x3_weight(self.weight, 64, 8) if k == 3 else packing.conv1x1_weight(self.weight, 8)
            b = packing.f32(self.bias)
            if b is not None and b.shape[0] != w.shape[0]:
                b = torch.nn.functional.pad(b, (0, w.shape[0] - b.shape[0]))
            return w, b
        return packing.cached(self, "conv", (self.weight, self.bias), build)

    def __call__(self, x):
        require_cuda(x, "x")
        if fp32.enabled():
            return fp32.conv_module(self, x)
        ctx = standalone_context()
        k, s = self._geometry()
        O, I = self.weight.shape[0], self.weight.shape[1]
        if k == 3 and I == 4 and s == 1 and x.dtype == F32:
            out = new_act_tensor(x.shape[0], x.shape[2], x.shape[3], O, device=x.device)
            self._run_smallcin(ctx, x.contiguous(), x.shape[0], out)
            return act_to_nchw(out, O)
        a = nchw_to_act(x, c_pad_to=64 if k == 3 else 8)
        w, _ = self._packed()
        Ho, Wo = (a.h + 2 * (k // 2) - k) // s + 1, (a.w + 2 * (k // 2) - k) // s + 1
        out = new_act_tensor(a.n, Ho, Wo, w.shape[0], device=x.device)
        self._run(ctx, a, out)
        return act_to_nchw(out, O)

    # fast path: NHWC fp16 Act -> NHWC fp16 (or fp32 with TF_EPI_OUT_F32) Act, epilogue fused
    def _run(self, ctx, x, out, residual=None, bias_ptr="own", flags=0, row_stats=None):
        k, s = self._geometry()
        w, b = self._packed()
        if bias_ptr == "own":
            bias_ptr = b.data_ptr() if b is not None else None
        if x.c != w.shape[-1] and k == 3:
            raise RuntimeError(f"Conv2d fast path: activation has {x.c} channels, packed weight expects {w.shape[-1]}")
        _conv_act(ctx, x, w, w.shape[0], k, s, out, bias_ptr, residual, flags, row_stats=row_stats if k == 1 else None)
        return out

    # Cin = 4 input convolution straight from the fp32 NCHW latent (UNet conv_in)
    def _run_smallcin(self, ctx, x_f32_nchw, n_out, out):
        if ctx.skip("misc"):
            return out
        wb = packing.cached(self, "conv_in", (self.weight, self.bias),
                            lambda: (packing.f32(self.weight), packing.f32(self.bias)))
        N, C, H, W = x_f32_nchw.shape
        gn = out.gn[0] if out.gn is not None and len(out.gn) == 1 and (H * W) % 32 == 0 else None
        if gn is not None:     # the output's GroupNorm statistics come out of this launch
            st = b200.tf_conv3x3_smallcin_gn_f32nchw(x_f32_nchw.data_ptr(), N, wb[0].data_ptr(),
                                                     wb[1].data_ptr() if wb[1] is not None else None, out.ptr, n_out, C, H,
                                                     W, self.weight.shape[0], out.stride, gn[2], gn[3], stream_ptr())
            b200.check(st, "tf_conv3x3_smallcin_gn_f32nchw")
            return out
        out.gn = None
        st = b200.tf_conv3x3_smallcin_f32nchw(x_f32_nchw.data_ptr(), N, wb[0].data_ptr(),
                                              wb[1].data_ptr() if wb[1] is not None else None, out.ptr, n_out, C, H, W,
                                              self.weight.shape[0], out.stride, stream_ptr())
        b200.check(st, "tf_conv3x3_smallcin_f32nchw")
        return out


def conv3x3_plus_skip(ctx, owner, conv, skip, x, x2, out):
    """out = conv(x) + skip(x2) with conv a 3x3 stride-1 Conv2d and skip a 1x1 Conv2d, as ONE implicit GEMM: the skip
    convolution's weight is appended to every row of the 3x3 weight (extra k-blocks read from x2), the biases add.
    `owner` caches the packed pair. Returns False when the geometry does not qualify (caller falls back to two launches)."""
    if x2.c % 64 != 0 or x.c % 64 != 0 or conv._geometry() != (3, 1) or skip._geometry() != (1, 1):
        return False
    def build():
        w3, b3 = conv._packed()
        w1, b1 = skip._packed()
        cout = w3.shape[0]
        w = torch.cat((w3.reshape(cout, -1), w1.reshape(w1.shape[0], -1)[:cout]), dim=1).contiguous()
        b = None
        if b3 is not None or b1 is not None:
            b = (b3 if b3 is not None else 0) + (b1[:cout] if b1 is not None else 0)
            b = b.contiguous()
        return w, b
    w, b = packing.cached(owner, "conv_skip", (conv.weight, conv.bias, skip.weight, skip.bias), build)
    if w.shape[1] != 9 * x.c + x2.c:
        raise RuntimeError(f"conv3x3_plus_skip: packed weight has {w.shape[1]} columns, expected {9 * x.c + x2.c}")
    gn = None
    if out.gn is not None and len(out.gn) == 1 and out.gn[0][0] == 0 and out.gn[0][1] == w.shape[0]:
        gn = (out.gn[0][2], out.gn[0][3])
    ctx.conv3x3_skip(x, x2, w.data_ptr(), w.shape[0], out, bias=b.data_ptr() if b is not None else None, gn=gn)
    return True
