"""Generates tests/golden/*.npz by running the REFERENCE'S OWN Python (imported from /root/reference).

Run here only (the GPU box has no /root/reference):   python oracle/make_golden.py

The reference cannot be imported as-is offline: CuPy, tinygrad and a CUDA driver are
missing, `tinyfusers.tensor` does not exist, and `attention/sdpa.py` asks its RawModule for a kernel name
that `softmax.cu` does not define. This script supplies the minimum stand-ins so that the reference's
model code — the part that decides WHAT is computed — runs unmodified on the CPU:

  * `cupy`            -> numpy (same array API; `cp.cuda.*` synchronisation calls become no-ops)
  * `tinyfusers.native.*` (ctypes over libcuda/libcublas/libnvrtc) -> empty stubs (never called on this path)
  * `tinyfusers.tensor.tensor` -> alias of `tinyfusers.storage.tensor` (the module the model files mean)
  * `cp.RawModule(...).get_function(...)` -> a numpy row softmax with the arithmetic of
    native/cuda/softmax.cu:24-112 (max, exp(x-max), sum, divide) writing into the output array
  * `cudnn` conv_fprop graph (vision/conv2d.py:9-28) -> torch.nn.functional.conv2d — the equivalence the
    reference's own test asserts (tests/conv2d.py:27-33)
  * `cudnn` layernorm graph (ff/layer_norm.py:8-32) -> numpy LayerNorm over the last dimension. Real cuDNN 9.x,
    fed the reference's exact descriptor on the B200 box (oracle/cudnn_probe.py ->
    tests/golden/cudnn_layernorm_probe.json), returns exactly this at B = 1 and REJECTS the descriptor
    (CUDNN_STATUS_NOT_SUPPORTED) at B > 1, where the declared strides [B*T*C, 1, B*C, B] stop being the
    contiguous layout; the stand-in applies the B = 1 semantics per batch element.

Everything else — GroupNorm, activations, Linear, GEGLU, SDPA wiring, CrossAttention's head reshape,
transformer / ResBlock / UNet composition, skip concatenation order, CFG combine, DDIM update, alpha
schedule, timestep embedding — is the reference's code executing.
Weights / inputs come from oracle.ref_ops' seeded generators so tests can regenerate them.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")


def install_shims():
    cp = types.ModuleType("cupy")
    for name in dir(np):
        if not name.startswith("__"):
            setattr(cp, name, getattr(np, name))
    cp.asnumpy = lambda a: np.asarray(a)
    cp.single = np.float32

    class _Stream:
        def use(self):
            return self

        def synchronize(self):
            pass

    class _Device:
        def synchronize(self):
            pass

    cuda = types.SimpleNamespace(get_current_stream=lambda: _Stream(), Device=lambda *a: _Device(),
                                 runtime=types.SimpleNamespace(deviceSynchronize=lambda: None))
    cp.cuda = cuda
    rnd = types.SimpleNamespace(uniform=lambda lo, hi, size=None, dtype=np.float32: np.zeros(size, dtype=dtype))
    cp.random = rnd  # constructor init values are irrelevant: update_state overwrites every weight

    def _softmax_kernel(grid=None, block=None, args=None, shared_mem=None):
        out, inp, N, C = args
        x = np.asarray(inp, dtype=np.float32).reshape(N, C)
        m = x.max(axis=1, keepdims=True)
        e = np.exp(x - m, dtype=np.float32)
        out.reshape(N, C)[...] = e / e.sum(axis=1, keepdims=True, dtype=np.float32)

    class _RawModule:
        def __init__(self, *a, **k):
            pass

        def get_function(self, name):
            return _softmax_kernel

    cp.RawModule = _RawModule
    sys.modules["cupy"] = cp

    cudnn = types.ModuleType("cudnn")
    cudnn.create_handle = lambda: None
    cudnn.data_type = types.SimpleNamespace(FLOAT=0, HALF=1)
    sys.modules["cudnn"] = cudnn

    # ctypes bindings over GPU driver libraries (libcuda / libcublas / libnvrtc): never reached on the UNet
    # path, but imported at package import time (tinyfusers/__init__.py:1 -> storage/device.py:6) -> stubs,
    # registered before the first `import tinyfusers`
    sys.path.insert(0, REF)
    native = types.ModuleType("tinyfusers.native")
    native.__path__ = []
    native.cuda = native.cudart = native.nvrtc = native.cublas = types.SimpleNamespace()
    sys.modules["tinyfusers.native"] = native
    for sub in ("cuda", "cublas", "nvrtc"):
        m = types.ModuleType(f"tinyfusers.native.{sub}")
        m.__path__ = []
        ops = types.ModuleType(f"tinyfusers.native.{sub}.ops")
        ops.cuda = ops.cudart = ops.nvrtc = ops.cublas = types.SimpleNamespace()
        m.ops = ops
        sys.modules[f"tinyfusers.native.{sub}"] = m
        sys.modules[f"tinyfusers.native.{sub}.ops"] = ops
    import tinyfusers  # noqa: F401
    import tinyfusers.storage.tensor as st
    pkg = types.ModuleType("tinyfusers.tensor")
    pkg.tensor = st
    sys.modules["tinyfusers.tensor"] = pkg
    sys.modules["tinyfusers.tensor.tensor"] = st

    # cuDNN-backed operators
    import tinyfusers.vision.conv2d as rconv
    import tinyfusers.ff.layer_norm as rln

    def conv_2d(X, W, padding, stride, dilation):
        y = F.conv2d(torch.from_numpy(np.ascontiguousarray(X, dtype=np.float32)),
                     torch.from_numpy(np.ascontiguousarray(W, dtype=np.float32)), None, stride=tuple(stride),
                     padding=tuple(padding), dilation=tuple(dilation))
        return y.numpy()

    def layer_norm(x, scale, bias, eps):
        # what real cuDNN executes for this graph: see the module docstring / oracle/cudnn_probe.py
        v = np.ascontiguousarray(x, dtype=np.float32)
        mean = v.mean(axis=-1, keepdims=True)
        var = ((v - mean) ** 2).mean(axis=-1, keepdims=True)
        y = (v - mean) / np.sqrt(var + np.float32(eps.reshape(-1)[0])) * scale.reshape(-1) + bias.reshape(-1)
        return y.astype(np.float32)

    rconv.conv_2d = conv_2d
    rln.layer_norm = layer_norm


def main():
    os.chdir(REF)  # attention/sdpa.py reads 'tinyfusers/native/cuda/softmax.cu' relative to the CWD
    install_shims()
    sys.path.insert(0, ROOT)
    from oracle import ref_ops as R
    from tinyfusers.attention.attention import BasicTransformerBlock, CrossAttention, SpatialTransformer
    from tinyfusers.attention.sdpa import scaled_dot_product_attention
    from tinyfusers.ff.group_norm import GroupNorm, group_norm
    from tinyfusers.ff.linear import Linear
    from tinyfusers.ff.nn import FeedForward
    from tinyfusers.storage.state import update_state
    from tinyfusers.storage.tensor import Tensor
    from tinyfusers.variants.sd import StableDiffusion, get_alphas_cumprod
    from tinyfusers.vision.resnet import ResBlock
    from tinyfusers.vision.unet import Downsample, UNetModel, Upsample, timestep_embedding
    import contextlib
    import io

    os.makedirs(OUT, exist_ok=True)
    G = {}

    def load(obj, sd, prefix):
        with contextlib.redirect_stdout(io.StringIO()):
            update_state(obj, sd, prefix)

    def rnd(seed, *shape, scale=1.0, shift=0.0):
        g = np.random.Generator(np.random.Philox(seed))
        return (g.standard_normal(shape, dtype=np.float32) * np.float32(scale) + np.float32(shift))

    # ---- activations (storage/tensor.py:64-86) ----
    x = np.linspace(-9, 9, 1801, dtype=np.float32)
    for name in ("sigmoid", "silu", "swish", "gelu", "quick_gelu"):
        G[f"act_{name}"] = getattr(Tensor, name)(x).astype(np.float32)

    # ---- group norm (ff/group_norm.py) ----
    x = rnd(101, 2, 320, 8, 8, scale=1.7, shift=0.6)
    G["group_norm_noaffine"] = group_norm(x, 32, 1e-5)
    gn = GroupNorm(32, 320)
    gn.weight, gn.bias = 1 + 0.1 * rnd(102, 320), 0.1 * rnd(103, 320)
    G["group_norm_affine"] = gn(x)

    # ---- linear / feed-forward ----
    sd = {}
    R._add_linear(sd, "lin", 320, 640, 201)
    lin = Linear(320, 640)
    load(lin, sd, "lin")
    G["linear"] = lin(rnd(202, 2, 16, 320))
    sd = {}
    R._add_linear(sd, "ff.net.0.proj", 320, 2560, 203)
    R._add_linear(sd, "ff.net.2", 1280, 320, 203)
    ff = FeedForward(320)
    load(ff, sd, "ff")
    xin = rnd(204, 2, 16, 320)
    G["geglu"] = ff.net[0](xin)
    G["feed_forward"] = ff(xin)

    # ---- timestep embedding / schedule / DDIM ----
    G["timestep_embedding"] = np.concatenate([timestep_embedding(np.array([t]), 320) for t in (1, 21, 501, 981)])
    G["alphas_cumprod"] = get_alphas_cumprod().astype(np.float32)
    fake = types.SimpleNamespace()
    ac = G["alphas_cumprod"]
    xp, p0 = StableDiffusion.get_x_prev_and_pred_x0(fake, rnd(301, 1, 4, 8, 8), rnd(302, 1, 4, 8, 8), ac[[501]], ac[[481]])
    G["ddim_x_prev"], G["ddim_pred_x0"] = xp.astype(np.float32), p0.astype(np.float32)

    # ---- sdpa (attention/sdpa.py:53-77 + softmax kernel) ----
    q, k, v = rnd(401, 2, 8, 64, 40), rnd(402, 2, 8, 77, 40), rnd(403, 2, 8, 77, 40)
    G["sdpa"] = scaled_dot_product_attention(q, k, v)

    # ---- attention blocks ----
    sd = {}
    R.add_spatial_transformer(sd, "st", 320, 768, seed=501)
    tb = "st.transformer_blocks.0"
    xt, ctx = rnd(502, 2, 64, 320), rnd(503, 2, 77, 768)
    a1 = CrossAttention(320, 320, 8, 40)
    load(a1, sd, tb + ".attn1")
    G["cross_attention_self"] = a1(xt)
    a2 = CrossAttention(320, 768, 8, 40)
    load(a2, sd, tb + ".attn2")
    G["cross_attention_ctx"] = a2(xt, ctx)
    blk = BasicTransformerBlock(320, 768, 8, 40)
    load(blk, sd, tb)
    G["transformer_block_b2"] = blk(xt, ctx)
    G["transformer_block_b1"] = blk(xt[:1], ctx[:1])
    st = SpatialTransformer(320, 768, 8, 40)
    load(st, sd, "st")
    G["spatial_transformer"] = st(rnd(504, 2, 320, 8, 8), ctx)

    # ---- ResBlock / up / down ----
    sd = {}
    R.add_res_block(sd, "rb", 320, 640, seed=601)
    rb = ResBlock(320, 1280, 640)
    load(rb, sd, "rb")
    G["res_block"] = rb(rnd(602, 2, 320, 8, 8), rnd(603, 1, 1280))
    sd = {}
    R._add_conv(sd, "u.conv", 64, 64, 3, 604)
    R._add_conv(sd, "d.op", 64, 64, 3, 604)
    up, down = Upsample(64), Downsample(64)
    load(up, sd, "u")
    load(down, sd, "d")
    G["upsample"] = up(rnd(605, 1, 64, 6, 6))
    G["downsample"] = down(rnd(606, 1, 64, 12, 12))

    # ---- full UNet, CFG and one sampler step (variants/sd.py:27-59) at a 16x16 latent ----
    usd = R.make_unet_state_dict(seed=1234)
    unet = UNetModel()
    load(unet, usd, "model.diffusion_model")
    lat, unc, cx = R.make_inputs(1, 16)
    lat, unc, cx = lat.numpy(), unc.numpy(), cx.numpy()
    x2, c2 = np.concatenate([lat, lat]), np.concatenate([unc, cx])
    G["unet_16"] = unet(x2, np.array([981]), c2).astype(np.float32)
    model = types.SimpleNamespace(model=types.SimpleNamespace(diffusion_model=unet))
    e_t = StableDiffusion.get_model_output(model, unc, cx, lat, np.array([501]), np.array([7.5]))
    G["cfg_e_t_16"] = np.asarray(e_t, dtype=np.float32)
    model.get_model_output = lambda *a: StableDiffusion.get_model_output(model, *a)
    model.get_x_prev_and_pred_x0 = lambda *a: StableDiffusion.get_x_prev_and_pred_x0(model, *a)
    xprev = StableDiffusion.__call__(model, unc, cx, lat, np.array([501]), ac[[25]], ac[[24]], np.array([7.5]))
    G["sampler_step_16"] = np.asarray(xprev, dtype=np.float32)

    np.savez_compressed(os.path.join(OUT, "reference_outputs.npz"), **{k: np.asarray(v, dtype=np.float32) for k, v in G.items()})
    for k, v in G.items():
        print(f"{k:28s} {str(np.asarray(v).shape):20s} absmax {np.abs(v).max():.4f}")


if __name__ == "__main__":
    main()
