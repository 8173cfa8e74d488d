"""tinyfusers_b200 — B200-native drop-in for the denoising hot path of Fatlonder/tinyfusers.

The module tree mirrors the reference (`tinyfusers/{variants,vision,attention,ff,storage,native}`), the
classes keep their names, constructor signatures, attribute names (== checkpoint keys) and call
signatures; every arithmetic operation runs in hand-written sm_100a CUDA kernels behind the ctypes C-ABI
in `tinyfusers_b200.native.b200` (include/tinyfusers_b200.h). torch tensors are containers only.

`set_quirks(True)` (default) reproduces the reference's CrossAttention head-major reshape
(attention/attention.py:39, SURVEY.md §8 parity note 2); `set_quirks(False)` gives the canonical head merge for
real checkpoints. LayerNorm is canonical: real cuDNN executes the reference's layernorm graph only at batch 1
(where it is canonical) and rejects its stride declaration at batch > 1 (oracle/cudnn_probe.py); the literal
reading of those strides stays available as `set_layernorm_strided(True)`.
"""
_QUIRKS = True
_LN_STRIDED = False


def set_quirks(flag: bool):
    global _QUIRKS
    _QUIRKS = bool(flag)


def get_quirks() -> bool:
    return _QUIRKS


def set_layernorm_strided(flag: bool):
    global _LN_STRIDED
    _LN_STRIDED = bool(flag)


def get_layernorm_strided() -> bool:
    return _LN_STRIDED


import os as _os

# OFF by default: measured on B200 (tools/ab_prefetch.py, profiles/ab_weight_prefetch_r2.json) 3.617 ms/step without, 3.623 with,
# at batch 16 18.54 vs 18.48 - the small-M layers are bound by the L2 -> SM operand feed (every CTA of an output row block
# re-reads the activation tile), not by the cold HBM fetch of their weights, which the TMA ring already overlaps.
_WEIGHT_PREFETCH = _os.environ.get("TINYFUSERS_B200_WEIGHT_PREFETCH", "0") == "1"


def set_weight_prefetch(flag: bool):
    """Captured steps ask L2 for the NEXT layer's weights while the current layer drains (include/tinyfusers_b200.h,
    tf_weight_prefetch_mode). Affects graphs captured after the call; results are identical either way."""
    global _WEIGHT_PREFETCH
    _WEIGHT_PREFETCH = bool(flag)


def weight_prefetch_enabled() -> bool:
    return _WEIGHT_PREFETCH


def set_precision(mode: str):
    """"fp16" (default: the tcgen05 path, <= 1e-2) or "fp32" (parity mode of BASELINE.json configs[0], <= 1e-5;
    plain-fp32 kernels in the reference's layouts - tinyfusers_b200/fp32.py, csrc/tf_fp32.cu)."""
    from . import fp32
    fp32.set_precision(mode)


def get_precision() -> str:
    from . import fp32
    return fp32.get_precision()


def __getattr__(name):
    # `tinyfusers.Tensor` (reference: tinyfusers/__init__.py:1), resolved lazily: importing the package itself must not
    # need the built shared library (the build script lives inside the package)
    if name == "Tensor":
        from .storage.tensor import Tensor
        return Tensor
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
