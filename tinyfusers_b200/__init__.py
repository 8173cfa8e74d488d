"""tinyfusers_b200 — B200-native drop-in for the denoising hot path of Fatlonder/tinyfusers.

The module tree mirrors the reference (`tinyfusers/{variants,vision,attention,ff,storage,native}`), the
classes keep their names, constructor signatures, attribute names (== checkpoint keys) and call
signatures; every arithmetic operation runs in hand-written sm_100a CUDA kernels behind the ctypes C-ABI
in `tinyfusers_b200.native.b200` (include/tinyfusers_b200.h). torch tensors are containers only.

`set_quirks(True)` (default) reproduces the reference literally, including its LayerNorm stride
declaration at batch > 1 and the CrossAttention head-major reshape (SURVEY.md §8 parity notes);
`set_quirks(False)` gives canonical Stable Diffusion semantics for real checkpoints.
"""
_QUIRKS = True


def set_quirks(flag: bool):
    global _QUIRKS
    _QUIRKS = bool(flag)


def get_quirks() -> bool:
    return _QUIRKS
