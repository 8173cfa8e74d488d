"""Embedding with the reference's signature (reference: tinyfusers/ff/embedding.py:6-23).

The reference builds a one-hot matrix on the host and multiplies it with the table through cuBLAS — and allocates
that matrix with the wrong shape (embed_sz x N instead of vocab x N, embedding.py:17), so it cannot run. What it
stands for is the row lookup `weight[idx]`; that is one gather kernel here."""
import numpy as np
import torch

from ..native.b200.ops import b200
from ..runtime import F16, F32, as_f32, stream_ptr
from ..storage.state import _default_device


def _ids_tensor(idx, device, vocab=None):
    """Token ids -> int32 device tensor. With `vocab`, ids outside [0, vocab) raise (the gather kernel would clamp them, and a
    silently clamped id is a wrong embedding): checked on the host copy the ids arrive as (numpy / CPU tensor), or with one
    device reduction for ids that are already on the GPU."""
    if isinstance(idx, torch.Tensor):
        t = idx
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(idx), dtype=np.int64))
    if vocab is not None and t.numel():
        lo, hi = int(t.min()), int(t.max())
        if lo < 0 or hi >= vocab:
            raise RuntimeError(f"tinyfusers_b200 embedding: token id out of range [0, {vocab}): min {lo}, max {hi}")
    return t.to(device=device, dtype=torch.int32).contiguous()


def embedding(weight, idx):
    """weight: (vocab, E) fp32 CUDA tensor; idx: (1, N) integer ids -> (N, E) fp32."""
    ids = _ids_tensor(idx, weight.device, weight.shape[0]).reshape(-1)
    N, E = ids.numel(), weight.shape[1]
    out = torch.empty((N, E), dtype=F16, device=weight.device)
    w = weight.to(F32).contiguous()
    st = b200.tf_embedding_f16(ids.data_ptr(), w.data_ptr(), None, out.data_ptr(), N, N, E, w.shape[0], stream_ptr())
    b200.check(st, "tf_embedding_f16")
    return as_f32(out)


class Embedding:
    def __init__(self, vocab_size: int, embed_size: int):
        self.vocab_sz = vocab_size
        self.embed_sz = embed_size
        self.weight = torch.ones((vocab_size, embed_size), dtype=F32, device=_default_device())

    def __call__(self, idx):
        return embedding(self.weight, idx)
