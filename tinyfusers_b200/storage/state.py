"""State-dict injection with the reference's signature (reference: tinyfusers/storage/state.py:4-23).

Walks the object tree — instance `__dict__`, namedtuples, lists/tuples (index becomes a key component),
dicts — and replaces every `weight` / `bias` entry by the tensor stored under the dotted checkpoint key.
Values may be torch tensors, numpy arrays or anything with `.numpy()` (tinygrad tensors in the reference)."""
from collections import OrderedDict

import numpy as np
import torch

from .. import packing


def _default_device():
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


def _to_device_tensor(v, device):
    if isinstance(v, torch.Tensor):
        return v.detach().to(device=device, dtype=torch.float32)
    arr = v.numpy() if hasattr(v, "numpy") else np.asarray(v)
    return torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32)).to(device)


def update_state(obj, state_dict, prefix=''):
    if hasattr(obj, '__dict__') and not isinstance(obj, torch.Tensor):
        update_state(obj.__dict__, state_dict, f"{prefix}")
    elif hasattr(obj, '_asdict'):
        update_state(obj._asdict(), state_dict, prefix)
    elif isinstance(obj, OrderedDict):
        update_state(dict(obj), state_dict, prefix)
    elif isinstance(obj, (list, tuple)):
        for i, x in enumerate(obj):
            update_state(x, state_dict, f"{prefix}.{str(i)}")
    elif isinstance(obj, dict):
        for k, v in obj.items():
            if isinstance(k, str) and k.startswith("_"):
                continue  # private attributes of this implementation (packed-weight caches, engines), not model state
            if k in {"weight", "bias"}:
                key = f"{prefix}.{k}"
                if key not in state_dict:
                    print(f"skipped: {key}")
                    continue
                obj[k] = _to_device_tensor(state_dict[key], _default_device())
                packing.bump_generation()   # captured sampler graphs bake in packed-weight addresses: they re-capture
            else:
                pre = f"{prefix}.{k}" if prefix != '' else f"{k}"
                update_state(v, state_dict, f"{pre}")


def get_state_dict(net):
    pass
