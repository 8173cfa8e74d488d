"""The C-ABI library loads on a CPU-only box and exports every symbol include/tinyfusers_b200.h declares;
the ctypes binding declares the same set (no compute calls here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "tinyfusers_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tf_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from tinyfusers_b200.csrc.build import build
    lib = ctypes.CDLL(build())
    names = _header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"


def test_binding_covers_header():
    from tinyfusers_b200.native.b200 import ops
    assert sorted(ops._SIGNATURES) == _header_functions()
    assert ops.b200.tf_version() >= 100


def test_no_cpu_fallback_without_gpu():
    import pytest
    import torch
    from tinyfusers_b200.native.b200.ops import b200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert b200.tf_init(0) != 0
    assert "no CUDA device" in b200.last_error() or "sm_" in b200.last_error()
    from tinyfusers_b200.ff.linear import Linear
    with pytest.raises(RuntimeError):
        Linear(8, 8)(torch.zeros(2, 8))


def test_header_cites_reference_for_each_operator():
    text = open(os.path.join(ROOT, "include", "tinyfusers_b200.h")).read()
    for cite in ("tinyfusers/ff/linear.py", "tinyfusers/vision/conv2d.py", "tinyfusers/ff/group_norm.py",
                 "tinyfusers/ff/layer_norm.py", "tinyfusers/attention/sdpa.py", "tinyfusers/variants/sd.py",
                 "tinyfusers/vision/unet.py", "tinyfusers/storage/tensor.py"):
        assert cite in text, cite
