"""Host logic of the precision switch (no GPU): fp16 is the default, unknown modes are rejected, and the fp32 parity
mode is no CPU fallback either - every operator still refuses host tensors."""
import pytest
import torch


def test_precision_switch_and_no_cpu_path():
    import tinyfusers_b200
    from tinyfusers_b200.ff.group_norm import GroupNorm
    from tinyfusers_b200.ff.linear import Linear
    from tinyfusers_b200.vision.conv2d import Conv2d
    assert tinyfusers_b200.get_precision() == "fp16"
    with pytest.raises(ValueError):
        tinyfusers_b200.set_precision("bf16")
    tinyfusers_b200.set_precision("fp32")
    try:
        assert tinyfusers_b200.get_precision() == "fp32"
        if not torch.cuda.is_available():
            for call in (lambda: Linear(8, 8)(torch.zeros(2, 8)),
                         lambda: Conv2d(8, 8, kernel_size=[3, 3], padding=[1, 1])(torch.zeros(1, 8, 4, 4)),
                         lambda: GroupNorm(4, 8)(torch.zeros(1, 8, 4, 4))):
                with pytest.raises(RuntimeError):
                    call()
    finally:
        tinyfusers_b200.set_precision("fp16")


def test_fp32_entry_points_are_bound():
    from tinyfusers_b200.native.b200.ops import b200
    for name in ("tf_gemm_f32", "tf_conv2d_nchw_f32", "tf_groupnorm_nchw_f32", "tf_layernorm_f32", "tf_softmax_rows_f32",
                 "tf_unary_f32", "tf_geglu_f32", "tf_cfg_combine_f32", "tf_embedding_f32", "tf_softmax_rows_f32_to_f16"):
        assert callable(getattr(b200, name))
    # argument validation runs before any device work: a null pointer is an argument error, not a crash
    assert b200.tf_layernorm_f32(None, None, None, None, 4, 8, 1e-5, None) == b200.TF_ERR_ARG
    assert "tf_layernorm_f32" in b200.last_error()
