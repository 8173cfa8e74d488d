"""scaled_dot_product_attention with the reference's signature (reference: tinyfusers/attention/sdpa.py:53-77).

q (B,NH,Tq,HS), k/v (B,NH,Tk,HS) -> (B,NH,Tq,HS) fp32. One fused tcgen05 kernel (tf_attention_f16) instead of
cuBLAS QK^T -> HBM scores -> softmax_kernel -> cuBLAS PV. The operand re-layout below (head padding,
V transpose) is container-level data movement for this stand-alone entry point only; inside the UNet
the projection GEMMs write these layouts directly."""
import torch

from ..runtime import F16, F32, require_cuda, standalone_context


def scaled_dot_product_attention(q_cp, k_cp, v_cp, attn_mask=None):
    require_cuda(q_cp, "q")
    if attn_mask is not None:
        raise RuntimeError("tinyfusers_b200 scaled_dot_product_attention: attn_mask is not built yet "
                           "(needed only by the CLIP text encoder, SURVEY.md §8f-2)")
    ctx = standalone_context()
    B, NH, Tq, HS = q_cp.shape
    Tk = k_cp.shape[-2]
    if HS % 8 != 0 or v_cp.shape[-1] != HS:
        raise RuntimeError(f"tinyfusers_b200 scaled_dot_product_attention: head size {HS} must be a multiple of 8 "
                           "and equal for q/k/v")
    dp = (HS + 15) // 16 * 16
    Tkp = (Tk + 7) // 8 * 8
    dev = q_cp.device
    Q = torch.zeros((B, Tq, NH, dp), dtype=F16, device=dev)
    Q[..., :HS] = q_cp.permute(0, 2, 1, 3)
    K = torch.zeros((B, Tkp, NH, dp), dtype=F16, device=dev)
    K[:, :Tk, :, :HS] = k_cp.permute(0, 2, 1, 3)
    Vt = torch.zeros((NH, dp, B, Tkp), dtype=F16, device=dev)
    Vt[:, :HS, :, :Tk] = v_cp.permute(1, 3, 0, 2)
    out = torch.empty((B, NH, Tq, HS), dtype=F16, device=dev)
    ctx.attention(Q.data_ptr(), NH * dp, K.data_ptr(), NH * dp, Vt.data_ptr(), B * Tkp, out.data_ptr(), B, NH, Tq, Tk,
                  Tkp, HS, dp, head_major=True)
    return out.to(F32)
