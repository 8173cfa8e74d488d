"""scaled_dot_product_attention with the reference's signature (reference: tinyfusers/attention/sdpa.py:53-77).

q (B,NH,Tq,HS), k/v (B,NH,Tk,HS) -> (B,NH,Tq,HS) fp32. One fused tcgen05 kernel (tf_attention_f16) instead of
cuBLAS QK^T -> HBM scores -> softmax_kernel -> cuBLAS PV. The operand re-layout below (head padding)
is container-level data movement for this stand-alone entry point only; inside the UNet
the projection GEMMs write these layouts directly."""
import torch

from .. import fp32
from ..runtime import F16, F32, as_f32, require_cuda, standalone_context


def scaled_dot_product_attention(q_cp, k_cp, v_cp, attn_mask=None):
    require_cuda(q_cp, "q")
    B, NH, Tq, HS = q_cp.shape
    Tk = k_cp.shape[-2]
    causal = False
    if attn_mask is not None:
        # the one mask the reference ever passes is CLIP's causal one (vae/encoder.py:79): additive triu(-inf, k=1), or
        # its boolean form (sdpa.py:67-68: True = keep). That is a kernel flag; arbitrary masks are not built.
        m = torch.as_tensor(attn_mask, device=q_cp.device)
        keep = m.reshape(-1, m.shape[-2], m.shape[-1])[0]
        keep = keep if keep.dtype == torch.bool else (keep == 0)
        tril = torch.ones((Tq, Tk), dtype=torch.bool, device=q_cp.device).tril()
        if Tq != Tk or m.numel() != Tq * Tk or not torch.equal(keep, tril) or \
                (m.dtype != torch.bool and not torch.isinf(m.reshape(Tq, Tk)[~tril]).all()):
            raise RuntimeError("tinyfusers_b200 scaled_dot_product_attention: only the causal mask "
                               "(triu(-inf, k=1) / its boolean form) is built")
        causal = True
    if fp32.enabled():
        return fp32.scaled_dot_product_attention(q_cp, k_cp, v_cp, causal)
    ctx = standalone_context()
    if HS % 8 != 0 or v_cp.shape[-1] != HS:
        raise RuntimeError(f"tinyfusers_b200 scaled_dot_product_attention: head size {HS} must be a multiple of 8 "
                           "and equal for q/k/v")
    dp = (HS + 15) // 16 * 16
    dvp = (HS + 63) // 64 * 64
    Tkp = (Tk + 7) // 8 * 8
    dev = q_cp.device
    Q = torch.zeros((B, Tq, NH, dp), dtype=F16, device=dev)
    Q[..., :HS] = q_cp.permute(0, 2, 1, 3)
    K = torch.zeros((B, Tkp, NH, dp), dtype=F16, device=dev)
    K[:, :Tk, :, :HS] = k_cp.permute(0, 2, 1, 3)
    V = torch.zeros((B, Tkp, NH, dvp), dtype=F16, device=dev)     # natural layout, head dim zero-padded to 64
    V[:, :Tk, :, :HS] = v_cp.permute(0, 2, 1, 3)
    out = torch.empty((B, NH, Tq, HS), dtype=F16, device=dev)
    ctx.attention_v(Q.data_ptr(), NH * dp, K.data_ptr(), NH * dp, V.data_ptr(), NH * dvp, out.data_ptr(), B, NH, Tq, Tk, Tkp,
                    HS, dp, dvp, head_major=True, causal=causal)
    return as_f32(out)
