"""GPU-box probe (test infrastructure): runs the reference's LayerNorm and conv graphs on the REAL cuDNN
(python frontend, as the reference does) to pin the two cuDNN-dependent pieces of the oracle:

  1. ff/layer_norm.py:8-32 — the graph is fed a contiguous (1,B,T,C) buffer but DECLARES the strides
     [B*T*C, 1, B*C, B] (layer_norm.py:10). oracle.ref_ops.layer_norm(quirks=True) models what cuDNN then
     does as "view memory as (T, C, B), normalise over C". This script measures the model's error against
     cuDNN itself for B = 1, 2, 4.
  2. vision/conv2d.py:9-28 — conv_fprop writes NHWC (frontend default) which the reference re-reads as NCHW
     (conv2d.py:27); net effect must equal F.conv2d.

Writes gpurun_out/cudnn_layernorm_probe.json (copied to tests/golden/ and checked by tests/test_oracle_golden.py).
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_ops as R  # noqa: E402


def cudnn_layer_norm(x, scale, bias, eps):
    import cudnn
    handle = cudnn.create_handle()
    _, B, T, C = x.shape
    x_stride = (B * T * C, 1, B * C, B)                        # layer_norm.py:10
    s_stride = (C, 1, C, 1)                                     # layer_norm.py:11 for a (1,1,1,C) tensor
    graph = cudnn.pygraph(intermediate_data_type=cudnn.data_type.FLOAT, compute_data_type=cudnn.data_type.FLOAT,
                          handle=handle)
    X = graph.tensor(name="X", dim=list(x.shape), stride=list(x_stride), data_type=cudnn.data_type.FLOAT)
    S = graph.tensor(name="scale", dim=[1, 1, 1, C], stride=list(s_stride), data_type=cudnn.data_type.FLOAT)
    Bt = graph.tensor(name="bias", dim=[1, 1, 1, C], stride=list(s_stride), data_type=cudnn.data_type.FLOAT)
    E = graph.tensor(name="epsilon", dim=[1, 1, 1, 1], stride=[1, 1, 1, 1], is_pass_by_value=True,
                     data_type=cudnn.data_type.FLOAT)
    Y, _, _ = graph.layernorm(name="layer_norm", norm_forward_phase=cudnn.norm_forward_phase.INFERENCE, input=X,
                              scale=S, bias=Bt, epsilon=E)
    Y.set_output(True).set_data_type(cudnn.data_type.FLOAT)
    graph.build([cudnn.heur_mode.A, cudnn.heur_mode.FALLBACK])
    y = torch.empty_like(x)
    ws = torch.empty(max(graph.get_workspace_size(), 1), dtype=torch.uint8, device=x.device)
    eps_cpu = torch.full((1, 1, 1, 1), eps, dtype=torch.float32)
    graph.execute({X: x, S: scale, Bt: bias, E: eps_cpu, Y: y}, ws, handle=handle)
    torch.cuda.synchronize()
    return y


def main():
    dev = torch.device("cuda:0")
    rec = {"cases": []}
    import cudnn
    rec["cudnn_backend_version"] = cudnn.backend_version()
    for B, T, C in ((1, 64, 320), (1, 16, 1280), (2, 64, 320), (2, 16, 1280), (4, 32, 640)):
        g = torch.Generator().manual_seed(B * 1000 + T + C)
        x = torch.randn(B, T, C, generator=g) * 2 + 0.3
        w, b = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
        case = {"B": B, "T": T, "C": C}
        try:
            y = cudnn_layer_norm(x.reshape(1, B, T, C).contiguous().to(dev), w.reshape(1, 1, 1, C).to(dev),
                                 b.reshape(1, 1, 1, C).to(dev), 1e-5).reshape(B, T, C).cpu()
            model = R.layer_norm(x, w, b, 1e-5, ln_strided=True)
            canon = R.layer_norm(x, w, b, 1e-5, ln_strided=False)
            case["cudnn"] = "executed"
            case["rel_err_vs_stride_model"] = ((y - model).abs().max() / model.abs().max()).item()
            case["rel_err_vs_canonical_layernorm"] = ((y - canon).abs().max() / canon.abs().max()).item()
        except Exception as exc:  # record what cuDNN says instead of hiding it
            case["cudnn"] = f"rejected: {type(exc).__name__}: {str(exc)[:160]}"
        rec["cases"].append(case)
        print(case)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rec, open("gpurun_out/cudnn_layernorm_probe.json", "w"), indent=1)


if __name__ == "__main__":
    main()
