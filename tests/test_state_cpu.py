"""Host logic on CPU: update_state's key walk over the drop-in module tree, sampler schedule, sharding."""
import contextlib
import io

import numpy as np
import torch


def test_update_state_walks_reference_key_names(oracle):
    from tinyfusers_b200.storage.state import update_state
    from tinyfusers_b200.vision.resnet import ResBlock
    from tinyfusers_b200.attention.attention import SpatialTransformer
    sd = {}
    oracle.add_res_block(sd, "model.diffusion_model.input_blocks.4.0", 320, 640, seed=5)
    oracle.add_spatial_transformer(sd, "model.diffusion_model.input_blocks.4.1", 640, 768, seed=5)

    class Holder:
        def __init__(self):
            from collections import namedtuple
            blocks = [[], [], [], [], [ResBlock(320, 1280, 640), SpatialTransformer(640, 768, 8, 80)]]
            unet = type("U", (), {})()
            unet.input_blocks = blocks
            self.model = namedtuple("DiffusionModel", ["diffusion_model"])(diffusion_model=unet)

    h = Holder()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        update_state(h, sd)
    skipped = [l.split(": ")[1] for l in buf.getvalue().splitlines() if l.startswith("skipped")]
    assert sorted(skipped) == sorted(f"model.diffusion_model.input_blocks.4.1.transformer_blocks.0.{a}.{p}.bias"
                                     for a in ("attn1", "attn2") for p in ("to_q", "to_k", "to_v"))
    rb = h.model.diffusion_model.input_blocks[4][0]
    assert torch.equal(rb.in_layers[2].weight.cpu(), sd["model.diffusion_model.input_blocks.4.0.in_layers.2.weight"])
    st = h.model.diffusion_model.input_blocks[4][1]
    key = "model.diffusion_model.input_blocks.4.1.transformer_blocks.0.ff.net.0.proj.weight"
    assert torch.equal(st.transformer_blocks[0].ff.net[0].proj.weight.cpu(), sd[key])


def test_update_state_accepts_numpy_and_numpy_like(oracle):
    from tinyfusers_b200.ff.linear import Linear
    from tinyfusers_b200.storage.state import update_state

    class TG:  # tinygrad-style value: only .numpy()
        def __init__(self, a):
            self.a = a

        def numpy(self):
            return self.a

    lin = Linear(4, 3)
    w = np.arange(12, dtype=np.float32).reshape(3, 4)
    update_state(lin, {"l.weight": TG(w), "l.bias": np.ones(3, dtype=np.float64)}, "l")
    assert np.array_equal(lin.weight.cpu().numpy(), w) and lin.bias.dtype == torch.float32


def test_full_model_key_set_matches_oracle(oracle):
    from tinyfusers_b200.storage.state import update_state
    from tinyfusers_b200.variants.sd import StableDiffusion
    m = StableDiffusion()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        update_state(m, {})
    walked = {l.split(": ")[1] for l in buf.getvalue().splitlines() if l.startswith("skipped")}
    # keys of the synthetic generator without materialising 3.4 GB: walk the structure tables
    want = set()
    P = "model.diffusion_model"
    def res(p, cin, cout):
        for s in ("in_layers.0", "in_layers.2", "emb_layers.1", "out_layers.0", "out_layers.3"):
            want.update({f"{p}.{s}.weight", f"{p}.{s}.bias"})
        if cin != cout:
            want.update({f"{p}.skip_connection.weight", f"{p}.skip_connection.bias"})
    def st(p):
        for s in ("norm", "proj_in", "proj_out"):
            want.update({f"{p}.{s}.weight", f"{p}.{s}.bias"})
        t = p + ".transformer_blocks.0"
        for a in ("attn1", "attn2"):
            for q in ("to_q", "to_k", "to_v"):
                want.add(f"{t}.{a}.{q}.weight")
            want.update({f"{t}.{a}.to_out.0.weight", f"{t}.{a}.to_out.0.bias"})
        for s in ("ff.net.0.proj", "ff.net.2", "norm1", "norm2", "norm3"):
            want.update({f"{t}.{s}.weight", f"{t}.{s}.bias"})
    def layer(p, l):
        if l[0] == "conv": want.update({p + ".weight", p + ".bias"})
        elif l[0] == "res": res(p, l[1], l[2])
        elif l[0] == "st": st(p)
        elif l[0] == "down": want.update({p + ".op.weight", p + ".op.bias"})
        elif l[0] == "up": want.update({p + ".conv.weight", p + ".conv.bias"})
    for n in ("time_embed.0", "time_embed.2", "out.0", "out.2"):
        want.update({f"{P}.{n}.weight", f"{P}.{n}.bias"})
    for i, b in enumerate(oracle.UNET_INPUT_BLOCKS):
        for j, l in enumerate(b): layer(f"{P}.input_blocks.{i}.{j}", l)
    for j, l in enumerate(oracle.UNET_MIDDLE_BLOCK): layer(f"{P}.middle_block.{j}", l)
    for i, b in enumerate(oracle.UNET_OUTPUT_BLOCKS):
        for j, l in enumerate(b): layer(f"{P}.output_blocks.{i}.{j}", l)
    assert len(want) == 686
    nobias = {k for k in walked if k.endswith((".to_q.bias", ".to_k.bias", ".to_v.bias"))}
    assert {k for k in walked if k.startswith(P)} - nobias == want
    # the rows next to the hot path: every key of the synthetic VAE-decoder / CLIP dicts has a slot in the tree
    assert set(oracle.make_vae_decoder_state_dict().keys()) <= walked
    assert set(oracle.make_clip_state_dict().keys()) <= walked
    # ... and the tree has no VAE-decoder / CLIP slot the generators do not fill (encoder / quant_conv are not built)
    extra = {k for k in walked if k.startswith(("first_stage_model.decoder", "first_stage_model.post_quant_conv",
                                                "cond_stage_model"))}
    assert extra == set(oracle.make_vae_decoder_state_dict().keys()) | set(oracle.make_clip_state_dict().keys())


def test_sampler_schedule_matches_reference_loop(oracle):
    ts, alphas, alphas_prev = oracle.sampler_schedule(50)
    assert ts == list(range(1, 1000, 20)) and len(ts) == 50
    ac = oracle.get_alphas_cumprod()
    assert torch.equal(alphas, ac[ts]) and alphas_prev[0] == 1.0 and torch.equal(alphas_prev[1:], alphas[:-1])


def test_flop_count_matches_survey(oracle):
    assert abs(oracle.unet_step_flops(2, 64, 64) / 1e9 - 1606.5) < 0.1
    assert abs(oracle.unet_step_flops(8, 96, 96) / 1e9 - 17184.6) < 1.0


def test_reference_import_paths():
    """`tinyfusers.Tensor` (reference __init__.py:1) and `tinyfusers.tensor.tensor.Tensor` (what the model files import)."""
    import tinyfusers_b200
    from tinyfusers_b200.storage.tensor import Tensor as A
    from tinyfusers_b200.tensor.tensor import Tensor as B
    assert tinyfusers_b200.Tensor is A is B
    assert Tensor_sequential_ok(A)


def Tensor_sequential_ok(T):
    return T.sequential([lambda x: x + 1, lambda x: x * 2], 3) == 8
