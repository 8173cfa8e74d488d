"""Generates tests/golden/reference_outputs_vae_clip.npz by running the REFERENCE'S OWN Python for the
"next" rows of SURVEY.md §8f: the VAE decoder (ResnetBlock, AttnBlock, Mid, Decoder, decode post-processing)
and the CLIP text encoder. Run here only:   python oracle/make_golden_vae.py

Uses the stand-ins of oracle/make_golden.py (numpy for CuPy, F.conv2d for the cuDNN conv graph, numpy LayerNorm
for the cuDNN layernorm graph, numpy row softmax for softmax.cu) plus two more, both documented in DESIGN.md:

  * `Embedding.__call__` (ff/embedding.py:15-23) cannot run in the reference: it allocates the one-hot matrix as
    (embed_sz, N) and indexes it with token ids up to 49407. The stand-in is the row lookup it stands for
    (weight[idx]), as SURVEY.md §8f rank 2 prescribes.
  * `StableDiffusion.decode` (variants/sd.py:48-54) hard-codes reshape(3, 512, 512). Its arithmetic is pinned by
    calling the reference method with the real post_quant_conv and a decoder stub that returns a fixed
    (1,3,512,512) field derived from its input; the Decoder itself is pinned separately at a 4x4 latent.
"""
import contextlib
import io
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle.make_golden import OUT, REF, install_shims  # noqa: E402


def main():
    os.chdir(REF)
    install_shims()
    from oracle import ref_ops as R
    import tinyfusers.ff.embedding as remb
    from tinyfusers.attention.attention import AttnBlock, CLIPAttention
    from tinyfusers.ff.nn import CLIPMLP
    from tinyfusers.storage.state import update_state
    from tinyfusers.vae.decoder import Decoder
    from tinyfusers.vae.encoder import CLIPTextTransformer
    from tinyfusers.vae.mid import Mid
    from tinyfusers.variants.sd import StableDiffusion
    from tinyfusers.vision.conv2d import Conv2d
    from tinyfusers.vision.resnet import ResnetBlock

    remb.Embedding.__call__ = lambda self, idx: np.asarray(self.weight)[np.asarray(idx).astype(np.int64)]

    G = {}

    def load(obj, sd, prefix):
        with contextlib.redirect_stdout(io.StringIO()):
            update_state(obj, sd, prefix)

    def rnd(seed, *shape, scale=1.0, shift=0.0):
        g = np.random.Generator(np.random.Philox(seed))
        return (g.standard_normal(shape, dtype=np.float32) * np.float32(scale) + np.float32(shift))

    # ---- ResnetBlock (vision/resnet.py:33-45), with and without nin_shortcut ----
    sd = {}
    R.add_resnet_block(sd, "rn", 64, 128, seed=701)
    R.add_resnet_block(sd, "rn2", 64, 64, seed=702)
    rn, rn2 = ResnetBlock(64, 128), ResnetBlock(64, 64)
    load(rn, sd, "rn")
    load(rn2, sd, "rn2")
    x = rnd(703, 1, 64, 8, 8, scale=1.3, shift=0.2)
    G["resnet_block_64_128"] = rn(x)
    G["resnet_block_64_64"] = rn2(x)

    # ---- AttnBlock (attention/attention.py:10-24): 4-D q/k/v into SDPA; non-square plane pins T=H, HS=W ----
    sd = {}
    R.add_attn_block(sd, "ab", 64, seed=711)
    ab = AttnBlock(64)
    load(ab, sd, "ab")
    G["attn_block_6x10"] = ab(rnd(712, 1, 64, 6, 10))
    G["attn_block_8x8"] = ab(rnd(713, 1, 64, 8, 8))

    # ---- Mid (vae/mid.py) ----
    sd = {}
    R.add_resnet_block(sd, "mid.block_1", 64, 64, seed=721)
    R.add_attn_block(sd, "mid.attn_1", 64, seed=721)
    R.add_resnet_block(sd, "mid.block_2", 64, 64, seed=721)
    mid = Mid(64)
    load(mid, sd, "mid")
    G["mid_64"] = mid(rnd(722, 1, 64, 8, 8))

    # ---- full Decoder at a 4x4 latent -> 32x32 (vae/decoder.py:22-34) ----
    vsd = R.make_vae_decoder_state_dict()
    dec = Decoder()
    load(dec, vsd, "first_stage_model.decoder")
    z = rnd(731, 1, 4, 4, 4)
    with contextlib.redirect_stdout(io.StringIO()):
        G["decoder_4x4"] = dec(z)

    # ---- decode post-processing (variants/sd.py:48-54) ----
    pq = Conv2d(4, 4, kernel_size=[1, 1])
    load(pq, vsd, "first_stage_model.post_quant_conv")
    yy, xx = np.meshgrid(np.linspace(-1.6, 1.6, 512, dtype=np.float32), np.linspace(-1.2, 1.2, 512, dtype=np.float32), indexing="ij")
    field = np.stack([yy * xx, yy + 0.3 * xx, np.sin(3 * yy) * np.cos(2 * xx)]).astype(np.float32)[None]

    def dec_stub(zq):
        G["decode_postquant_in"] = np.asarray(zq, dtype=np.float32)
        return field * np.float32(1.0 + 0.01 * float(np.asarray(zq).mean()))
    fake = types.SimpleNamespace(first_stage_model=types.SimpleNamespace(post_quant_conv=pq, decoder=dec_stub))
    zlat = rnd(741, 1, 4, 64, 64)
    G["decode_uint8"] = np.asarray(StableDiffusion.decode(fake, zlat)).astype(np.float32)
    G["decode_field_scale"] = np.float32(1.0 + 0.01 * float(G["decode_postquant_in"].mean())).reshape(1)

    # ---- CLIP text encoder (vae/encoder.py:36-81, attention.py:78-99, ff/nn.py:25-34) ----
    csd = R.make_clip_state_dict()
    P = "cond_stage_model.transformer.text_model"
    mlp = CLIPMLP()
    load(mlp, csd, P + ".encoder.layers.0.mlp")
    h = rnd(751, 1, 77, 768)
    G["clip_mlp"] = mlp(h)
    att = CLIPAttention()
    load(att, csd, P + ".encoder.layers.0.self_attn")
    mask = np.triu(np.full((1, 1, 77, 77), float("-inf")), k=1).astype(np.float32)
    G["clip_attention"] = att(h, mask)
    clip = CLIPTextTransformer()
    load(clip, csd, P)
    ids = np.array([[49406, 320, 1125, 539, 320, 2368, 6765, 525, 320, 11795] + [49407] * 67])
    G["clip_ids"] = ids.astype(np.float32)
    G["clip_text_transformer"] = clip(ids)

    np.savez_compressed(os.path.join(OUT, "reference_outputs_vae_clip.npz"),
                        **{k: np.asarray(v, dtype=np.float32) for k, v in G.items()})
    for k, v in G.items():
        print(f"{k:28s} {str(np.asarray(v).shape):20s} absmax {np.abs(np.asarray(v, dtype=np.float32)).max():.4f}")


if __name__ == "__main__":
    main()
