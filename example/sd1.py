"""Text-to-image entry point of the reference (reference: example/sd1.py:23-79) on the B200 path:
prompt -> CLIP text encoder -> CFG / DDIM sampler loop (one captured CUDA graph replayed per step) -> VAE decode -> PNG.

    python -m example.sd1 --steps 50 --seed 42 --guidance 7.5 [--ckpt sd-v1-4.ckpt] [--bpe bpe_simple_vocab_16e6.txt.gz]
                          [--timing] [--no-graph] [--canonical] [--fp32] [--out rendered.png] [--latent-out latent.npy]

The reference downloads the checkpoint and the BPE merges file; there is no network here. Without --ckpt the three
models get seeded synthetic weights (same generators the parity tests use), without --bpe
(or TINYFUSERS_BPE_PATH) the prompt is mapped to deterministic stand-in token ids. The arithmetic path is the same.
"""
import argparse
import contextlib
import io
import os
import sys
import time
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def standin_token_ids(prompt):
    """Deterministic ids for a prompt when no merges file is available: one id per whitespace-separated word."""
    words = prompt.lower().split()[:75]
    ids = [zlib.crc32(w.encode()) % 49000 + 256 for w in words]
    return [49406] + ids + [49407] * (77 - len(ids) - 1)


def main():
    parser = argparse.ArgumentParser(description="Run Stable Diffusion 1.x on B200",
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument('--steps', type=int, default=5, help="Number of steps in diffusion")
    parser.add_argument('--prompt', type=str, default="a horse sized cat eating a bagel", help="Phrase to render")
    parser.add_argument('--out', type=str, default="rendered.png", help="Output filename")
    parser.add_argument('--noshow', action='store_true', help="Don't show the image (never shown here: headless)")
    parser.add_argument('--fp16', action='store_true', help="fp16 operands (always on: the B200 kernels are fp16 / fp32-accumulate)")
    parser.add_argument('--timing', action='store_true', help="Print timing per step")
    parser.add_argument('--seed', type=int, default=42, help="Set the random latent seed")
    parser.add_argument('--guidance', type=float, default=7.5, help="Prompt strength")
    parser.add_argument('--ckpt', type=str, default=None, help="sd-v1-x .ckpt (torch zip); default: seeded synthetic weights")
    parser.add_argument('--bpe', type=str, default=None, help="CLIP merges file (bpe_simple_vocab_16e6.txt.gz)")
    parser.add_argument('--size', type=int, default=512, help="image size (multiple of 64)")
    parser.add_argument('--no-graph', action='store_true', help="launch every step eagerly instead of replaying a CUDA graph")
    parser.add_argument('--canonical', action='store_true', help="canonical head merge (real checkpoints) instead of the reference's reshape")
    parser.add_argument('--latent-out', type=str, default=None, help="also save the final latent as .npy")
    parser.add_argument('--fp32', action='store_true', help="fp32 parity mode (the reference's dtype; plain-fp32 kernels, slow): tinyfusers_b200.set_precision('fp32')")
    args = parser.parse_args()

    import numpy as np
    import torch
    import tinyfusers_b200
    from tinyfusers_b200.storage.state import update_state
    from tinyfusers_b200.variants.sd import StableDiffusion

    if args.canonical:
        tinyfusers_b200.set_quirks(False)
    if args.fp32:
        tinyfusers_b200.set_precision("fp32")
    model = StableDiffusion()

    # load in weights (reference: sd1.py:38-41)
    if args.ckpt:
        from tinyfusers_b200.storage.unpicker import load_weights
        state = load_weights(args.ckpt)
        state = state.get('state_dict', state)
    else:
        from oracle import ref_ops as R   # synthetic weight generators only (no checkpoint offline)
        state = {}
        state.update(R.make_unet_state_dict(seed=1234))
        state.update(R.make_vae_decoder_state_dict())
        state.update(R.make_clip_state_dict())
    with contextlib.redirect_stdout(io.StringIO()) as skipped:
        update_state(model, state)
    print(f"weights loaded ({len(state)} tensors, {skipped.getvalue().count('skipped')} slots without a tensor)")

    # run through CLIP to get context (reference: sd1.py:43-50)
    bpe = args.bpe or os.environ.get("TINYFUSERS_BPE_PATH")
    if bpe:
        from tinyfusers_b200.tokenizer.clip import ClipTokenizer
        encode = ClipTokenizer(bpe).encode
    else:
        encode = standin_token_ids
    text_model = model.cond_stage_model.transformer.text_model
    context = text_model(np.array([encode(args.prompt)]))
    unconditional_context = text_model(np.array([encode("")]))
    print(f"CLIP context: {tuple(context.shape)}, unconditional CLIP context: {tuple(unconditional_context.shape)}")

    timesteps = list(range(1, 1000, 1000 // args.steps))
    print(f"running for {timesteps} timesteps")
    alphas = model.alphas_cumprod[timesteps]
    alphas_prev = torch.cat((torch.tensor([1.0], device=alphas.device), alphas[:-1])).float()

    hw = args.size // 8
    g = np.random.Generator(np.random.Philox(args.seed))
    latent = torch.from_numpy(g.standard_normal((1, 4, hw, hw), dtype=np.float32)).cuda()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if args.timing or args.no_graph:
        for index, timestep in list(enumerate(timesteps))[::-1]:
            ts = time.perf_counter()
            latent = model(unconditional_context, context, latent, torch.tensor([timestep]), alphas[[index]],
                           alphas_prev[[index]], torch.tensor([args.guidance]))
            if args.timing:
                torch.cuda.synchronize()
                print(f"{index:3d} {timestep:3d} step in {(time.perf_counter() - ts) * 1e3:.2f} ms")
    else:
        latent = model.sample(unconditional_context, context, latent, timesteps, alphas, alphas_prev, args.guidance)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    x = model.decode(latent)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{args.steps} steps in {(t1 - t0) * 1e3:.1f} ms ({args.steps / (t1 - t0):.1f} steps/s), decode {(t2 - t1) * 1e3:.1f} ms; "
          f"latent mean {latent.mean().item():.4f} std {latent.std().item():.4f}; image {tuple(x.shape)}")
    if args.latent_out:
        np.save(args.latent_out, latent.cpu().numpy())
    from PIL import Image
    im = Image.fromarray(x.cpu().numpy().astype(np.uint8, copy=False))
    print(f"saving {args.out}")
    im.save(args.out)


if __name__ == "__main__":
    main()
