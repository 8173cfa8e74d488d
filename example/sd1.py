"""Sampler loop of the reference's example (reference: example/sd1.py:23-79) on the B200 path.

The reference's script also downloads a checkpoint, tokenises a prompt, runs CLIP and decodes with the VAE;
none of that is on the denoising hot path (and none of it is available offline), so this entry point keeps the
flags and the loop structure (`:54-73`) and feeds synthetic prompt embeddings / seeded synthetic weights
(SURVEY.md §8d). Output: the final latent (saved as .npy with --out).

    python -m example.sd1 --steps 50 --seed 42 --guidance 7.5 [--timing] [--no-graph] [--canonical]
"""
import argparse
import contextlib
import io
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    parser = argparse.ArgumentParser(description="Run the Stable Diffusion 1.x denoising loop on B200",
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument('--steps', type=int, default=5, help="Number of steps in diffusion")
    parser.add_argument('--prompt', type=str, default="a horse sized cat eating a bagel", help="(unused: synthetic embeddings)")
    parser.add_argument('--noshow', action='store_true', help="Don't show the image")
    parser.add_argument('--fp16', action='store_true', help="fp16 operands (always on: the B200 kernels are fp16 / fp32-accumulate)")
    parser.add_argument('--timing', action='store_true', help="Print timing per step")
    parser.add_argument('--seed', type=int, default=42, help="Set the random latent seed")
    parser.add_argument('--guidance', type=float, default=7.5, help="Prompt strength")
    parser.add_argument('--no-graph', action='store_true', help="launch every step eagerly instead of replaying a CUDA graph")
    parser.add_argument('--canonical', action='store_true', help="canonical head merge (real checkpoints) instead of the reference's reshape")
    parser.add_argument('--out', type=str, default=None, help="save the final latent as .npy")
    args = parser.parse_args()

    import numpy as np
    import torch
    import tinyfusers_b200
    from oracle import ref_ops as R   # synthetic weights / inputs only (no checkpoint offline)
    from tinyfusers_b200.storage.state import update_state
    from tinyfusers_b200.variants.sd import StableDiffusion

    if args.canonical:
        tinyfusers_b200.set_quirks(False)
    model = StableDiffusion()
    with contextlib.redirect_stdout(io.StringIO()):
        update_state(model, R.make_unet_state_dict(seed=1234))
    latent, unconditional_context, context = R.make_inputs(1, 64, seed=args.seed, ctx_seed=args.seed + 1)
    print(f"CLIP context: {tuple(context.shape)}, unconditional CLIP context: {tuple(unconditional_context.shape)}")

    timesteps = list(range(1, 1000, 1000 // args.steps))
    print(f"running for {timesteps} timesteps")
    alphas = model.alphas_cumprod[timesteps]
    alphas_prev = torch.cat((torch.tensor([1.0], device=alphas.device), alphas[:-1])).float()

    latent, unconditional_context, context = latent.cuda(), unconditional_context.cuda(), context.cuda()
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
    dt = time.perf_counter() - t0
    print(f"{args.steps} steps in {dt * 1e3:.1f} ms ({args.steps / dt:.1f} steps/s); latent {tuple(latent.shape)} "
          f"mean {latent.mean().item():.4f} std {latent.std().item():.4f}")
    if args.out:
        np.save(args.out, latent.cpu().numpy())
        print(f"saving {args.out}")


if __name__ == "__main__":
    main()
