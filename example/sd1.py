"""Text-to-image entry point of the reference (reference: example/sd1.py:23-79) on the B200 path:
prompt -> CLIP text encoder -> CFG / DDIM sampler loop (one captured CUDA graph replayed per step) -> VAE decode -> PNG.

    python -m example.sd1 --steps 50 --seed 42 --guidance 7.5 [--ckpt sd-v1-4.ckpt] [--bpe bpe_simple_vocab_16e6.txt.gz]
                          [--timing] [--no-graph] [--canonical] [--fp32] [--out rendered.png] [--latent-out latent.npy]
                          [--batch B] [--gpus N]

--batch B renders B images of the prompt (seeds seed .. seed+B-1); --gpus N shards them over N GPUs of this node: the
script re-launches itself under torch.distributed.run (one process, one UNet replica and one captured sampler graph per GPU),
each rank denoises its share and the final latents are all-gathered over NCCL (tinyfusers_b200/dp.py); rank 0 decodes and saves.

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
    parser.add_argument('--batch', type=int, default=1, help="images to render (seeds seed .. seed+batch-1)")
    parser.add_argument('--gpus', type=int, default=1, help="GPUs of this node to shard the batch over (re-launches under torch.distributed.run)")
    parser.add_argument('--fp32', action='store_true', help="fp32 parity mode (the reference's dtype; plain-fp32 kernels, slow): tinyfusers_b200.set_precision('fp32')")
    args = parser.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # one process per GPU: re-launch this module under torchrun with the same arguments
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29533"), "-m", "example.sd1"] + sys.argv[1:]
        sys.exit(subprocess.call(cmd, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    rank, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))

    import numpy as np
    import torch
    import tinyfusers_b200
    from tinyfusers_b200.storage.state import update_state
    from tinyfusers_b200.variants.sd import StableDiffusion

    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    say = print if rank == 0 else (lambda *a, **k: None)
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
        from tinyfusers_b200 import synthetic   # seeded synthetic weights (no checkpoint offline)
        state = {}
        state.update(synthetic.make_unet_state_dict(seed=1234))
        state.update(synthetic.make_vae_decoder_state_dict())
        state.update(synthetic.make_clip_state_dict())
    with contextlib.redirect_stdout(io.StringIO()) as skipped:
        update_state(model, state)
    say(f"weights loaded ({len(state)} tensors, {skipped.getvalue().count('skipped')} slots without a tensor)")

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
    say(f"CLIP context: {tuple(context.shape)}, unconditional CLIP context: {tuple(unconditional_context.shape)}")

    timesteps = list(range(1, 1000, 1000 // args.steps))
    say(f"running for {timesteps} timesteps")
    alphas = model.alphas_cumprod[timesteps]
    alphas_prev = torch.cat((torch.tensor([1.0], device=alphas.device), alphas[:-1])).float()

    hw = args.size // 8
    B = args.batch
    lats = []
    for i in range(B):   # image i = the single-image run with seed + i
        g = np.random.Generator(np.random.Philox(args.seed + i))
        lats.append(torch.from_numpy(g.standard_normal((1, 4, hw, hw), dtype=np.float32)))
    latent = torch.cat(lats).cuda()
    if B > 1 or world > 1:
        context = context.expand(B, -1, -1).contiguous()
        unconditional_context = unconditional_context.expand(B, -1, -1).contiguous()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if world > 1:
        latent = model.sample_dp(unconditional_context, context, latent, timesteps, alphas, alphas_prev, args.guidance)
    elif args.timing or args.no_graph:
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
    if rank != 0:
        torch.distributed.destroy_process_group()
        return
    x = model.decode(latent)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{B} image(s) x {args.steps} steps on {world} GPU(s) in {(t1 - t0) * 1e3:.1f} ms ({B * args.steps / (t1 - t0):.1f} image-steps/s), "
          f"decode {(t2 - t1) * 1e3:.1f} ms; latent mean {latent.mean().item():.4f} std {latent.std().item():.4f}; image {tuple(x.shape)}")
    if args.latent_out:
        np.save(args.latent_out, latent.cpu().numpy())
    from PIL import Image
    imgs = x.cpu().numpy().astype(np.uint8, copy=False)
    imgs = imgs[None] if imgs.ndim == 3 else imgs
    for i, im in enumerate(imgs):
        name = args.out if len(imgs) == 1 else "{}_{}{}".format(os.path.splitext(args.out)[0], i, os.path.splitext(args.out)[1])
        print(f"saving {name}")
        Image.fromarray(im).save(name)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
