"""Data parallelism for the sampler: independent UNet replicas, one process per GPU (torchrun env).

The denoising path shards by image (and, inside a step, by the cond / uncond halves of CFG) with no
per-step exchange: every rank owns `shard_range(...)` of the images and runs its own captured sampler
graph. The only collective is one all-gather of the final latents (64 KiB per 512x512 image) over
NCCL / NVLink after the last step. The reference itself is single-GPU
(tinyfusers/storage/device.py:23)."""
import os

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_range(n_items, rank, world):
    """Contiguous, balanced split: the first (n_items % world) ranks take one extra item."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_latents(latent, group=None):
    """All-gather equally-shaped per-rank latents -> (world * B_local, 4, H, W) on every rank (rank order): ONE collective
    into one output tensor (64 KiB per 512x512 image over NVLink - latency-bound, so no list of per-rank buffers)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return latent
    world = dist.get_world_size(group)
    src = latent.contiguous()
    out = torch.empty((world * src.shape[0],) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    try:
        dist.all_gather_into_tensor(out, src, group=group)
    except (RuntimeError, NotImplementedError):   # backend without the fused form
        dist.all_gather(list(out.chunk(world, dim=0)), src, group=group)
    return out


def gather_ragged_latents(latent, n_total, group=None):
    """Same for ragged shards (n_total not divisible by world): pads to the largest shard, gathers, trims."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return latent
    world = dist.get_world_size(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    biggest = max(e - s for s, e in sizes)
    pad = torch.zeros((biggest,) + tuple(latent.shape[1:]), dtype=latent.dtype, device=latent.device)
    pad[:latent.shape[0]] = latent
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:e - s] for p, (s, e) in zip(parts, sizes)], dim=0)


def sample_sharded(model, unconditional_context, context, latent, timesteps, alphas, alphas_prev, guidance, group=None,
                   sample_fn=None):
    """Data-parallel `StableDiffusion.sample`: every rank holds the same (B, ...) inputs, denoises its own
    `shard_range(B, rank, world)` of the images on its own UNet replica (its own captured graph, no per-step
    communication) and all ranks return all B final latents (one all-gather). `sample_fn(unc, ctx, lat, ...)` defaults to
    `model.sample`; the CPU tests pass a stand-in."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        fn = sample_fn or model.sample
        return fn(unconditional_context, context, latent, timesteps, alphas, alphas_prev, guidance)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    B = latent.shape[0]
    s, e = shard_range(B, rank, world)
    fn = sample_fn or model.sample
    if e > s:
        mine = fn(unconditional_context[s:e], context[s:e], latent[s:e], timesteps, alphas, alphas_prev, guidance)
    else:
        mine = latent[:0].clone()
    if B % world == 0:
        return gather_latents(mine, group)
    return gather_ragged_latents(mine, B, group)
