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
    """All-gather equally-shaped per-rank latents -> (world * B_local, 4, H, W) on every rank (rank order)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return latent
    world = dist.get_world_size(group)
    parts = [torch.empty_like(latent) for _ in range(world)]
    dist.all_gather(parts, latent.contiguous(), group=group)
    return torch.cat(parts, dim=0)


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
