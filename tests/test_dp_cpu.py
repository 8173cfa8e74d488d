"""N > 1 path on CPU: world_size-2 gloo processes exercise the sharding + final-latent gather logic
(tinyfusers_b200/dp.py) that bench.py / example use with NCCL on the GPU box."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_images, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from tinyfusers_b200 import dp
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s, e = dp.shard_range(n_images, rank, world)
    # each rank "denoises" its own images: latent value encodes the global image index
    mine = torch.stack([torch.full((4, 8, 8), float(i)) for i in range(s, e)]) if e > s else torch.zeros(0, 4, 8, 8)
    if n_images % world == 0:
        full = dp.gather_latents(mine)
    else:
        full = dp.gather_ragged_latents(mine, n_images)
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([10.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    q.put((rank, full[:, 0, 0, 0].tolist(), float(t.item()), (s, e)))
    dist.destroy_process_group()


def _run(n_images, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_images, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(out)


def test_even_shards_gather_in_rank_order():
    out = _run(8)
    for rank, vals, tmax, (s, e) in out:
        assert vals == [float(i) for i in range(8)]
        assert tmax == 11.0
    assert [o[3] for o in out] == [(0, 4), (4, 8)]


def test_ragged_shards():
    out = _run(5)
    for rank, vals, tmax, _ in out:
        assert vals == [float(i) for i in range(5)]
    assert [o[3] for o in out] == [(0, 3), (3, 5)]


def test_shard_range_covers_everything():
    from tinyfusers_b200.dp import shard_range
    for n in (1, 7, 8, 64):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _sample_worker(rank, world, port, n_images, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from tinyfusers_b200 import dp
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)          # every rank holds the same full batch
    lat = torch.randn(n_images, 4, 8, 8, generator=g)
    unc, ctx = torch.randn(n_images, 77, 16, generator=g), torch.randn(n_images, 77, 16, generator=g)
    calls = []

    def fake_sample(u, c, x, ts, a, ap, guidance):   # stands in for StableDiffusion.sample: per-image, order-preserving
        calls.append(x.shape[0])
        return x * guidance + u.mean(dim=(1, 2)).reshape(-1, 1, 1, 1) - c.mean(dim=(1, 2)).reshape(-1, 1, 1, 1)
    out = dp.sample_sharded(None, unc, ctx, lat, [1, 2], None, None, 7.5, sample_fn=fake_sample)
    ref = fake_sample(unc, ctx, lat, None, None, None, 7.5)
    q.put((rank, bool(torch.equal(out, ref)), calls[0]))
    dist.destroy_process_group()


def _run_sample(n_images, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sample_worker, args=(r, world, port, n_images, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(out)


def test_sample_sharded_returns_all_images_on_every_rank():
    """StableDiffusion.sample_dp's host logic (dp.sample_sharded): each rank runs only its shard, every rank gets all latents."""
    for n in (8, 5):
        out = _run_sample(n)
        assert all(ok for _, ok, _ in out)
        assert sorted(c for _, _, c in out) == sorted([n // 2, n - n // 2])
