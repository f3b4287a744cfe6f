"""World-size-2 gloo test (CPU) of the tiled multi-GPU path's host logic: round-robin tile
ownership, padded all-gather, re-assembly in global tile order and reference-order blending must
reproduce the single-process tiled model exactly (bit-for-bit), for an odd tile count."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffbir_b200.sampler.sampler import assemble_units, tile_slots, tiles_of_rank
from diffbir_b200.utils.common import gaussian_weights, sliding_windows


def _stub(x, t, c_img):
    return 0.3 * torch.tanh(x) + 0.05 * c_img + 1e-4 * float(t)


def _reference_tiled(x, c_img, t, size, stride):
    out, cnt = torch.zeros_like(x), torch.zeros_like(x)
    w = torch.tensor(gaussian_weights(size, size)[None, None], dtype=x.dtype)
    for a, b, c, d in sliding_windows(x.shape[2], x.shape[3], size, stride):
        out[..., a:b, c:d] += _stub(x[..., a:b, c:d], t, c_img[..., a:b, c:d]) * w
        cnt[..., a:b, c:d] += w
    return out / cnt


def _worker(rank, world, port, q, H=24, W=40):
    """Mirrors EngineEval's tiled sharding: unit u = branch * T + tile, owner u % world, padded all-gather,
    assemble_units, per-branch blend in the reference's row-major accumulation order."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)               # same data on every rank
    size, stride, B, C, nbr = 16, 8, 1, 4, 2
    x = torch.randn(B, C, H, W, generator=g)
    c_imgs = [torch.randn(B, C, H, W, generator=g) for _ in range(nbr)]
    wins = sliding_windows(H, W, size, stride)
    T = len(wins)
    U = nbr * T
    mine, slots = tiles_of_rank(U, rank, world), tile_slots(U, world)
    send = torch.zeros(slots, B, C, size, size)
    for s, u in enumerate(mine):
        j, t = divmod(u, T)
        a, b, c, d = wins[t]
        send[s] = _stub(x[..., a:b, c:d], 7 + j, c_imgs[j][..., a:b, c:d])
    parts = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(parts, send)                        # gloo: list form (NCCL path uses _into_tensor)
    tiles = assemble_units(torch.stack(parts, 0), U).view(nbr, T, B, C, size, size)
    w = torch.tensor(gaussian_weights(size, size)[None, None], dtype=x.dtype)
    ok = True
    for j in range(nbr):
        out, cnt = torch.zeros_like(x), torch.zeros_like(x)
        for t, (a, b, c, d) in enumerate(wins):        # reference accumulation order
            out[..., a:b, c:d] += tiles[j, t] * w
            cnt[..., a:b, c:d] += w
        ok = ok and bool(torch.equal(out / cnt, _reference_tiled(x, c_imgs[j], 7 + j, size, stride)))
    q.put((rank, ok, T, len(mine)))
    dist.destroy_process_group()


def _run_world(world, H, W):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, H, W)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _, _ in results), results
    return results


def test_two_rank_tile_sharding_matches_single_process():
    results = _run_world(2, 24, 40)
    counts = sorted(n for _, _, _, n in results)
    assert sum(counts) == 2 * results[0][2] and counts[1] - counts[0] <= 1


def test_four_rank_tile_sharding_matches_single_process():
    """The N = 4 point of the driver's scaling run (only N = 2 and N = 8 ran on GPUs during development)."""
    results = _run_world(4, 24, 40)
    counts = sorted(n for _, _, _, n in results)
    assert sum(counts) == 2 * results[0][2] and counts[-1] - counts[0] <= 1


def test_fewer_tiles_than_ranks():
    """T < world (e.g. a 768^2 image on 8 GPUs): ranks without a tile send a zero buffer, still join
    the all-gather, and every rank blends the same result (used to dead-lock)."""
    results = _run_world(5, 16, 24)                    # 2 tiles x 2 branches = 4 units on 5 ranks
    assert results[0][2] == 2
    assert sorted(n for _, _, _, n in results) == [0, 1, 1, 1, 1]


def test_ownership_covers_every_tile_once():
    for T in (1, 2, 4, 7, 49):
        for world in (1, 2, 4, 8):
            owned = sorted(t for r in range(world) for t in tiles_of_rank(T, r, world))
            assert owned == list(range(T))
            assert max(len(tiles_of_rank(T, r, world)) for r in range(world)) == tile_slots(T, world) >= 1
    # 49 whole tiles over 8 ranks: 7,6,..,6 tiles = 14 / 12 forwards -> 87.5 % ideal efficiency; the 98
    # (tile, CFG branch) units the sampler shards instead: 13,13,12,..,12 -> 94 %
    assert [len(tiles_of_rank(49, r, 8)) for r in range(8)] == [7] + [6] * 7
    assert [len(tiles_of_rank(98, r, 8)) for r in range(8)] == [13, 13] + [12] * 6


# ---- batch sharding (BASELINE configs[4]: images x CFG branches over the ranks) -----------------
def _unit_worker(rank, world, port, q, B):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(1)
    nbr, C, H, W = 2, 4, 8, 8
    x = torch.randn(B, C, H, W, generator=g)
    c_img = torch.randn(nbr, B, C, H, W, generator=g)
    U = nbr * B
    mine, slots = tiles_of_rank(U, rank, world), tile_slots(U, world)
    send = torch.zeros(slots, C, H, W)
    for s, u in enumerate(mine):
        j, b = divmod(u, B)
        send[s] = _stub(x[b], 3 + j, c_img[j, b])
    parts = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(parts, send)
    units = assemble_units(torch.stack(parts, 0), U).view(nbr, B, C, H, W)
    ref = torch.stack([torch.stack([_stub(x[b], 3 + j, c_img[j, b]) for b in range(B)]) for j in range(nbr)])
    q.put((rank, bool(torch.equal(units, ref)), U, len(mine)))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,B", [(2, 4), (3, 4), (3, 1), (4, 4), (8, 4)])     # (4, 4) / (8, 4): configs[4] at N = 4 / 8
def test_batch_unit_sharding_matches_single_process(world, B):
    """(CFG branch, image) units round-robin over the ranks + padded all-gather == the un-sharded batch,
    including U not divisible by world and U < world (ranks without a unit still join the collective)."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_unit_worker, args=(r, world, port, q, B)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _, _ in results), results
    assert sum(n for _, _, _, n in results) == 2 * B
