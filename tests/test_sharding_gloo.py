"""CPU only, world_size 2 over gloo: the multi-rank host logic (frame / row-band sharding, framebuffer gather and
reassembly).  The per-rank "renderer" here is a deterministic stand-in for the kernel -- the kernels themselves are
covered by the -m gpu tests; what is checked is that shards tile the work exactly once and the gather reassembles it."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ray_tracing_octrees_b200 import sharding


def _fake_plane(frame, y0, y1, width):
    yy, xx = np.meshgrid(np.arange(y0, y1), np.arange(width), indexing="ij")
    return (frame * 1_000_003 + yy * 4099 + xx).astype(np.int32)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        H, W, F = 37, 24, 7
        # (1) frames sharded over ranks, gathered to rank 0
        mine = sharding.shard_frames(F, world, rank)
        pad = (F + world - 1) // world
        local = np.zeros((pad, H, W), np.int32)
        for i, k in enumerate(mine):
            local[i] = _fake_plane(k, 0, H, W)
        got = sharding.gather_planes(torch.from_numpy(local), dst=0)
        ok1 = True
        if rank == 0:
            frames = sharding.interleave_frames([g.numpy() for g in got], F, world)
            ok1 = all(np.array_equal(frames[k], _fake_plane(k, 0, H, W)) for k in range(F))
        else:
            ok1 = got is None
        # (2) one frame sharded by interleaved row bands
        bands = sharding.shard_rows(H, world, rank, band=8)
        rows = max(sum(y1 - y0 for y0, y1 in sharding.shard_rows(H, world, r, band=8)) for r in range(world))
        local = np.zeros((rows, W), np.int32)
        off = 0
        for (y0, y1) in bands:
            local[off:off + y1 - y0] = _fake_plane(3, y0, y1, W); off += y1 - y0
        got = sharding.gather_planes(torch.from_numpy(local), dst=0)
        ok2 = True
        if rank == 0:
            full = sharding.assemble_rows(H, W, world, [g.numpy() for g in got], band=8)
            ok2 = np.array_equal(full, _fake_plane(3, 0, H, W))
        # (3) max-over-ranks timing reduction used by bench.py
        t = torch.tensor([1.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok3 = float(t) == float(world)
        q.put((rank, bool(ok1), bool(ok2), bool(ok3)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True, True, True), (1, True, True, True)]


def _subgroup_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ranks 1 and 2 form a group of their own; its rank 0 is GLOBAL rank 1: gather_planes takes the group's rank and hands torch the global one
        group = dist.new_group(ranks=[1, 2])
        ok = True
        if rank in (1, 2):
            got = sharding.gather_planes(torch.full((4, 5), rank, dtype=torch.int32), dst=0, group=group)
            if rank == 1:
                ok = got is not None and len(got) == 2 and int(got[0][0, 0]) == 1 and int(got[1][0, 0]) == 2
            else:
                ok = got is None
        dist.barrier()
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_gather_inside_a_subgroup_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_subgroup_worker, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True), (2, True)]


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_shards_tile_exactly_once(world):
    for H in (1, 7, 8, 9, 135, 1080, 2160):
        seen = np.zeros(H, np.int32)
        for r in range(world):
            for (y0, y1) in sharding.shard_rows(H, world, r):
                seen[y0:y1] += 1
        assert (seen == 1).all()
    for F in (0, 1, 5, 64):
        allf = sorted(k for r in range(world) for k in sharding.shard_frames(F, world, r))
        assert allf == list(range(F))
        sizes = [len(sharding.shard_frames(F, world, r)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


def test_weighted_row_ranges_tile_the_batch_exactly_once():
    """The pure host logic of the gathered renderer (sharding.py; csrc/rto_group.cu does the same in C++): weighted contiguous ranges
    of 8-row tiles, each cut into at most three launches, cover every row of every frame once, on tile boundaries."""
    rng = np.random.default_rng(5)
    for H in (1080, 2160, 203, 8, 1):
        tpf = (H + 7) // 8
        for n, w in ((16, [0.5, 1, 1, 1]), (1, [1, 1]), (3, [0.02, 1, 5]), (128, [0.55] + [1] * 7), (5, list(rng.uniform(0.05, 3, 8)))):
            cuts = sharding.deal_tiles(tpf * n, w)
            assert cuts[0] == 0 and cuts[-1] == tpf * n and all(a <= b for a, b in zip(cuts, cuts[1:]))
            rows = [sharding.tile_row_to_row(c, H) for c in cuts]
            assert rows[-1] == n * H
            done = {}
            for r in range(len(w)):
                parts = sharding.split_rows(rows[r], rows[r + 1], H)
                assert len(parts) <= 3
                for (f, k, y0, y1) in parts:
                    assert y0 % 8 == 0 and (y1 % 8 == 0 or y1 == H) and 0 <= y0 < y1 <= H and (k == 1 or (y0 == 0 and y1 == H))
                    for ff in range(f, f + k):
                        assert done.get(ff, 0) == y0, "gap or overlap in frame %d" % ff
                        done[ff] = y1
            assert len(done) == n and all(v == H for v in done.values())


def test_rebalance_moves_work_away_from_slow_ranks():
    w = [1.0, 1.0, 1.0, 1.0]
    speed = [0.5, 1.0, 1.0, 1.25]                       # rank 0 is half as fast (it also expands everyone's codes)
    for _ in range(12):
        t = [wi / sum(w) / si for wi, si in zip(w, speed)]
        w = sharding.rebalance(w, t)
    t = [wi / sum(w) / si for wi, si in zip(w, speed)]
    assert max(t) / min(t) < 1.02 and w[0] < w[1] < w[3]
