"""Multi-GPU host logic on CPU: shard plans, and the optional gather over a world_size-2 gloo group."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vision_inspection_system_b200 import sharding as S
from vision_inspection_system_b200 import synth


def test_contiguous_shards_cover_everything():
    for n in (0, 1, 7, 256, 1000):
        for world in (1, 2, 4, 8):
            got = [i for r in range(world) for i in S.contiguous_shard(n, r, world)]
            assert got == list(range(n))
            sizes = [len(S.contiguous_shard(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        S.contiguous_shard(4, 2, 2)


def test_balanced_shards_on_config5_mix():
    shapes = synth.mixed_resolution_shapes(8192, seed=9000)
    costs = [S.frame_bytes(h, w) for h, w in shapes]
    assert S.frame_bytes(1080, 1920) == 29213952
    assert S.frame_bytes(2160, 3840) == 47876352
    assert S.frame_bytes(1080, 1920, max_pixels=12845056) == 56854656
    shards = S.balanced_shards(costs, 8)
    assert sorted(i for s in shards for i in s) == list(range(8192))
    loads = [sum(costs[i] for i in s) for s in shards]
    assert (max(loads) - min(loads)) / max(loads) < 0.002           # greedy largest-first is tight on 8192 items
    assert shards == S.balanced_shards(costs, 8)                    # deterministic: no communication needed


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        grids = [(1, int(a), int(b)) for a, b in rng.integers(1, 5, (7, 2)) * 2]
        rows = [g[1] * g[2] for g in grids]
        full = torch.arange(sum(rows) * 1176, dtype=torch.float32).reshape(-1, 1176)
        starts = np.concatenate([[0], np.cumsum(rows)])
        mine = S.balanced_shards(rows, world)[rank]
        pv = torch.cat([full[starts[i]:starts[i + 1]] for i in mine]) if mine else full[:0]
        grid = torch.tensor([grids[i] for i in mine], dtype=torch.int64).reshape(-1, 3)
        out, ogrid = S.gather_patches(pv, grid, mine, dst=0)
        if rank == 0:
            q.put((torch.equal(out, full), ogrid.tolist() == [list(g) for g in grids]))
        else:
            q.put((out is None, ogrid is None))
    finally:
        dist.destroy_process_group()


def _worker_subgroup(rank, world, port, q):
    """world 3, gather inside the sub-group {1, 2}: group rank 0 is GLOBAL rank 1 (P2POp peers are global)."""
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        group = dist.new_group([1, 2])
        ok = True
        if rank in (1, 2):
            grids = [(1, 2, 4), (1, 4, 2), (1, 2, 2)]
            rows = [g[1] * g[2] for g in grids]
            full = torch.arange(sum(rows) * 1176, dtype=torch.float32).reshape(-1, 1176)
            starts = np.concatenate([[0], np.cumsum(rows)])
            mine = [0, 2] if rank == 1 else [1]
            pv = torch.cat([full[starts[i]:starts[i + 1]] for i in mine])
            grid = torch.tensor([grids[i] for i in mine], dtype=torch.int64).reshape(-1, 3)
            out, ogrid = S.gather_patches(pv, grid, mine, dst=0, group=group)
            ok = (torch.equal(out, full) and ogrid.tolist() == [list(g) for g in grids]) if rank == 1 else out is None
            # nobody has a frame: empty tensors on dst, not an exception
            out, ogrid = S.gather_patches(full[:0], torch.zeros((0, 3), dtype=torch.int64), [], dst=0, group=group)
            ok = ok and ((out.shape == (0, 1176) and ogrid.shape == (0, 3)) if rank == 1 else out is None)
        q.put((ok, True))
    finally:
        dist.destroy_process_group()


def test_gather_patches_subgroup_world3():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker_subgroup, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(a and b for a, b in results)


def test_gather_patches_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(a and b for a, b in results)
