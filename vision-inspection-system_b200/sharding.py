"""Image sharding across the GPUs of one box (one process per GPU, torch.distributed for plumbing only).

Every frame is independent, so the hot path has NO collective: each rank preprocesses its own shard and keeps its
``pixel_values`` in its own HBM.  ``gather_patches`` is the optional exchange step for a consumer that lives on one
rank (variable-length gather over NCCL send/recv; works on gloo/CPU tensors too, which is how it is tested).

* uniform batches      -> ``contiguous_shard``: equal contiguous chunks, order preserved
* mixed resolutions    -> ``balanced_shards``: greedy largest-first on algorithmic bytes (BASELINE config 5)
"""
from __future__ import annotations

import heapq

import torch

from . import geometry as G


def contiguous_shard(n_items: int, rank: int, world: int) -> range:
    """Items [lo, hi) of rank ``rank`` when ``n_items`` are split into ``world`` near-equal contiguous chunks."""
    if not (0 <= rank < world):
        raise ValueError("rank outside world")
    return range(n_items * rank // world, n_items * (rank + 1) // world)


def frame_bytes(h: int, w: int, min_pixels: int = G.DEFAULT_MIN_PIXELS, max_pixels: int = G.DEFAULT_MAX_PIXELS) -> int:
    """Algorithmic HBM bytes of one frame -> pixel_values: uint8 read + fp32 write."""
    dh, dw = G.smart_resize(h, w, G.FACTOR, min_pixels, max_pixels)
    return h * w * 3 + (dh // G.PATCH_SIZE) * (dw // G.PATCH_SIZE) * G.ROW_FLOATS * 4


def balanced_shards(costs, world: int) -> list:
    """Greedy largest-first assignment of items to ``world`` ranks; returns a list of sorted index lists.
    Deterministic (ties broken by index / rank), so every rank computes the same plan without communicating."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    heap = [(0, r) for r in range(world)]
    heapq.heapify(heap)
    shards = [[] for _ in range(world)]
    for i in order:
        load, r = heapq.heappop(heap)
        shards[r].append(i)
        heapq.heappush(heap, (load + costs[i], r))
    return [sorted(s) for s in shards]


def gather_patches(pixel_values: torch.Tensor, image_grid_thw: torch.Tensor, frame_ids, dst: int = 0, group=None):
    """Optional exchange step: collect every rank's patch rows on rank ``dst`` in original frame order.

    ``pixel_values`` [n_local_rows, 1176] (device of the backend), ``image_grid_thw`` [n_local_frames, 3] int64,
    ``frame_ids``: global index of each local frame.  Returns (pixel_values, image_grid_thw) on ``dst`` and
    (None, None) elsewhere.  Receiver ingress bounds this step (~770 GB/s measured peer copy on NVLink 5), which is
    why it is not part of the headline metric.
    """
    import torch.distributed as dist
    # `dst` and the ranks below are GROUP ranks; P2POp peers are GLOBAL ranks: translate (identity for the default group)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    peer = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
    dev = pixel_values.device
    ids = torch.as_tensor(list(frame_ids), dtype=torch.int64)
    grid = image_grid_thw.to(torch.int64).cpu()
    if ids.numel() != grid.shape[0]:
        raise ValueError("one frame id per grid row expected")
    meta = torch.tensor([pixel_values.shape[0], ids.numel()], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    metas = [m.cpu().tolist() for m in metas]
    payload = torch.cat([ids, grid.reshape(-1)]).to(dev)            # ids then grid, one small message per rank
    # one batched group of point-to-point operations: the peers' transfers run concurrently into the receiver
    # (unbatched send/recv pairs on the default group are serialised one after the other)
    if rank != dst:
        if ids.numel():
            ops = [dist.P2POp(dist.isend, payload, peer(dst), group),
                   dist.P2POp(dist.isend, pixel_values.contiguous(), peer(dst), group)]
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return None, None
    parts, ops, pending = [], [], []
    for r in range(world):
        rows, frames = metas[r]
        if frames == 0:
            continue
        if r == rank:
            parts.append((ids, grid, pixel_values))
            continue
        buf = torch.empty(frames * 4, dtype=torch.int64, device=dev)
        pv = torch.empty((rows, pixel_values.shape[1]), dtype=pixel_values.dtype, device=dev)
        ops += [dist.P2POp(dist.irecv, buf, peer(r), group), dist.P2POp(dist.irecv, pv, peer(r), group)]
        pending.append((frames, buf, pv))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for frames, buf, pv in pending:
        buf = buf.cpu()
        parts.append((buf[:frames], buf[frames:].reshape(frames, 3), pv))
    # restore original frame order: rows of a frame are contiguous inside its rank's tensor
    entries = []
    for ids_r, grid_r, pv_r in parts:
        counts = (grid_r[:, 0] * grid_r[:, 1] * grid_r[:, 2]).tolist()
        at = 0
        for k, n in enumerate(counts):
            entries.append((int(ids_r[k]), grid_r[k], pv_r[at:at + n]))
            at += n
    entries.sort(key=lambda e: e[0])
    if not entries:                                  # no rank had a frame
        return pixel_values.new_zeros((0, pixel_values.shape[1])), torch.zeros((0, 3), dtype=torch.int64)
    return torch.cat([e[2] for e in entries]), torch.stack([e[1] for e in entries])
