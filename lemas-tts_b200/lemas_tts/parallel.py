"""Multi-GPU plumbing: the path shards by independent utterances — one process per GPU, a full
weight replica per rank, ONE broadcast of the packed weight blob from rank 0 at init, and no per-step collective.

Works on any torch.distributed backend (NCCL over NVLink on the box, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_utterances(lengths: list[int], world_size: int) -> list[list[int]]:
    """Length-balanced partition of utterance indices: longest first onto the currently lightest rank
    (cost model: N * (a + b*N), dense + attention FLOPs of SURVEY.md §8d).  Deterministic; every rank computes the
    same answer locally, so no communication is needed.  Within a rank the original order is kept."""
    def cost(n):
        return n * (378_888_192 + 90_112 * n)

    load = [0] * world_size
    shards: list[list[int]] = [[] for _ in range(world_size)]
    for i in sorted(range(len(lengths)), key=lambda i: (-lengths[i], i)):
        r = min(range(world_size), key=lambda r: (load[r], r))
        shards[r].append(i)
        load[r] += cost(lengths[i])
    return [sorted(s) for s in shards]


def broadcast_state_dict(sd: dict | None, src: int = 0, device="cpu", group=None) -> dict:
    """Rank `src` passes its state dict (CPU or device tensors), the others pass None; every rank returns the same
    dict on `device`.  One metadata exchange (key names / shapes / dtypes) + ONE tensor broadcast of a flat blob."""
    rank = dist.get_rank(group)
    meta = [None]
    if rank == src:
        meta[0] = [(k, tuple(v.shape), str(v.dtype).replace("torch.", "")) for k, v in sd.items()]
    dist.broadcast_object_list(meta, src=src, group=group)
    entries = meta[0]
    # everything travels as raw bytes so integer / fp16 / fp32 tensors share one blob
    sizes = [int(torch.empty(0, dtype=getattr(torch, dt)).element_size()) * int(torch.Size(shape).numel())
             for _, shape, dt in entries]
    offsets, total = [], 0
    for sz in sizes:
        offsets.append(total)
        total += (sz + 15) // 16 * 16
    blob = torch.empty(total, dtype=torch.uint8, device=device)
    if rank == src:
        for (k, shape, dt), off, sz in zip(entries, offsets, sizes):
            if sz:
                blob[off:off + sz] = sd[k].detach().contiguous().reshape(-1).view(torch.uint8).to(device)
    dist.broadcast(blob, src=src, group=group)
    out = {}
    for (k, shape, dt), off, sz in zip(entries, offsets, sizes):
        out[k] = blob[off:off + sz].view(getattr(torch, dt)).reshape(shape)
    return out


def max_over_ranks(value: float, device="cpu", group=None) -> float:
    """Step time of the job = the slowest rank's device time."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
