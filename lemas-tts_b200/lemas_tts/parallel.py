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


# ------------------------------------------------------------------------------------------- sharded synthesis
#
# BASELINE.json configs[3] / SURVEY.md §8e: a LIST of independent utterances is dealt to the ranks by
# `shard_utterances`, every rank runs CFM.sample + Vocos.decode on its own utterances (the reference's B > 1 path,
# cfm.py:336-339), and the waveforms are gathered on the host.  No collective touches the data path.
#
# Sharding must not change a single bit of any utterance.  The reference's batch padding is
# "leaky" (an utterance padded inside a longer batch differs from its solo run), so utterances are only
# ever batched with utterances of the SAME shape (reference frames, total frames): such a batch has no padding, every
# kernel tiles per sequence, and a row's result does not depend on which or how many other rows share the launch.
# The noise of utterance i is drawn from its own generator seeded with (seed, i), independent of rank and batch.


def utterance_noise(seed: int, index: int, frames: int, mel_dim: int, device) -> torch.Tensor:
    """y0 of utterance `index` (cfm.py:430-435 draws randn(duration, mel) per row): a per-utterance device generator,
    so the draw does not depend on how the list is sharded or batched."""
    g = torch.Generator(device=device)
    g.manual_seed((int(seed) * 1_000_003 + int(index)) & 0x7FFFFFFFFFFFFFFF)
    return torch.randn(frames, mel_dim, generator=g, device=device, dtype=torch.float32)


def make_synth_fn(model, vocoder, *, steps: int = 32, cfg_strength: float = 2.0, sway_sampling_coef=3.0, seed: int = 0,
                  batch_size: int = 32, use_acc_grl: bool = False):
    """Binds a CFM model and a vocoder into the callable `synthesize_sharded` drives:
    fn([(global_index, utterance), ...]) -> {global_index: waveform (1-D fp32, CPU)}.
    An utterance is a dict: cond = reference mel [Tc, mel] (fp32), text = token ids [nt] (int64, no padding),
    duration = total frames N."""
    device = model.device
    mel_dim = model.num_channels

    def fn(items):
        groups: dict[tuple, list] = {}
        for idx, u in items:  # same (Tc, N) only: no padding inside a batch (see the header)
            groups.setdefault((int(u["cond"].shape[0]), int(u["duration"])), []).append((idx, u))
        out = {}
        for (tc, n), members in sorted(groups.items()):
            for i in range(0, len(members), batch_size):
                chunk = members[i:i + batch_size]
                cond = torch.stack([u["cond"] for _, u in chunk]).to(device, torch.float32)
                nt = max(int(u["text"].numel()) for _, u in chunk)
                text = torch.full((len(chunk), nt), -1, dtype=torch.long)
                for r, (_, u) in enumerate(chunk):
                    text[r, : u["text"].numel()] = u["text"]
                dur = max(n, tc + 1, nt + 1)  # cfm.py:300 raises the duration; the noise must match the final length
                noise = torch.stack([utterance_noise(seed, idx, min(dur, 4096), mel_dim, device) for idx, _ in chunk])
                mel, _ = model.sample(cond=cond, text=text.to(device), duration=n, steps=steps,
                                      cfg_strength=cfg_strength, sway_sampling_coef=sway_sampling_coef, noise=noise,
                                      use_acc_grl=use_acc_grl, use_prosody_encoder=False, return_trajectory=False)
                wav = vocoder.decode(mel[:, tc:, :].permute(0, 2, 1))  # utils_infer.py:545-549
                wav = wav.cpu()
                for r, (idx, _) in enumerate(chunk):
                    out[idx] = wav[r].clone()
        return out

    return fn


def synthesize_sharded(utterances: list[dict], synth_fn, *, group=None, dst: int | None = 0) -> list | None:
    """Shard -> synthesise -> host gather.  Every rank passes the same `utterances` list (host data) and its own
    `synth_fn`; rank r synthesises `shard_utterances(durations, world)[r]`.  Returns the waveforms in list order on
    rank `dst` (None elsewhere), or on every rank with dst=None.  Without an initialised process group it simply runs
    everything locally (world size 1)."""
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    mine = shard_utterances([int(u["duration"]) for u in utterances], world)[rank]
    local = synth_fn([(i, utterances[i]) for i in mine])
    if set(local) != set(mine):
        raise RuntimeError(f"rank {rank}: synth_fn returned utterances {sorted(local)} for shard {mine}")
    if world == 1:
        return [local[i] for i in range(len(utterances))]
    if dst is None:
        parts = [None] * world
        dist.all_gather_object(parts, local, group=group)
    else:
        parts = [None] * world if rank == dst else None
        dist.gather_object(local, parts, dst=dst, group=group)
        if rank != dst:
            return None
    merged = {}
    for part in parts:
        merged.update(part)
    if len(merged) != len(utterances):
        raise RuntimeError(f"gather returned {len(merged)} of {len(utterances)} utterances")
    return [merged[i] for i in range(len(utterances))]


# ------------------------------------------------------------------------------------------- two-GPU latency mode
class CfgSplit:
    """Two-GPU latency mode for ONE utterance: the conditional and the unconditional DiT forward of
    every Euler step (cfm.py:393-417) run on two GPUs and swap their `pred` over NVLink inside one kernel per step
    (`cfg_split_exchange_kernel`: peer stores + system-scope flags — no NCCL call on the data path, graph-replayable).

    Two processes (ranks 0 and 1 of `group`), one GPU each, same weights, SAME `sample()` calls with the same inputs
    and noise; rank 0 computes the conditional variant, rank 1 the unconditional one, and both return the same `out`,
    bit-identical to the single-GPU result.  Set-up (once): each rank allocates its exchange buffer and flag words and
    hands them to the peer through CUDA IPC (the handles travel over the process group)."""

    FLAG_BYTES = 256   # flag words first (ready, data), the two pred slots behind them

    def __init__(self, device, group=None, max_rows: int = 4096):
        import ctypes as C

        from . import _native as nv

        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world != 2:
            raise RuntimeError(f"CfgSplit needs exactly 2 ranks, got {world}")
        self.variant = rank
        self.device = torch.device(device)
        self.max_rows = int(max_rows)
        self._lib = nv.load()
        own, peer, handle = C.c_void_p(), C.c_void_p(), C.create_string_buffer(64)
        with torch.cuda.device(self.device):   # allocate here, map the peer's block for kernels of THIS device
            nv.check(self._lib.lemas_peer_alloc(self.FLAG_BYTES + 2 * self.max_rows * 128 * 4, C.byref(own), handle))
            both = [None, None]
            dist.all_gather_object(both, handle.raw, group=group)
            nv.check(self._lib.lemas_peer_open(both[1 - rank], C.byref(peer)))
        self._own, self._peer = own.value, peer.value
        self.flags_ptr, self.xchg_ptr = self._own, self._own + self.FLAG_BYTES
        self.peer_flags_ptr, self.peer_xchg_ptr = self._peer, self._peer + self.FLAG_BYTES
        dist.barrier(group)

    def close(self):
        """Both ranks call this (after their last split sample() has been synchronised)."""
        if getattr(self, "_peer", None):
            self._lib.lemas_peer_close(self._peer)
            self._peer = None
        if getattr(self, "_own", None):
            self._lib.lemas_peer_free(self._own)
            self._own = None
