"""world_size-2 gloo tests of the multi-GPU host logic (lemas_tts.parallel): utterance sharding and the single
weight broadcast.  The data path itself has no collective."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_utterances_balanced_and_complete():
    from lemas_tts.parallel import shard_utterances

    lengths = [2187, 940, 768, 768, 2814, 300, 1500, 1500, 64]
    shards = shard_utterances(lengths, 4)
    assert sorted(i for s in shards for i in s) == list(range(len(lengths)))
    cost = lambda n: n * (378_888_192 + 90_112 * n)
    loads = [sum(cost(lengths[i]) for i in s) for s in shards]
    assert max(loads) <= cost(max(lengths)) * 1.05  # the longest utterance bounds the makespan here
    # uniform case (config C4): exact split
    assert [len(s) for s in shard_utterances([768] * 256, 8)] == [32] * 8
    assert shard_utterances([5], 2) == [[0], []]


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lemas_tts import synthetic as syn
        from lemas_tts.parallel import broadcast_state_dict, max_over_ranks, shard_utterances

        sd = syn.make_dit_state_dict(syn.TINY_ARCH, seed=5) if rank == 0 else None
        if rank == 0:
            sd["some.int.buffer"] = torch.arange(7)
            sd["some.half"] = torch.randn(3, 5).half()
        got = broadcast_state_dict(sd, src=0)
        want = syn.make_dit_state_dict(syn.TINY_ARCH, seed=5)
        assert set(want) | {"some.int.buffer", "some.half"} == set(got)
        for k, v in want.items():
            assert torch.equal(got[k], v), k
        assert got["some.int.buffer"].tolist() == list(range(7)) and got["some.half"].dtype == torch.float16
        mine = shard_utterances([100, 900, 500, 300], world)[rank]
        assert max_over_ranks(float(rank + 1)) == float(world)
        torch.save(mine, os.path.join(tmp, f"shard{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_broadcast_and_sharding_two_ranks(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = torch.load(tmp_path / "shard0.pt")
    b = torch.load(tmp_path / "shard1.pt")
    assert sorted(a + b) == [0, 1, 2, 3] and a and b


# ------------------------------------------------------------------------------- sharded synthesis (host logic)
def _fake_synth(items):
    """Stand-in for make_synth_fn(model, vocoder): a waveform that depends on the utterance only."""
    out = {}
    for idx, u in items:
        n = (int(u["duration"]) - int(u["cond"].shape[0]) - 1) * 4
        out[idx] = torch.full((n,), float(idx)) + u["text"].float().sum() * 1e-3
    return out


def _utterances():
    g = torch.Generator().manual_seed(0)
    durs = [700, 300, 300, 512, 900, 300, 512, 128]
    return [dict(cond=torch.randn(d // 3, 100, generator=g), text=torch.randint(0, 50, (d // 10,), generator=g), duration=d)
            for d in durs]


def test_synthesize_sharded_single_process_runs_everything_locally():
    from lemas_tts.parallel import synthesize_sharded

    utts = _utterances()
    wavs = synthesize_sharded(utts, _fake_synth)
    want = _fake_synth(list(enumerate(utts)))
    assert len(wavs) == len(utts) and all(torch.equal(wavs[i], want[i]) for i in range(len(utts)))


def _shard_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lemas_tts.parallel import shard_utterances, synthesize_sharded

        utts = _utterances()
        seen = []

        def synth(items):
            seen.extend(i for i, _ in items)
            return _fake_synth(items)

        got = synthesize_sharded(utts, synth, dst=0)
        assert sorted(seen) == shard_utterances([u["duration"] for u in utts], world)[rank]
        assert (got is None) == (rank != 0)
        everywhere = synthesize_sharded(utts, _fake_synth, dst=None)
        assert len(everywhere) == len(utts)
        if rank == 0:
            torch.save(dict(got=got, everywhere=everywhere), os.path.join(tmp, "gathered.pt"))
    finally:
        dist.destroy_process_group()


def test_synthesize_sharded_two_ranks_gathers_in_list_order(tmp_path):
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_shard_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    res = torch.load(tmp_path / "gathered.pt")
    utts = _utterances()
    want = _fake_synth(list(enumerate(utts)))
    for key in ("got", "everywhere"):
        assert len(res[key]) == len(utts)
        for i in range(len(utts)):
            assert torch.equal(res[key][i], want[i]), (key, i)


def test_shard_utterances_properties():
    """Property test (hypothesis): for any list of lengths and world size the shards are a partition of the indices, in
    ascending order inside a rank, identical on every call (ranks compute them independently), and balanced the way
    longest-first greedy guarantees: no rank exceeds the lightest one by more than one utterance's cost."""
    from hypothesis import given, settings, strategies as st

    from lemas_tts.parallel import shard_utterances

    def cost(n):
        return n * (378_888_192 + 90_112 * n)

    @settings(max_examples=200, deadline=None)
    @given(st.lists(st.integers(min_value=1, max_value=4096), min_size=0, max_size=70), st.integers(1, 8))
    def check(lengths, world):
        shards = shard_utterances(lengths, world)
        assert len(shards) == world
        flat = [i for s in shards for i in s]
        assert sorted(flat) == list(range(len(lengths)))
        assert all(s == sorted(s) for s in shards)
        assert shards == shard_utterances(list(lengths), world)
        loads = [sum(cost(lengths[i]) for i in s) for s in shards]
        if lengths:
            assert max(loads) - min(loads) <= cost(max(lengths))

    check()
