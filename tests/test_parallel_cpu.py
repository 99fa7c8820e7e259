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
