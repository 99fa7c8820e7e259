"""Sharded synthesis on hardware (BASELINE.json configs[3]): every utterance of a list must
come out BIT-IDENTICAL whether the list is synthesised on one GPU or dealt to several ranks by
`lemas_tts.parallel.synthesize_sharded` (the reference's B > 1 path, /root/reference/lemas_tts/model/cfm.py:336-339).

* one-GPU test (runs in the driver's tier): the shards of a 2-rank and a 3-rank split are run one after the other in this
  process with different batch sizes and compared with the unsharded run — batch composition must not change a bit;
* two-GPU test (skipped with fewer than 2 devices; `gpurun --gpus 2`): two NCCL ranks, one weight broadcast,
  host gather on rank 0, compared with rank 0's own single-GPU run.
"""
import os

import pytest
import torch

from lemas_tts import synthetic as syn

pytestmark = pytest.mark.gpu


def _utterances(arch, shapes, seed=0):
    """shapes: list of (ref frames, total frames, tokens)."""
    utts = []
    for i, (tc, n, nt) in enumerate(shapes):
        utts.append(dict(cond=syn.synthetic_ref_mel(1, tc, arch.mel_dim, seed=seed + i)[0],
                         text=syn.synthetic_text_ids(1, nt, arch.text_num_embeds, seed=seed + i)[0], duration=n))
    return utts


def _build(arch, varch, device):
    from lemas_tts.model.backbones.dit import DiT
    from lemas_tts.model.cfm import CFM
    from lemas_tts.vocoder import Vocos

    model = CFM(transformer=DiT(**arch.to_kwargs()), mel_spec_kwargs=dict(mel_spec_type="vocos"))
    model.load_state_dict(syn.make_dit_state_dict(arch, seed=3), strict=True)
    voc = Vocos(input_channels=varch.input_channels, dim=varch.dim, intermediate_dim=varch.intermediate_dim,
                num_layers=varch.num_layers)
    voc.load_state_dict(syn.make_vocos_state_dict(varch, seed=7), strict=True)
    return model.to(device), voc.to(device).eval()


SHAPES = [(40, 128, 30), (64, 200, 45), (40, 128, 22), (40, 128, 30), (64, 200, 50), (40, 128, 9), (33, 97, 20),
          (64, 200, 45)]


@pytest.mark.parametrize("arch_name", ["TINY_ARCH", "FULL_ARCH"])
def test_shards_are_bit_identical_to_the_unsharded_run(arch_name):
    from lemas_tts.parallel import make_synth_fn, shard_utterances

    arch = getattr(syn, arch_name)
    varch = syn.TINY_VOCOS if arch_name == "TINY_ARCH" else syn.FULL_VOCOS
    model, voc = _build(arch, varch, "cuda")
    utts = _utterances(arch, SHAPES if arch_name == "TINY_ARCH" else SHAPES[:5])
    steps = 4 if arch_name == "TINY_ARCH" else 3
    whole = make_synth_fn(model, voc, steps=steps, seed=11, batch_size=8)(list(enumerate(utts)))
    for world, bs in ((2, 3), (3, 1)):
        shards = shard_utterances([u["duration"] for u in utts], world)
        fn = make_synth_fn(model, voc, steps=steps, seed=11, batch_size=bs)
        for shard in shards:
            part = fn([(i, utts[i]) for i in shard])
            for i in shard:
                assert part[i].shape == ((utts[i]["duration"] - utts[i]["cond"].shape[0] - 1) * 256,)
                assert torch.equal(part[i], whole[i]), f"utterance {i}: world {world}, batch size {bs}"
    assert all(torch.isfinite(w).all() for w in whole.values())


def _rank_main(rank, world, port, out_path):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from lemas_tts.model.backbones.dit import DiT
        from lemas_tts.model.cfm import CFM
        from lemas_tts.parallel import broadcast_state_dict, make_synth_fn, synthesize_sharded
        from lemas_tts.vocoder import Vocos

        arch, varch = syn.TINY_ARCH, syn.TINY_VOCOS
        sd = None
        if rank == 0:  # rank 0 "loads" the checkpoint; ONE broadcast of the packed blob; everybody keeps a replica
            sd = dict(syn.make_dit_state_dict(arch, seed=3))
            sd.update({"vocos." + k: v for k, v in syn.make_vocos_state_dict(varch, seed=7).items()})
        sd = broadcast_state_dict(sd, src=0, device=dev)
        model = CFM(transformer=DiT(**arch.to_kwargs()), mel_spec_kwargs=dict(mel_spec_type="vocos"))
        model.load_state_dict({k: v for k, v in sd.items() if not k.startswith("vocos.")}, strict=True)
        voc = Vocos(input_channels=varch.input_channels, dim=varch.dim, intermediate_dim=varch.intermediate_dim,
                    num_layers=varch.num_layers)
        voc.load_state_dict({k[6:]: v for k, v in sd.items() if k.startswith("vocos.")}, strict=True)
        model, voc = model.to(dev), voc.to(dev).eval()
        utts = _utterances(arch, SHAPES)
        fn = make_synth_fn(model, voc, steps=4, seed=11, batch_size=4)
        got = synthesize_sharded(utts, fn, dst=0)
        if rank == 0:
            whole = fn(list(enumerate(utts)))
            same = all(torch.equal(got[i], whole[i]) for i in range(len(utts)))
            torch.save(dict(same=same, n=len(got)), out_path)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_ranks_nccl_bit_identical_to_one_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp

    out = tmp_path / "res.pt"
    mp.spawn(_rank_main, args=(2, 29700 + os.getpid() % 1000, str(out)), nprocs=2, join=True)
    res = torch.load(out)
    assert res["n"] == len(SHAPES) and res["same"], "sharded utterances differ from the single-GPU run"


# ------------------------------------------------------------------------------------------- two-GPU latency mode
def _split_main(rank, world, port, out_path):
    import time

    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from lemas_tts.model.backbones.dit import DiT
        from lemas_tts.model.cfm import CFM
        from lemas_tts.parallel import CfgSplit

        res = {}
        split = CfgSplit(dev)
        for arch_name, tc, n, nt, steps in (("TINY_ARCH", 60, 300, 40, 6), ("FULL_ARCH", 937, 2187, 350, 8)):
            arch = getattr(syn, arch_name)
            model = CFM(transformer=DiT(**arch.to_kwargs()), mel_spec_kwargs=dict(mel_spec_type="vocos"))
            model.load_state_dict(syn.make_dit_state_dict(arch, seed=3), strict=True)   # same seeded weights on both ranks
            model = model.to(dev)
            kw = dict(cond=syn.synthetic_ref_mel(1, tc, arch.mel_dim, seed=1).to(dev),
                      text=syn.synthetic_text_ids(1, nt, arch.text_num_embeds, seed=1).to(dev), duration=n, steps=steps,
                      cfg_strength=2.0, sway_sampling_coef=3.0, noise=syn.synthetic_noise([n], arch.mel_dim, seed=1),
                      use_acc_grl=False, return_trajectory=False)
            whole, _ = model.sample(**kw)                  # both CFG variants on this GPU
            model.cfg_split = split
            for _ in range(2):
                got, _ = model.sample(**kw)                # one variant here, the other on the peer
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            got, _ = model.sample(**kw)
            torch.cuda.synchronize()
            t_split = time.perf_counter() - t0
            model.cfg_split = None
            t0 = time.perf_counter()
            model.sample(**kw)
            torch.cuda.synchronize()
            t_whole = time.perf_counter() - t0
            res[arch_name] = dict(same=bool(torch.equal(got, whole)), t_split=t_split, t_whole=t_whole)
            dist.barrier()
        if rank == 0:
            torch.save(res, out_path)
        # rank 1 must have produced the same state as well
        flag = torch.tensor([int(all(v["same"] for v in res.values()))], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        assert int(flag.item()) == 1
        split.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_gpu_cfg_split_is_bit_identical_and_faster(tmp_path):
    """cond / uncond forwards on two GPUs with a fused NVLink `pred` exchange per step."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp

    out = tmp_path / "split.pt"
    mp.spawn(_split_main, args=(2, 29800 + os.getpid() % 1000, str(out)), nprocs=2, join=True)
    res = torch.load(out)
    for name, r in res.items():
        print(f"{name}: split {r['t_split'] * 1e3:.1f} ms vs one GPU {r['t_whole'] * 1e3:.1f} ms, identical {r['same']}")
        assert r["same"], f"{name}: the two-GPU split changed the result"
