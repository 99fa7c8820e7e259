#!/usr/bin/env python
"""bench.py — mel-frames/s and RTF of the LEMAS-TTS acoustic hot path (CFM.sample + Vocos.decode) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C1|C2|C3|C4|C5] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one full synthesis of the workload's utterance batch on each GPU: the 32-NFE ODE loop (64 co-batched DiT
forwards) and the vocoder.  Workload at N=1 = BASELINE.json configs[1] ("C2": batch 1, 10 s reference = 937 mel
frames, 350 phones, 1250 generated frames, NFE 32, CFG 2.0, sway 5 -> 3.4856).  N>1 = one such replica per GPU
(utterance sharding, no data-path collective, weak scaling); weights are broadcast once from rank 0 over NCCL.

Prints ONE JSON line (rank 0).  `value` = generated mel frames / s with inputs resident in HBM; `e2e` = the same
through the public API (CFM.sample + vocoder.decode) from pinned HOST buffers, H2D and D2H copies inside the timed
region; `roofline` = the attention kernel (the largest tensor-core kernel of the step) from CUDA events recorded
around each of its launches in a profiled step of the same workload; `cpu_baseline` = the CPU oracle port timed on
this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for p in (ROOT, ROOT / "lemas-tts_b200"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

import torch  # noqa: E402

from lemas_tts import synthetic as syn  # noqa: E402

METRIC, UNIT = "mel_frames_per_sec", "mel-frames/s"
HOP, SR = 256, 24000


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", choices=list(syn.CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU utterance batch (C4: default 32)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--all-rows", action="store_true",
                    help="C3 (ragged batch): also compute the padded rows that cannot reach a valid row (reference-"
                         "identical padding; default: skip them, lemas_sample_args.flags)")
    ap.add_argument("--no-c4", action="store_true", help="skip the sharded C4 block (256 utterances over the ranks)")
    ap.add_argument("--vocoder", default="vocos", choices=["vocos", "bigvgan"],
                    help="vocos (both shipped configs) or the BigVGAN-v2 generator of the `mel_spec_type: bigvgan` branch "
                         "(112 M parameters, seeded weights)")
    ap.add_argument("--latency-split", action="store_true",
                    help="--gpus 2 only: add a `latency_split` block — ONE utterance of the workload with the conditional "
                         "and unconditional forwards on the two GPUs (lemas_tts.parallel.CfgSplit) against one GPU")
    return ap.parse_args()


class Workload:
    """One BASELINE.json config made concrete: host inputs, CFM.sample keywords and frame accounting."""

    def __init__(self, name: str, batch_override: int = 0, seed_offset: int = 0):
        cfg = syn.CONFIGS[name]
        self.cfg, self.name = cfg, name
        arch = syn.FULL_ARCH
        batch = cfg.batch
        if name == "C4":
            batch = 32  # 256 utterances over 8 GPUs (BASELINE.json configs[3]); one rank runs its 32-utterance shard
        if batch_override:
            batch = batch_override
        self.batch = batch
        seed = cfg.seed + seed_offset
        self.prosody = name == "C3"
        self.edit_mask = None
        self.lens = None
        if name == "C3":  # ragged, raw reference audio, prosody model
            b = syn.c3_batch(batch, seed=seed)
            self.cond, self.text, self.lens, self.duration = b["audio"], b["text"], b["lens"], b["duration"]
            self.N = int(self.duration.max())
            self.gen_slices = [(int(l), int(d)) for l, d in zip(self.lens, self.duration)]
        else:
            self.cond = syn.synthetic_ref_mel(batch, cfg.ref_frames, arch.mel_dim, seed=seed)
            self.text = syn.synthetic_text_ids(batch, cfg.n_text, arch.text_num_embeds, seed=seed)
            self.duration = cfg.total_frames
            self.N = cfg.total_frames
            if name == "C5":  # speech edit: regenerate 3 s in the middle, keep the rest; whole utterance re-vocoded
                self.edit_mask = torch.ones(batch, cfg.ref_frames, dtype=torch.bool)
                self.edit_mask[:, 1125:1406] = False
                self.gen_slices = [(0, cfg.total_frames)] * batch
            else:
                self.gen_slices = [(cfg.ref_frames, cfg.total_frames)] * batch
        self.frames = sum(e - s for s, e in self.gen_slices)  # mel frames that are vocoded = the metric's numerator
        self.uniform = len(set(self.gen_slices)) == 1

    def describe(self) -> str:
        c = self.cfg
        shape = (f"ragged N<={self.N} (raw 10 s reference audio, prosody encoder on)" if self.name == "C3"
                 else f"ref {c.ref_frames} frames, N {c.total_frames}, {c.n_text} phones")
        return (f"{self.name}: batch {self.batch}/GPU, {shape}, NFE {c.steps}, cfg {c.cfg_strength}, sway {c.sway_coef}")

    def sample_kwargs(self, dev):
        c = self.cfg
        kw = dict(steps=c.steps, cfg_strength=c.cfg_strength, sway_sampling_coef=c.sway_coef, use_acc_grl=False,
                  return_trajectory=False, use_prosody_encoder=self.prosody)
        kw["duration"] = self.duration.to(dev) if torch.is_tensor(self.duration) else self.duration
        if self.lens is not None:
            kw["lens"] = self.lens.to(dev)
        if self.edit_mask is not None:
            kw["edit_mask"] = self.edit_mask.to(dev)
        return kw


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(tflops=d["bf16_tflops_sustained"], tflops_burst=d["bf16_tflops"], hbm=d["hbm_gbs"],
                    source="MEASURED_PEAKS.json (sustained bf16 GEMM; kernel timed inside a long step)")
    return dict(tflops=1400.0, tflops_burst=1590.0, hbm=6650.0, source="fallback of B200_PROFILING.md")


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU reference arm

_REF_MODELS = {}


def _reference_model(wl: "Workload"):
    """The reference's own CFM(DiT) (oracle/verbatim.py: /root/reference, else the vendored oracle/_ref), fp32 on the
    CPU, with the same seeded weights as the B200 arm."""
    import dataclasses
    import tempfile

    from oracle import verbatim

    key = wl.prosody
    if key not in _REF_MODELS:
        arch = dataclasses.replace(syn.FULL_ARCH, use_prosody_encoder=wl.prosody)
        sd = dict(syn.make_dit_state_dict(arch, seed=0))
        paths = None
        if wl.prosody:
            tmp = tempfile.mkdtemp(prefix="lemas_prosody_ref_")
            paths = syn.write_prosody_assets(tmp)
            sd.update({"prosody_encoder.encoder." + k: v for k, v in syn.make_prosody_state_dict().items()})
        _REF_MODELS[key] = verbatim.build_reference_cfm(arch, sd, prosody_paths=paths)
    return _REF_MODELS[key]


def cpu_reference_sample(wl: "Workload", euler_steps: int = 4, max_batch: int = 2):
    """One bounded sample of the workload on the host cores: the reference's own `CFM.sample`
    (/root/reference/lemas_tts/model/cfm.py:206-473, all of it: mel / prosody front-end, text embedding, `euler_steps`
    REAL Euler steps of two DiT forwards each) on the first `max_batch` utterances, then the vocoder on the generated
    frames (oracle/vocos_oracle.py: pip `vocos` is absent offline).  Returns the MEASURED seconds and, separately, the
    time scaled to the workload's step count — only the Euler loop (timed inside the torchdiffeq shim) is scaled, the
    once-per-utterance parts are not."""
    from oracle import verbatim
    from oracle import vocos_oracle as vo

    if not verbatim.available():
        return cpu_port_sample(wl, euler_steps, max_batch)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = wl.cfg
    model = _reference_model(wl)
    vsd = syn.make_vocos_state_dict(syn.FULL_VOCOS, seed=7)
    nb = min(wl.batch, max_batch)
    slices = wl.gen_slices[:nb]
    kw = dict(steps=euler_steps, cfg_strength=cfg.cfg_strength, sway_sampling_coef=cfg.sway_coef, use_acc_grl=False,
              use_prosody_encoder=wl.prosody)
    if torch.is_tensor(wl.duration):
        kw["duration"] = wl.duration[:nb] if nb > 1 else int(wl.duration[0])
    else:
        kw["duration"] = wl.duration
    if wl.lens is not None:
        kw["lens"] = wl.lens[:nb]
    if wl.edit_mask is not None:
        kw["edit_mask"] = wl.edit_mask[:nb]
    with torch.inference_mode():
        t0 = time.perf_counter()
        out, _ = model.sample(cond=wl.cond[:nb], text=wl.text[:nb], **kw)
        t_sample = time.perf_counter() - t0
        t_loop = float(verbatim.LAST_LOOP_SECONDS)
        t0 = time.perf_counter()
        for b, (s0, e0) in enumerate(slices):
            vo.vocos_decode(vsd, out[b:b + 1, s0:e0].float().transpose(1, 2).contiguous())
        t_voc = time.perf_counter() - t0
    measured = t_sample + t_voc
    scaled = (t_sample - t_loop) + t_loop * cfg.steps / euler_steps + t_voc
    frames = sum(e - s0 for s0, e in slices)
    return dict(kind="reference", seconds=scaled, measured_seconds=measured, loop_seconds=t_loop, t_vocos=t_voc,
                euler_steps=euler_steps, frames=frames, cores=cores,
                sample=f"reference CFM.sample (verbatim /root/reference modules, fp32, {cores} threads) on the first {nb} of "
                       f"{wl.batch} utterance(s) with {euler_steps} real Euler steps (2 DiT forwards each, N={out.shape[1]}) + "
                       f"Vocos decode of {frames} frames (oracle port); the Euler-loop time is scaled x{cfg.steps}/"
                       f"{euler_steps} to the workload's {cfg.steps} steps, everything else counted once")


def cpu_port_sample(wl: "Workload", euler_steps: int = 4, max_batch: int = 2):
    """Fallback when the reference sources are not available (no /root/reference, no oracle/_ref): the CPU oracle port
    (oracle/lemas_oracle.py + vocos_oracle.py, fp32, all host threads), same bounded sample, same scaling rule."""
    from oracle import lemas_oracle as orc
    from oracle import vocos_oracle as vo

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = wl.cfg
    arch = syn.FULL_ARCH
    sd = syn.make_dit_state_dict(arch, seed=0)
    vsd = syn.make_vocos_state_dict(syn.FULL_VOCOS, seed=7)
    nb = min(wl.batch, max_batch)
    slices = wl.gen_slices[:nb]
    N = max(e for _, e in slices)
    text = wl.text[:nb]
    if wl.name == "C3":  # mel of the raw reference audio (its prosody encoder pass is not part of the sample)
        cond = orc.mel_spectrogram(wl.cond[:nb]).permute(0, 2, 1)
        for b, (l, _) in enumerate(slices):
            cond[b, l:] = 0
    else:
        cond = wl.cond[:nb]
    durs = [e for _, e in slices]
    noise = syn.synthetic_noise(durs, arch.mel_dim, seed=cfg.seed)
    condp = torch.nn.functional.pad(cond, (0, 0, 0, N - cond.shape[1]))[:, :N]
    mask = orc.lens_to_mask(torch.tensor(durs)) if nb > 1 else None
    tgrid = orc.time_grid(cfg.steps, cfg.sway_coef)
    with torch.inference_mode():
        t0 = time.perf_counter()
        tc = orc.text_embedding(sd, arch, text, N, drop_text=False)
        tu = orc.text_embedding(sd, arch, text, N, drop_text=True)
        t_text = time.perf_counter() - t0
        y = noise
        t0 = time.perf_counter()
        for i in range(euler_steps):
            t = tgrid[i]
            pc = orc.dit_forward(sd, arch, y, condp, tc, t, mask, False)
            pu = orc.dit_forward(sd, arch, y, condp, tu, t, mask, True)
            y = y + (tgrid[i + 1] - t) * (pc + (pc - pu) * (cfg.cfg_strength * (1 - t) ** 2)).clamp(-20, 20)
        t_loop = time.perf_counter() - t0
        t0 = time.perf_counter()
        for b, (s0, e0) in enumerate(slices):
            vo.vocos_decode(vsd, y[b:b + 1, s0:e0].transpose(1, 2).contiguous())
        t_voc = time.perf_counter() - t0
    frames = sum(e - s0 for s0, e in slices)
    return dict(kind="port", seconds=t_text + t_loop * cfg.steps / euler_steps + t_voc,
                measured_seconds=t_text + t_loop + t_voc, loop_seconds=t_loop, t_vocos=t_voc, euler_steps=euler_steps,
                frames=frames, cores=cores,
                sample=f"oracle port, fp32, {cores} threads, first {nb} of {wl.batch} utterance(s): text-embed x2 + "
                       f"{euler_steps} real Euler step(s) (2 DiT forwards each, B={nb}, N={N}) + Vocos decode of {frames} "
                       f"frames; the Euler-loop time is scaled x{cfg.steps}/{euler_steps}")


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores (all threads).
    Every step is a bounded sample (4 real Euler steps of the workload); `ms_per_step` is what ELAPSED per sample,
    `value` uses the time scaled to the workload's step count (both are printed).  One host process: under torchrun
    only rank 0 runs, and the number does not scale with --gpus."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = Workload(args.workload, args.batch)
    scaled, measured, info = [], [], None
    for i in range(args.warmup + args.steps):
        info = cpu_reference_sample(wl, euler_steps=4)
        if i >= args.warmup:
            scaled.append(info["seconds"])
            measured.append(info["measured_seconds"])
    sec = sum(scaled) / len(scaled)
    msec = sum(measured) / len(measured)
    value = info["frames"] / sec
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": msec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "rtf": sec / (info["frames"] * HOP / SR),
            "config": {"workload": wl.describe()},
            "measured": {"seconds_per_sample": msec, "euler_steps_per_sample": info["euler_steps"],
                         "euler_loop_seconds": info["loop_seconds"], "vocoder_seconds": info["t_vocos"],
                         "scaled_seconds_full_nfe": sec,
                         "note": "ms_per_step is the measured sample; value = frames / scaled_seconds_full_nfe"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": info["cores"], "kind": info["kind"],
                             "sample": info["sample"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if args.gpus > 1:
        line["cpu_arm_scope"] = "one host process (rank 0); not applicable to --gpus scaling"
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ B200 arm


def run_b200(args):
    import tempfile

    import torch.distributed as dist

    from lemas_tts import _native as nv
    from lemas_tts.model.backbones.dit import DiT
    from lemas_tts.model.cfm import CFM
    from lemas_tts.parallel import broadcast_state_dict, max_over_ranks
    from lemas_tts.vocoder import Vocos

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    nv.require_device()
    if world > 1:
        # NCCL prints its version banner on stdout at init; stdout must carry exactly one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    wl = Workload(args.workload, args.batch, seed_offset=100 * rank)  # every rank synthesises its own utterances
    cfg, batch = wl.cfg, wl.batch
    import dataclasses

    arch = dataclasses.replace(syn.FULL_ARCH, use_prosody_encoder=wl.prosody)
    # weights: rank 0 "loads" (seeded factory), ONE NCCL broadcast of the packed blob, every rank keeps a replica
    def all_weights():
        sd = dict(syn.make_dit_state_dict(arch, seed=0))
        if wl.prosody:
            sd.update({"prosody_encoder.encoder." + k: v for k, v in syn.make_prosody_state_dict().items()})
        sd.update({"vocos." + k: v for k, v in syn.make_vocos_state_dict(syn.FULL_VOCOS, seed=7).items()})
        return sd

    if world > 1:
        sd_all = broadcast_state_dict(all_weights() if rank == 0 else None, src=0, device=dev)
    else:
        sd_all = all_weights()
    sd = {k: v for k, v in sd_all.items() if not k.startswith("vocos.")}
    vsd = {k[6:]: v for k, v in sd_all.items() if k.startswith("vocos.")}
    pros = {}
    if wl.prosody:
        tmp = tempfile.mkdtemp(prefix="lemas_prosody_")
        cfg_path, ckpt_path = syn.write_prosody_assets(tmp)
        pros = dict(use_prosody_encoder=True, prosody_cfg_path=str(cfg_path), prosody_ckpt_path=str(ckpt_path))
    model = CFM(transformer=DiT(**arch.to_kwargs()), mel_spec_kwargs=dict(mel_spec_type="vocos"), **pros)
    model.load_state_dict(sd, strict=True)
    model = model.to(dev)
    model.skip_padded_rows = wl.name == "C3" and not args.all_rows
    if args.vocoder == "bigvgan":   # every rank builds the same seeded generator (no broadcast needed for the bench)
        from lemas_tts.bigvgan import BigVGAN

        big = BigVGAN()
        big.load_state_dict(syn.make_bigvgan_state_dict(syn.FULL_BIGVGAN, seed=17), strict=True)
        big = big.to(dev).eval()

        class _AsDecode:   # same call shape as Vocos.decode for the code below: [B, 100, T] -> [B, samples]
            def decode(self, mel):
                return big(mel)[:, 0, : (mel.shape[-1] - 1) * HOP]

        voc = _AsDecode()
    else:
        voc = Vocos()
        voc.load_state_dict(vsd, strict=True)
        voc = voc.to(dev).eval()
    del sd, vsd, sd_all

    cond_h, text_h = wl.cond.pin_memory(), wl.text.pin_memory()
    cond_d, text_d = cond_h.to(dev), text_h.to(dev)
    kw = wl.sample_kwargs(dev)
    frames = wl.frames
    wav_h = torch.empty(sum((e - s - 1) * HOP for s, e in wl.gen_slices)).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    use_graph = not args.no_graph

    def synth(cond, text, seed):
        out, _ = model.sample(cond=cond, text=text, seed=seed, **kw)
        if wl.uniform:
            s0, e0 = wl.gen_slices[0]
            return [voc.decode(out[:, s0:e0, :].permute(0, 2, 1))]
        return [voc.decode(out[b:b + 1, s0:e0, :].permute(0, 2, 1)) for b, (s0, e0) in enumerate(wl.gen_slices)]

    def step_resident(i):
        return synth(cond_d, text_d, 1000 + i)

    def step_e2e(i):
        wavs = synth(cond_h.to(dev, non_blocking=True), text_h.to(dev, non_blocking=True), 1000 + i)
        off = 0
        for w in wavs:
            n = w.numel()
            wav_h[off:off + n].copy_(w.reshape(-1), non_blocking=True)
            off += n
        return wavs

    if not use_graph:
        eng0 = model.transformer.engine()
        orig = eng0.sample_loop
        eng0.sample_loop = lambda *a, **k: orig(*a, **{**k, "use_graph": False})

    def timed(fn, n_warm, n_steps):
        for i in range(n_warm):
            fn(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        total = 0.0
        l0 = nv.load().lemas_launch_count()
        for i in range(n_steps):
            flush.zero_()  # L2 flush between timed iterations (not timed)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(n_warm + i)
            e1.record()
            e1.synchronize()
            total += e0.elapsed_time(e1)
        torch.cuda.synchronize()
        launches = nv.load().lemas_launch_count() - l0
        if world > 1:
            dist.barrier()
            total = max_over_ranks(total, device=dev)
        return total, launches

    sampler = ClockSampler(local) if rank == 0 else None
    ms_total, launches = timed(step_resident, args.warmup, args.steps)
    clocks = sampler.stop() if sampler else None
    ms_e2e, _ = timed(step_e2e, 1, args.steps)

    # profiled step: CUDA events around every launch of the sampler, by kernel kind (same workload, same process).
    # The event pairs are recorded INSIDE the captured ODE-step graph and read back after every replay
    # (lemas_engine_profile mode 2), so the per-kernel times describe the graph-replayed step that is timed above —
    # not an eager step with host launch gaps.  Step 0 (run eagerly before the capture) is not counted.
    eng = model.transformer.engine()
    eng.profile(2 if use_graph else 1)
    eng.profile_read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush.zero_()
    e0.record()
    step_resident(10_000)
    e1.record()
    prof = eng.profile_read()
    eng.profile(0)
    prof_ms = e0.elapsed_time(e1)

    # BASELINE.json configs[3]: 256 utterances (256 reference + 512 generated frames) dealt to the ranks by
    # shard_utterances and synthesised in batches of 32 through lemas_tts.parallel (strong scaling: the list is fixed,
    # the ranks share it).  Device time of each rank's shard (CFM.sample + Vocos.decode + D2H of the waveforms), max
    # over ranks; the host-side gather of the waveforms follows outside the timed region.
    c4 = None
    if args.workload == "C2" and not args.no_c4:
        from lemas_tts.parallel import make_synth_fn, shard_utterances, synthesize_sharded

        c4cfg = syn.CONFIGS["C4"]
        n_utt = c4cfg.batch
        cond4 = syn.synthetic_ref_mel(n_utt, c4cfg.ref_frames, arch.mel_dim, seed=c4cfg.seed)
        text4 = syn.synthetic_text_ids(n_utt, c4cfg.n_text, arch.text_num_embeds, seed=c4cfg.seed)
        utts = [dict(cond=cond4[i], text=text4[i], duration=c4cfg.total_frames) for i in range(n_utt)]
        fn = make_synth_fn(model, voc, steps=c4cfg.steps, cfg_strength=c4cfg.cfg_strength,
                           sway_sampling_coef=c4cfg.sway_coef, seed=c4cfg.seed, batch_size=32)
        mine = shard_utterances([u["duration"] for u in utts], world)[rank]
        fn([(i, utts[i]) for i in mine[:32]])  # warm-up: one batch (graph capture, workspace growth)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        local = fn([(i, utts[i]) for i in mine])
        e1.record()
        e1.synchronize()
        ms_c4 = e0.elapsed_time(e1)
        if world > 1:
            ms_c4 = max_over_ranks(ms_c4, device=dev)
        wavs = synthesize_sharded(utts, lambda items: {i: local[i] for i, _ in items}, dst=0)  # host gather only
        gen = n_utt * (c4cfg.total_frames - c4cfg.ref_frames)
        if rank == 0:
            assert len(wavs) == n_utt and all(w.numel() == (c4cfg.total_frames - c4cfg.ref_frames - 1) * HOP for w in wavs)
            c4 = {"workload": f"C4: {n_utt} utterances (ref {c4cfg.ref_frames} frames, N {c4cfg.total_frames}, "
                              f"{c4cfg.n_text} phones, NFE {c4cfg.steps}) sharded over {world} rank(s), batches of 32",
                  "scaling": "strong", "seconds": ms_c4 * 1e-3, "value": gen / (ms_c4 * 1e-3), "unit": UNIT,
                  "utterances_per_rank": len(mine), "x_realtime": (gen * HOP / SR) / (ms_c4 * 1e-3)}

    # SURVEY.md §8 f4, two-GPU latency mode: both ranks synthesise the SAME utterance; rank 0 runs the conditional DiT
    # forward of every Euler step, rank 1 the unconditional one, `pred` swapped over NVLink inside one kernel per step.
    lat = None
    if args.latency_split:
        if world != 2 or cfg.cfg_strength < 1e-5:
            raise SystemExit("--latency-split needs --gpus 2 and a workload with classifier-free guidance")
        from lemas_tts.parallel import CfgSplit

        wl0 = wl if rank == 0 else Workload(args.workload, args.batch)   # rank 0's utterance on both ranks
        c0, t0_ = wl0.cond.to(dev), wl0.text.to(dev)
        kw0 = wl0.sample_kwargs(dev)

        def one(i):
            out, _ = model.sample(cond=c0, text=t0_, seed=5000 + i, **kw0)
            s0, e0 = wl0.gen_slices[0]
            return out, voc.decode(out[:, s0:e0, :].permute(0, 2, 1))

        ref_out, _ = one(0)
        ms_one, _ = timed(one, args.warmup, args.steps)
        split = CfgSplit(dev)
        model.cfg_split = split
        got_out, _ = one(0)
        same = torch.tensor([int(torch.equal(got_out, ref_out))], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        ms_split, _ = timed(one, args.warmup, args.steps)
        model.cfg_split = None
        torch.cuda.synchronize()
        dist.barrier()
        split.close()
        lat = {"workload": wl0.describe() + " — one utterance, cond / uncond forwards on 2 GPUs, fused NVLink pred exchange",
               "ms_one_gpu": ms_one / args.steps, "ms_two_gpus": ms_split / args.steps,
               "speedup": ms_one / ms_split, "bit_identical_to_one_gpu": bool(same.item()),
               "xchg_bytes_per_step_per_direction": wl0.batch * wl0.N * 128 * 4}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    N, D, L = wl.N, arch.dim, arch.depth
    variants = 2 if cfg.cfg_strength >= 1e-5 else 1
    att_ms, att_n = prof["attention"]
    if wl.name == "C3":  # ragged keys: only kv_len_b keys (and query tiles below kv_len_b) are computed
        att_flops = sum(4.0 * d * d * (arch.heads * 64) for _, d in wl.gen_slices) * variants
    else:
        att_flops = 4.0 * N * N * (arch.heads * 64) * batch * variants  # QK^T + PV, per launch (one layer)
    att_tflops = att_flops / (att_ms / att_n * 1e-3) / 1e12 if att_n else 0.0
    # what the measurement itself adds: an EMPTY event pair recorded once per ODE step in the same graph
    tare_ms, tare_n = prof.get("tare", (0.0, 0))
    tare_us = tare_ms / tare_n * 1e3 if tare_n else None
    att_tflops_tare = (att_flops / ((att_ms / att_n - tare_ms / tare_n) * 1e-3) / 1e12
                       if att_n and tare_n and att_ms / att_n > tare_ms / tare_n else None)
    traffic = None
    tp = ROOT / "profiles" / "roofline_traffic.json"
    if tp.exists():
        traffic = json.loads(tp.read_text()).get(f"attention_{cfg.name}_b{batch}")
    flops_fwd = N * (378_888_192 + 90_112 * N)
    dit_flops = variants * cfg.steps * batch * flops_fwd
    sec_step = ms_total / args.steps * 1e-3
    kinds = {}
    for k, (ms, n) in prof.items():
        if n and k != "tare":
            kinds[k] = {"ms": round(ms, 3), "launches": n, "share": round(ms / prof_ms, 4)}
    gemm_flops = {"gemm_qkv": 2.0 * 3 * D * D, "gemm_out": 2.0 * D * D, "gemm_ff1": 2.0 * D * D * arch.ff_mult,
                  "gemm_ff2": 2.0 * D * D * arch.ff_mult}
    # rows the GEMMs usefully process per launch: every row of a uniform batch; the VALID rows of a ragged one (padded
    # rows are either skipped or computed for nothing — neither counts as work)
    rows_alg = sum(d for _, d in wl.gen_slices) if wl.name == "C3" else N * batch
    for k, per_tok in gemm_flops.items():
        ms, n = prof[k]
        if n:
            kinds[k]["tflops"] = round(per_tok * rows_alg * variants / (ms / n * 1e-3) / 1e12, 1)
    if att_n:
        kinds["attention"]["tflops"] = round(att_tflops, 1)
    ms, n = prof["ln_mod"]
    if n:  # algorithmic bytes: read fp32 x, write fp16 (6 B / element)
        kinds["ln_mod"]["gbs"] = round(6.0 * D * rows_alg * variants / (ms / n * 1e-3) / 1e9, 1)

    value = world * frames * args.steps / (ms_total * 1e-3)
    e2e = world * frames * args.steps / (ms_e2e * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16", "data": "synthetic",
        "rtf": sec_step / (frames * HOP / SR), "x_realtime": (frames * HOP / SR) / sec_step,
        "config": {"workload": wl.describe(),
                   "padded_rows": ("skipped beyond the influence cone of the position-embedding convolutions (valid rows "
                                   "unchanged, tests/test_fullnfe_gpu.py)" if model.skip_padded_rows else "computed"),
                   "weights": "full 336M-parameter DiT (22 layers) + " + ("BigVGAN-v2 (112M)" if args.vocoder == "bigvgan" else "Vocos")
                              + ", seeded random init (no checkpoints offline)",
                   "step": f"CFM.sample ({variants * cfg.steps} co-batched DiT forwards, graph-replayed ODE steps) + "
                           + ("BigVGAN" if args.vocoder == "bigvgan" else "Vocos.decode") + " of one utterance batch per GPU",
                   "l2": "256 MiB buffer written between timed iterations; per-step working set (0.7 GB weights) > L2",
                   "parallelism": f"utterance sharding x{world}, one weight broadcast, no per-step collective"},
        "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": world * (cond_h.numel() * cond_h.element_size() + text_h.numel() * 8),
                "d2h_bytes_per_step": world * wav_h.numel() * 4},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor",
                     "kernel": ("attention9_kernel (tcgen05, csrc/attention9.cu: short / ragged sequences)"
                                if (wl.name == "C3" or N <= 1024) else "attention_kernel (tcgen05, csrc/attention.cu)"),
                     "achieved": att_tflops, "peak": pk["tflops"], "unit": "TFLOP/s",
                     "frac": att_tflops / pk["tflops"], "traffic": traffic, "peak_source": pk["source"],
                     "flops_per_launch": att_flops, "launch_ms": att_ms / att_n if att_n else None,
                     "event_pair_tare_us": tare_us, "achieved_minus_tare": att_tflops_tare,
                     "note": "head dim 64: a 128x128 score block needs 1024 SFU clk (16 exp2/clk/SM, measured) against 512 MMA clk; "
                             "P is kept in TMEM (TS-form P V) and 1/4 of the exp2 run on the FMA pipe — see DESIGN.md §6; `achieved` uses the "
                             "bracketed launch time as measured; `event_pair_tare_us` is an empty event pair in the same graph "
                             "(each bracket also cuts the programmatic-dependent-launch overlap with its neighbours)"},
        "sampler_tensor_frac": dit_flops / sec_step / 1e12 / pk["tflops"],
        "kernels": kinds, "profiled_step_ms": prof_ms,
        "kernels_note": ("per-kernel CUDA events recorded inside the replayed step graph; steps 1.." + str(cfg.steps - 1) +
                         " of one synthesis (step 0 runs eagerly before the capture and is not counted); `share` is of "
                         "profiled_step_ms, which includes a stream synchronisation after every step") if use_graph
                        else "per-kernel CUDA events around eager launches",
        "clocks": clocks,
    }
    if c4 is not None:
        line["c4_sharded"] = c4
    if lat is not None:
        line["latency_split"] = lat
    if world == 1 and not args.no_cpu_baseline:
        info = cpu_reference_sample(wl, euler_steps=2)
        line["cpu_baseline"] = {"value": info["frames"] / info["seconds"], "unit": UNIT, "cores": info["cores"],
                                "kind": info["kind"], "sample": info["sample"],
                                "measured_seconds": info["measured_seconds"],
                                "scaled_seconds_full_nfe": info["seconds"]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
