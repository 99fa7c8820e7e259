"""Mint FULL-SIZE, FULL-NFE goldens (tests/golden/full_*.pt) from the VERBATIM reference (build container only).

    python oracle/gen_golden_full.py [case ...]          # ~45 min of CPU for all four

BASELINE.json configs at the sizes and step counts they name, through the reference's own loop
(/root/reference/lemas_tts/model/cfm.py:382-456, fp32 on the CPU, noise injected):
  full_C1      B=1, raw 4 s reference audio (96 000 samples -> 376 frames through MelSpec), N=940, 16 steps, cfg 2, sway 5
  full_C2      B=1, N=2187 (937 reference + 1250 generated frames), 32 steps, cfg 2, sway 5
  full_C5      B=1, N=2814, edit mask False on [1125,1406), 64 steps, cfg 5, sway 3
  full_C4_b4   slice of C4: B=4, N=768 uniform (mask path cfm.py:336-339), 32 steps, cfg 2, sway 3
  full_C3_b4   slice of C3: B=4 ragged, raw reference audio, prosody encoder (full Pretssel ECAPA), 32 steps
Only `out` is stored (fp32) plus input checksums; inputs and weights are re-derived from seeds by
lemas_tts.synthetic, exactly as tests/golden_cases.py:full_inputs does.
"""
from __future__ import annotations

import json
import sys
import tempfile
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle.gen_golden import syn, checksum, GOLDEN  # noqa: E402  (loads lemas_tts/synthetic.py standalone)
from oracle import verbatim  # noqa: E402

C3_PICK = [4, 2, 0, 27]  # utterances of c3_batch(32, seed=1): durations 1673, 951, 1165, 950 (ragged, N = 1673)

FULL_CASES = {
    "full_C1": dict(kind="raw", wseed=0, seed=0, batch=1, samples=96000, ref_frames=376, frames=940, n_text=150, steps=16,
                    cfg=2.0, sway=5.0),
    "full_C2": dict(kind="mel", wseed=0, seed=0, batch=1, ref_frames=937, frames=2187, n_text=350, steps=32, cfg=2.0,
                    sway=5.0),
    "full_C5": dict(kind="mel", wseed=0, seed=3, batch=1, ref_frames=2813, frames=2814, n_text=450, steps=64, cfg=5.0,
                    sway=3.0, edit=[1125, 1406], duration=2812),
    "full_C4_b4": dict(kind="mel", wseed=0, seed=2, batch=4, ref_frames=256, frames=768, n_text=120, steps=32, cfg=2.0,
                       sway=3.0),
    "full_C3_b4": dict(kind="c3", wseed=0, pseed=13, seed=1, batch=4, pick=C3_PICK, steps=32, cfg=2.0, sway=3.0),
}


def full_inputs(case: dict) -> dict:
    """Shared with tests/golden_cases.py (which imports nothing from oracle/: the derivation is repeated there and
    checked through `in_sums`)."""
    import dataclasses

    if case["kind"] == "c3":
        arch = dataclasses.replace(syn.FULL_ARCH, use_prosody_encoder=True)
        b = syn.c3_batch(32, seed=case["seed"])
        idx = torch.tensor(case["pick"])
        lens, dur = b["lens"][idx], b["duration"][idx]
        text = b["text"][idx]
        text = text[:, : int((text >= 0).sum(1).max())]
        cond = b["audio"][idx]
        noise = syn.synthetic_noise(dur.tolist(), arch.mel_dim, seed=case["seed"])
        return dict(arch=arch, cond=cond, text=text, lens=lens, duration=dur, noise=noise, edit_mask=None,
                    durations=dur.tolist())
    arch = syn.FULL_ARCH
    B, Tc, N = case["batch"], case["ref_frames"], case["frames"]
    if case["kind"] == "raw":   # raw reference audio: the mel front-end (modules.py:75-101) is part of the call
        cond = syn.synthetic_ref_audio(B, case["samples"], seed=case["seed"])
    else:
        cond = syn.synthetic_ref_mel(B, Tc, arch.mel_dim, seed=case["seed"])
    text = syn.synthetic_text_ids(B, case["n_text"], arch.text_num_embeds, seed=case["seed"])
    noise = syn.synthetic_noise([N] * B, arch.mel_dim, seed=case["seed"])
    edit_mask = None
    if case.get("edit"):
        edit_mask = torch.ones(B, Tc, dtype=torch.bool)
        edit_mask[:, case["edit"][0]: case["edit"][1]] = False
    return dict(arch=arch, cond=cond, text=text, lens=None, duration=case.get("duration", N), noise=noise,
                edit_mask=edit_mask, durations=[N] * B)


def run(case: dict, threads: int) -> dict:
    torch.set_num_threads(threads)
    inp = full_inputs(case)
    arch = inp["arch"]
    sd = dict(syn.make_dit_state_dict(arch, seed=case["wseed"]))
    pros = None
    tmp = None
    if case["kind"] == "c3":
        tmp = tempfile.TemporaryDirectory()
        pros = syn.write_prosody_assets(Path(tmp.name), syn.PROSODY_CFG, seed=case["pseed"])
        sd.update({"prosody_encoder.encoder." + k: v for k, v in
                   syn.make_prosody_state_dict(syn.PROSODY_CFG, case["pseed"]).items()})
    model = verbatim.build_reference_cfm(arch, sd, prosody_paths=pros)
    durs = inp["durations"]
    queue = [inp["noise"][b, : durs[b]].clone() for b in range(case["batch"])]
    real_randn = torch.randn

    def fake_randn(*size, **kw):
        want = queue.pop(0)
        assert tuple(want.shape) == tuple(int(s) for s in size), (want.shape, size)
        return want

    torch.randn = fake_randn
    t0 = time.perf_counter()
    try:
        with torch.no_grad():
            out, traj = model.sample(cond=inp["cond"], text=inp["text"], duration=inp["duration"], lens=inp["lens"],
                                     steps=case["steps"], cfg_strength=case["cfg"], sway_sampling_coef=case["sway"],
                                     edit_mask=inp["edit_mask"], use_acc_grl=False,
                                     use_prosody_encoder=case["kind"] == "c3")
    finally:
        torch.randn = real_randn
        if tmp is not None:
            tmp.cleanup()
    sec = time.perf_counter() - t0
    return dict(out=out.float().clone(),
                in_sums=torch.tensor([checksum(inp["cond"]), checksum(inp["noise"]), float(inp["text"].sum())],
                                     dtype=torch.float64),
                cpu_seconds=torch.tensor(sec), cpu_threads=torch.tensor(threads))


def main():
    names = sys.argv[1:] or list(FULL_CASES)
    GOLDEN.mkdir(parents=True, exist_ok=True)
    mpath = GOLDEN / "MANIFEST_full.json"
    manifest = json.loads(mpath.read_text()) if mpath.exists() else {
        "generator": "oracle/gen_golden_full.py", "reference": "/root/reference (verbatim import)",
        "torch": torch.__version__, "cases": {}}
    import os

    threads = int(os.environ.get("GOLDEN_THREADS", os.cpu_count() or 1))
    for name in names:
        case = FULL_CASES[name]
        res = run(case, threads)
        torch.save(res, GOLDEN / f"{name}.pt")
        manifest["cases"][name] = dict(case, reference_cpu_seconds=round(float(res["cpu_seconds"]), 1),
                                       cpu_threads=threads)
        mpath.write_text(json.dumps(manifest, indent=1))
        print(name, tuple(res["out"].shape), "out|sum|=%.6f" % checksum(res["out"]),
              "%.0f s" % float(res["cpu_seconds"]), flush=True)


if __name__ == "__main__":
    main()
