"""Mint tests/golden/*.pt by running the VERBATIM reference (/root/reference, via oracle/verbatim.py).

Run in the build container only:   python oracle/gen_golden.py
Inputs and weights are NOT stored: they are re-derived from seeds by lemas_tts.synthetic (CPU
generator streams are bit-reproducible), only reference OUTPUTS are committed, plus input checksums.
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "lemas-tts_b200"))

import importlib.util  # noqa: E402

_spec = importlib.util.spec_from_file_location("lemas_synthetic", ROOT / "lemas-tts_b200/lemas_tts/synthetic.py")
syn = importlib.util.module_from_spec(_spec)
sys.modules["lemas_synthetic"] = syn
_spec.loader.exec_module(syn)

from oracle import verbatim  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"


def checksum(t: torch.Tensor) -> float:
    return float(t.double().abs().sum())


def case_inputs(case: dict):
    """Shared with tests/golden_cases.py — keep in sync (the test imports THIS function)."""
    arch = getattr(syn, case["arch"])
    B, Tc, N = case["batch"], case["ref_frames"], case["frames"]
    lens = case.get("lens")
    cond = syn.synthetic_ref_mel(B, Tc, arch.mel_dim, seed=case["seed"])
    if lens is not None:
        for b, l in enumerate(lens):
            cond[b, l:] = 0.0
    text = syn.synthetic_text_ids(B, case["n_text"], arch.text_num_embeds, seed=case["seed"],
                                  lengths=case.get("text_lens"))
    durations = case.get("durations") or [N] * B
    noise = syn.synthetic_noise(durations, arch.mel_dim, seed=case["seed"])
    return arch, cond, text, durations, noise


CASES = [
    dict(name="sample_tiny_b1", arch="TINY_ARCH", wseed=11, seed=1, batch=1, ref_frames=40, frames=97, n_text=30,
         steps=4, cfg=2.0, sway=5.0, use_acc_grl=False),
    dict(name="sample_tiny_b3_ragged", arch="TINY_ARCH", wseed=11, seed=2, batch=3, ref_frames=50, frames=131,
         n_text=40, lens=[50, 31, 44], text_lens=[40, 22, 35], durations=[131, 80, 117], steps=3, cfg=2.0,
         sway=3.0, use_acc_grl=True),
    dict(name="sample_tiny_edit", arch="TINY_ARCH", wseed=11, seed=3, batch=1, ref_frames=150, frames=151,
         n_text=60, steps=4, cfg=5.0, sway=3.0, use_acc_grl=False, edit=[60, 90]),
    dict(name="sample_tiny_nocfg", arch="TINY_ARCH", wseed=11, seed=4, batch=1, ref_frames=30, frames=64,
         n_text=20, steps=2, cfg=0.0, sway=None, use_acc_grl=True),
    dict(name="sample_full_small", arch="FULL_ARCH", wseed=0, seed=5, batch=1, ref_frames=100, frames=260,
         n_text=80, steps=2, cfg=2.0, sway=5.0, use_acc_grl=False),
]


def run_sample_case(case: dict) -> dict:
    arch, cond, text, durations, noise = case_inputs(case)
    sd = syn.make_dit_state_dict(arch, seed=case["wseed"])
    model = verbatim.build_reference_cfm(arch, sd)
    B = case["batch"]
    lens = torch.tensor(case["lens"]) if case.get("lens") else None
    duration = torch.tensor(durations) if B > 1 else durations[0]
    edit_mask = None
    if case.get("edit"):
        edit_mask = torch.ones(1, case["ref_frames"], dtype=torch.bool)
        edit_mask[:, case["edit"][0]: case["edit"][1]] = False

    # inject the noise: the reference draws y0 with torch.randn on its own device (cfm.py:434)
    queue = [noise[b, : durations[b]].clone() for b in range(B)]
    real_randn = torch.randn

    def fake_randn(*size, **kw):
        want = queue.pop(0)
        assert tuple(want.shape) == tuple(int(s) for s in size), (want.shape, size)
        return want

    torch.randn = fake_randn
    try:
        out, traj = model.sample(cond=cond, text=text, duration=duration, lens=lens, steps=case["steps"],
                                 cfg_strength=case["cfg"], sway_sampling_coef=case["sway"],
                                 edit_mask=edit_mask, use_acc_grl=case["use_acc_grl"],
                                 use_prosody_encoder=False)
    finally:
        torch.randn = real_randn

    # one extra DiT.forward (cond + uncond) at t=0.37 on the final state, to pin the backbone alone
    tr = model.transformer
    N = out.shape[1]
    mask = None
    if B > 1:
        mask = torch.arange(N)[None] < torch.tensor(durations)[:, None]
    cond_p = torch.nn.functional.pad(cond, (0, 0, 0, N - cond.shape[1]))
    tt = torch.tensor(0.37)
    with torch.no_grad():
        f_c = tr(x=traj[-1], cond=cond_p, text=text, time=tt, mask=mask, drop_audio_cond=False, drop_text=False)
        f_u = tr(x=traj[-1], cond=cond_p, text=text, time=tt, mask=mask, drop_audio_cond=True, drop_text=True)
    return dict(out=out.clone(), last=traj[-1].clone(), first_step=traj[1].clone(), fwd_cond=f_c, fwd_uncond=f_u,
                in_sums=torch.tensor([checksum(cond), checksum(noise), float(text.sum())], dtype=torch.float64))


PROSODY_CASE = dict(name="sample_tiny_prosody", wseed=17, pseed=13, seed=6, batch=2, samples=12000, n_text=24,
                    text_lens=[24, 15], lens=[47, 30], durations=[110, 75], steps=3, cfg=2.0, sway=3.0)


def prosody_case_inputs(case: dict):
    """Shared with tests/golden_cases.py."""
    import dataclasses

    arch = dataclasses.replace(syn.TINY_ARCH, use_prosody_encoder=True)
    audio = syn.synthetic_ref_audio(case["batch"], case["samples"], seed=case["seed"])
    text = syn.synthetic_text_ids(case["batch"], case["n_text"], arch.text_num_embeds, seed=case["seed"],
                                  lengths=case["text_lens"])
    noise = syn.synthetic_noise(case["durations"], arch.mel_dim, seed=case["seed"])
    return arch, audio, text, noise


def run_prosody_case(case: dict, tmp: Path) -> dict:
    """CFM.sample from RAW AUDIO with the prosody encoder on (config C3's path), both use_acc_grl settings."""
    arch, audio, text, noise = prosody_case_inputs(case)
    paths = syn.write_prosody_assets(tmp, syn.TINY_PROSODY_CFG, seed=case["pseed"])
    sd = syn.make_dit_state_dict(arch, seed=case["wseed"])
    sd.update({"prosody_encoder.encoder." + k: v for k, v in
               syn.make_prosody_state_dict(syn.TINY_PROSODY_CFG, case["pseed"]).items()})
    model = verbatim.build_reference_cfm(arch, sd, prosody_paths=paths)
    res = {}
    real_randn = torch.randn
    for grl in (False, True):
        queue = [noise[b, : case["durations"][b]].clone() for b in range(case["batch"])]
        torch.randn = lambda *size, **kw: queue.pop(0)
        try:
            out, traj = model.sample(cond=audio, text=text, duration=torch.tensor(case["durations"]),
                                     lens=torch.tensor(case["lens"]), steps=case["steps"], cfg_strength=case["cfg"],
                                     sway_sampling_coef=case["sway"], use_acc_grl=grl, use_prosody_encoder=True)
        finally:
            torch.randn = real_randn
        res[f"out_grl{int(grl)}"] = out.clone()
        res[f"last_grl{int(grl)}"] = traj[-1].clone()
    # the embeddings themselves (cfm.py:248-262), to pin the ECAPA-TDNN restatement alone
    import torchaudio
    from lemas_tts.model.backbones.prosody_encoder import extract_fbank_16k

    embeds = []
    for b in range(case["batch"]):
        a16 = torchaudio.functional.resample(audio[b:b + 1], 24000, 16000).squeeze(0)
        embeds.append(model.prosody_encoder(extract_fbank_16k(a16).unsqueeze(0))[0])
    res["embeds"] = torch.stack(embeds).detach().clone()
    res["in_sums"] = torch.tensor([checksum(audio), checksum(noise), float(text.sum())], dtype=torch.float64)
    return res


def main():
    GOLDEN.mkdir(parents=True, exist_ok=True)
    manifest = {"generator": "oracle/gen_golden.py", "reference": "/root/reference (verbatim import)",
                "torch": torch.__version__, "cases": []}
    torch.set_num_threads(8)
    for case in CASES:
        res = run_sample_case(case)
        torch.save(res, GOLDEN / f"{case['name']}.pt")
        manifest["cases"].append(case)
        print(case["name"], {k: tuple(v.shape) for k, v in res.items()}, "out|sum|=%.6f" % checksum(res["out"]))

    import tempfile

    with tempfile.TemporaryDirectory() as tmp:
        with torch.no_grad():
            res = run_prosody_case(PROSODY_CASE, Path(tmp))
    torch.save(res, GOLDEN / f"{PROSODY_CASE['name']}.pt")
    manifest["prosody_case"] = PROSODY_CASE
    print(PROSODY_CASE["name"], {k: tuple(v.shape) for k, v in res.items()})

    # mel front-end (modules.py:75-101) on synthetic audio
    verbatim.install()
    from lemas_tts.model.modules import MelSpec

    wav = syn.synthetic_ref_audio(2, 24000, seed=9)
    mel = MelSpec()(wav)
    torch.save(dict(mel=mel), GOLDEN / "melspec.pt")
    print("melspec", tuple(mel.shape))

    # time grids (cfm.py:445-453) for the step counts / coefficients the entry points use
    grids = {}
    arch = syn.TINY_ARCH
    model = verbatim.build_reference_cfm(arch, syn.make_dit_state_dict(arch, seed=11))
    cond = syn.synthetic_ref_mel(1, 8, arch.mel_dim, seed=0)
    text = syn.synthetic_text_ids(1, 4, arch.text_num_embeds, seed=0)
    for steps, coef in [(16, 5.0), (32, 5.0), (32, 3.0), (32, -1.0), (32, None), (64, 3.0), (32, 1.0), (7, 0.5)]:
        model.sample(cond=cond, text=text, duration=12, steps=steps, cfg_strength=0.0, sway_sampling_coef=coef)
        grids[f"{steps}_{coef}"] = verbatim.LAST_T_GRID
    torch.save(grids, GOLDEN / "time_grids.pt")
    print("time_grids", list(grids))
    (GOLDEN / "MANIFEST.json").write_text(json.dumps(manifest, indent=1))


if __name__ == "__main__":
    main()
