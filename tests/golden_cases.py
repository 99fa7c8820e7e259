"""Re-derive the inputs of the committed golden cases (tests/golden/MANIFEST.json) from their seeds."""
from __future__ import annotations

import json
from pathlib import Path

import torch

from lemas_tts import synthetic as syn

GOLDEN = Path(__file__).resolve().parent / "golden"
MANIFEST = json.loads((GOLDEN / "MANIFEST.json").read_text())
CASES = {c["name"]: c for c in MANIFEST["cases"]}


def load(name: str) -> dict:
    return torch.load(GOLDEN / f"{name}.pt", weights_only=True)


def inputs(case: dict):
    """Same derivation as oracle/gen_golden.py:case_inputs (checked through `in_sums`)."""
    arch = getattr(syn, case["arch"])
    B, Tc, N = case["batch"], case["ref_frames"], case["frames"]
    cond = syn.synthetic_ref_mel(B, Tc, arch.mel_dim, seed=case["seed"])
    if case.get("lens"):
        for b, l in enumerate(case["lens"]):
            cond[b, l:] = 0.0
    text = syn.synthetic_text_ids(B, case["n_text"], arch.text_num_embeds, seed=case["seed"],
                                  lengths=case.get("text_lens"))
    durations = case.get("durations") or [N] * B
    noise = syn.synthetic_noise(durations, arch.mel_dim, seed=case["seed"])
    lens = torch.tensor(case["lens"]) if case.get("lens") else None
    duration = torch.tensor(durations) if B > 1 else durations[0]
    edit_mask = None
    if case.get("edit"):
        edit_mask = torch.ones(1, Tc, dtype=torch.bool)
        edit_mask[:, case["edit"][0]: case["edit"][1]] = False
    return dict(arch=arch, cond=cond, text=text, durations=durations, noise=noise, lens=lens,
                duration=duration, edit_mask=edit_mask)


PROSODY_CASE = MANIFEST.get("prosody_case")


def prosody_inputs(case: dict):
    """Same derivation as oracle/gen_golden.py:prosody_case_inputs."""
    import dataclasses

    arch = dataclasses.replace(syn.TINY_ARCH, use_prosody_encoder=True)
    audio = syn.synthetic_ref_audio(case["batch"], case["samples"], seed=case["seed"])
    text = syn.synthetic_text_ids(case["batch"], case["n_text"], arch.text_num_embeds, seed=case["seed"],
                                  lengths=case["text_lens"])
    noise = syn.synthetic_noise(case["durations"], arch.mel_dim, seed=case["seed"])
    sd = syn.make_dit_state_dict(arch, seed=case["wseed"])
    return arch, audio, text, noise, sd


# ----------------------------------------------------------------------------------- full-size, full-NFE goldens
_MF = GOLDEN / "MANIFEST_full.json"
FULL_CASES = json.loads(_MF.read_text())["cases"] if _MF.exists() else {}


def full_inputs(case: dict) -> dict:
    """Same derivation as oracle/gen_golden_full.py:full_inputs (checked through `in_sums`)."""
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
