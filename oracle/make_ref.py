"""Vendor the reference's OWN implementation of the hot path into oracle/_ref/ so it can run on the GPU box.

    python oracle/make_ref.py            (build container only: reads /root/reference, which the GPU box does not have)

The reference is pure Python and cannot be pip-installed offline (no setup.py / pyproject.toml, >= 20 absent
dependencies — DESIGN.md §2).  Its hot path, however, lives in six self-contained files; this recipe copies them (and the two CLI scripts),
byte for byte, to oracle/_ref/lemas_tts/model/... together with a MANIFEST (source path + sha256).  oracle/_ref/ is
git-ignored (reference sources never enter the history) but NOT gpurun-ignored, so it travels with the snapshot like a
built .so.  oracle/verbatim.py imports the files from /root/reference when that exists and from oracle/_ref/ otherwise;
`bench.py --impl reference` and the `cpu_baseline` leg then time the reference's own `CFM.sample`
(/root/reference/lemas_tts/model/cfm.py:206-473) on the box's host cores (`cpu_baseline.kind = "reference"`).
TEST INFRASTRUCTURE ONLY — nothing under lemas-tts_b200/ imports it.
"""
from __future__ import annotations

import hashlib
import json
import shutil
from pathlib import Path

SRC = Path("/root/reference/lemas_tts")
DST = Path(__file__).resolve().parent / "_ref" / "lemas_tts"
FILES = ["model/cfm.py", "model/modules.py", "model/utils.py", "model/backbones/dit.py",
         "model/backbones/prosody_encoder.py", "model/backbones/ecapa_tdnn.py",
         # the two command-line entry points, executed UNCHANGED against this repo's package by
         # tests/test_reference_scripts_gpu.py
         "scripts/tts_multilingual.py", "scripts/speech_edit_multilingual.py"]


def make(verbose: bool = True) -> bool:
    if not SRC.is_dir():
        if verbose:
            print("oracle/make_ref.py: /root/reference absent — keeping whatever oracle/_ref/ holds")
        return False
    manifest = {"source": str(SRC), "files": {}}
    for rel in FILES:
        dst = DST / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(SRC / rel, dst)
        manifest["files"][rel] = hashlib.sha256(dst.read_bytes()).hexdigest()
    (DST.parent / "MANIFEST.json").write_text(json.dumps(manifest, indent=1))
    if verbose:
        print(f"oracle/_ref: {len(FILES)} reference files vendored (git-ignored)")
    return True


if __name__ == "__main__":
    make()
