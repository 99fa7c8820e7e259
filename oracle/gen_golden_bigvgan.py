"""Mint tests/golden/bigvgan_mel.pt from the VERBATIM reference (build container only):

    python oracle/gen_golden_bigvgan.py

`MelSpec(mel_spec_type="bigvgan")` = get_bigvgan_mel_spectrogram (/root/reference/lemas_tts/model/modules.py:30-72) on a
seeded waveform, with `librosa.filters.mel` supplied by the Slaney filterbank restated in oracle/bigvgan_oracle.py
(librosa is absent; the restatement is cross-checked against torchaudio's implementation in tests/test_bigvgan_cpu.py).
The BigVGAN generator itself cannot be minted: its source is an un-vendored submodule (parity unpinned)."""
from __future__ import annotations

import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle.gen_golden import syn, GOLDEN  # noqa: E402
from oracle import verbatim  # noqa: E402


def main():
    verbatim.install()
    from lemas_tts.model.modules import MelSpec  # the reference's file

    wav = syn.synthetic_ref_audio(2, 24000 + 77, seed=21)
    mel = MelSpec(mel_spec_type="bigvgan")(wav)
    torch.save(dict(mel=mel.float().clone()), GOLDEN / "bigvgan_mel.pt")
    print("bigvgan_mel", tuple(mel.shape), float(mel.mean()))


if __name__ == "__main__":
    main()
