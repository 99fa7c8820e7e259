"""Runs one of the reference's command-line scripts UNCHANGED against this repo's `lemas_tts` package.

    python tests/ref_script_runner.py <script.py from the reference> <script args ...>

The script file is copied next to a mirror of the package (a directory of symlinks to lemas-tts_b200/lemas_tts plus
`scripts/<script>`), because the reference scripts put `Path(__file__).parents[2]` first on sys.path
(scripts/tts_multilingual.py:24-27) and expect `lemas_tts` there.  Host-side libraries that are absent offline are
shimmed — none of them is on the hot path:
  * soundfile / cached_path        import-only (WAV writing goes through lemas_tts.infer.utils_infer.save_audio)
  * torchaudio.load / .save        need torchcodec in this torchaudio build: replaced by the package's WAV reader/writer
  * lemas_tts.infer.frontend       the espeak / jieba text frontend (out of scope): a deterministic
                                   stand-in with the same TextNorm(dtype).text2phn / text2norm interface
"""
import os
import runpy
import shutil
import sys
import tempfile
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "lemas-tts_b200" / "lemas_tts"


def main():
    script = Path(sys.argv[1]).resolve()
    mirror = Path(tempfile.mkdtemp(prefix="lemas_mirror_"))
    pkg = mirror / "lemas_tts"
    pkg.mkdir()
    for entry in PKG.iterdir():
        if entry.name not in ("__pycache__", "scripts"):
            os.symlink(entry, pkg / entry.name)
    (pkg / "scripts").mkdir()
    target = pkg / "scripts" / script.name
    shutil.copyfile(script, target)

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    sys.path.insert(0, str(mirror))
    import torch
    import torchaudio

    from lemas_tts.infer import utils_infer

    def sf_write(path, data, sr, *a, **k):
        utils_infer.save_audio(path, torch.as_tensor(data), sr)

    stub("soundfile", write=sf_write)
    stub("cached_path", cached_path=lambda p, *a, **k: (_ for _ in ()).throw(FileNotFoundError(p)))
    torchaudio.load = lambda path, *a, **k: utils_infer.load_audio(path)
    torchaudio.save = lambda path, wav, sr, *a, **k: utils_infer.save_audio(path, wav, sr)

    class TextNorm:
        """Deterministic stand-in for lemas_tts/infer/frontend.py:18-251 (same surface: dtype, text2phn, text2norm)."""

        def __init__(self, dtype="phone"):
            self.dtype = dtype

        def text2phn(self, text):
            toks = ["(en)"]
            for ch in text:
                toks.append("_" if ch.isspace() else ("." if ch in ".!?" else f"p{ord(ch) % 700 + 1}"))
            return "|".join(toks)

        def text2norm(self, text):
            return "en", text.strip()

    stub("lemas_tts.infer.frontend", TextNorm=TextNorm)
    sys.argv = [str(target)] + sys.argv[2:]
    try:
        runpy.run_path(str(target), run_name="__main__")
    finally:
        shutil.rmtree(mirror, ignore_errors=True)


if __name__ == "__main__":
    main()
