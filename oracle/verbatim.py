"""ORACLE helper — imports the *verbatim* reference modules: from /root/reference in the build container, from
the vendored copy oracle/_ref/ (oracle/make_ref.py; git-ignored, travels with the gpurun snapshot) on the GPU box.

Used by oracle/gen_golden*.py to mint tests/golden/* (build container) and by bench.py's CPU legs
(`--impl reference`, `cpu_baseline`) to time the reference's own CFM.sample.  Never imported by the product or by
the `-m gpu` tests.

The reference package's __init__ pulls in hydra/soundfile/pydub/..., none of which are installed, so
a namespace stub for `lemas_tts` is registered whose __path__ points at the reference tree, and the
un-vendored third-party modules the model files import are replaced by restatements:

* torchdiffeq.odeint            — fixed-grid Euler (torchdiffeq 0.2.4, requirements.txt:167)
* x_transformers.x_transformers — RotaryEmbedding / apply_rotary_pos_emb (>=1.31.14, requirements.txt:180)
* librosa.filters.mel — the Slaney filterbank restated in oracle/bigvgan_oracle.py (bigvgan mel front-end only)
* jieba, pypinyin — import-only stubs (pinyin paths are never taken)
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import torch

REFERENCE_ROOT = Path("/root/reference")
VENDORED_ROOT = Path(__file__).resolve().parent / "_ref"


def reference_root() -> Path | None:
    for root in (REFERENCE_ROOT, VENDORED_ROOT):
        if (root / "lemas_tts" / "model" / "cfm.py").is_file():
            return root
    return None


def available() -> bool:
    return reference_root() is not None



LAST_T_GRID = None  # the t grid of the most recent odeint call (golden generator reads it)
LAST_LOOP_SECONDS = None  # wall time of the most recent Euler loop (bench.py scales the loop, not the prologue)


def _euler_odeint(func, y0, t, **kwargs):
    global LAST_T_GRID, LAST_LOOP_SECONDS
    import time

    assert kwargs.get("method", "euler") == "euler"
    LAST_T_GRID = t.detach().clone()
    ys = [y0]
    y = y0
    t_start = time.perf_counter()
    for i in range(len(t) - 1):
        t0, t1 = t[i], t[i + 1]
        y = y + (t1 - t0) * func(t0.to(y.dtype), y)
        ys.append(y)
    LAST_LOOP_SECONDS = time.perf_counter() - t_start
    return torch.stack(ys)


class _Rotary(torch.nn.Module):
    def __init__(self, dim, base=10000):
        super().__init__()
        self.register_buffer("inv_freq", 1.0 / (base ** (torch.arange(0, dim, 2).float() / dim)))

    def forward_from_seq_len(self, seq_len):
        pos = torch.arange(seq_len, device=self.inv_freq.device).type_as(self.inv_freq)
        ang = torch.einsum("i,j->ij", pos, self.inv_freq)
        ang = torch.stack((ang, ang), dim=-1).flatten(-2)
        return ang[None], 1.0


def _rotate_pairs(x):
    x = x.unflatten(-1, (-1, 2))
    a, b = x.unbind(dim=-1)
    return torch.stack((-b, a), dim=-1).flatten(-2)


def _apply_rotary(t, freqs, scale=1):
    rot_dim, seq_len, dtype = freqs.shape[-1], t.shape[-2], t.dtype
    freqs = freqs[:, -seq_len:, :]
    if t.ndim == 4 and freqs.ndim == 3:
        freqs = freqs[:, None]
    head, tail = t[..., :rot_dim], t[..., rot_dim:]
    head = (head * freqs.cos() * scale) + (_rotate_pairs(head) * freqs.sin() * scale)
    return torch.cat((head, tail), dim=-1).type(dtype)


def install() -> None:
    """Register stubs; afterwards `from lemas_tts.model.cfm import CFM` loads the reference file."""
    root = reference_root()
    if root is None:
        raise RuntimeError("reference sources not found: neither /root/reference nor oracle/_ref "
                           "(run `python oracle/make_ref.py` in the build container)")
    if "lemas_tts" in sys.modules and getattr(sys.modules["lemas_tts"], "__verbatim__", False):
        return
    for name in [m for m in sys.modules if m == "lemas_tts" or m.startswith("lemas_tts.")]:
        del sys.modules[name]
    pkg = types.ModuleType("lemas_tts")
    pkg.__path__ = [str(root / "lemas_tts")]
    pkg.__verbatim__ = True
    sys.modules["lemas_tts"] = pkg

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    stub("torchdiffeq", odeint=_euler_odeint)
    xt = stub("x_transformers")
    xt.x_transformers = stub("x_transformers.x_transformers", RotaryEmbedding=_Rotary,
                             apply_rotary_pos_emb=_apply_rotary)
    lib = stub("librosa")
    def _librosa_mel(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, **kw):
        # librosa.filters.mel with its defaults (Slaney scale, Slaney norm), restated in oracle/bigvgan_oracle.py
        from oracle.bigvgan_oracle import slaney_mel_filterbank

        return slaney_mel_filterbank(sr, n_fft, n_mels, fmin, fmax).numpy()

    lib.filters = stub("librosa.filters", mel=_librosa_mel)
    stub("jieba")
    stub("pypinyin", lazy_pinyin=None, Style=None)


def build_reference_cfm(arch, state_dict, vocab_char_map=None, prosody_paths=None):
    """Instantiate the reference CFM(DiT(...)) and strict-load a synthetic state dict (fp32, CPU).
    prosody_paths = (pretssel_cfg.json, checkpoint) builds the reference ProsodyEncoder too."""
    install()
    from lemas_tts.model.cfm import CFM  # noqa: the reference's file
    from lemas_tts.model.backbones.dit import DiT

    kw = arch.to_kwargs()
    model = CFM(
        transformer=DiT(**kw),
        mel_spec_kwargs=dict(n_fft=1024, hop_length=256, win_length=1024, n_mel_channels=arch.mel_dim,
                             target_sample_rate=24000, mel_spec_type="vocos"),
        odeint_kwargs=dict(method="euler"),
        vocab_char_map=vocab_char_map,
        use_prosody_encoder=prosody_paths is not None,
        prosody_cfg_path=None if prosody_paths is None else str(prosody_paths[0]),
        prosody_ckpt_path=None if prosody_paths is None else str(prosody_paths[1]),
    )
    model.load_state_dict(state_dict, strict=True)
    return model.eval()
