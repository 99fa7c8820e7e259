"""Parameter containers and the mel front-end (reference lemas_tts/model/modules.py).

The reference's nn.Modules compute with stock PyTorch ops; here the modules below only HOLD parameters under the
reference's state-dict names so that `load_checkpoint` (strict) works unchanged — the arithmetic of the DiT blocks
runs in liblemas_b200.so (csrc/*.cu), driven by lemas_tts.engine.  Their `forward` is deliberately absent.

MelSpec (modules.py:75-143): CUDA waveforms go through the native STFT + mel kernel (csrc/frontend.cu, SURVEY.md §8
row a1); CPU waveforms (host-side data preparation, as in the reference) use torchaudio like the reference does.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
import torchaudio
from torch import nn


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("lemas_b200: this module only holds parameters; the computation runs in liblemas_b200.so")


# ----------------------------------------------------------------------------- mel front-end

_MEL_CACHE: dict = {}


def get_vocos_mel_spectrogram(waveform, n_fft=1024, n_mel_channels=100, target_sample_rate=24000, hop_length=256,
                              win_length=1024):
    """modules.py:75-101: |STFT| (power 1, hann, center) -> HTK mel (no norm) -> log(clamp 1e-5)."""
    if waveform.dim() == 3:
        waveform = waveform.squeeze(1)
    assert waveform.dim() == 2
    if waveform.is_cuda and (n_fft, hop_length, win_length) == (1024, 256, 1024) and waveform.shape[-1] > 512:
        from lemas_tts import ops
        return ops.mel_spectrogram_1024(waveform.float().contiguous(), n_mel_channels, target_sample_rate)
    key = (str(waveform.device), n_fft, n_mel_channels, target_sample_rate, hop_length, win_length)
    tf = _MEL_CACHE.get(key)
    if tf is None:
        tf = torchaudio.transforms.MelSpectrogram(sample_rate=target_sample_rate, n_fft=n_fft, win_length=win_length,
                                                  hop_length=hop_length, n_mels=n_mel_channels, power=1, center=True,
                                                  normalized=False, norm=None).to(waveform.device)
        _MEL_CACHE[key] = tf
    if waveform.dim() == 3:
        waveform = waveform.squeeze(1)
    assert waveform.dim() == 2
    return tf(waveform.float()).clamp(min=1e-5).log()


def slaney_mel_filterbank(sample_rate: int, n_fft: int, n_mels: int, fmin: float = 0.0, fmax=None) -> torch.Tensor:
    """What `librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)` returns with its defaults (Slaney mel scale, unit-area
    triangles): [n_mels, n_fft // 2 + 1].  librosa is not a dependency here; torchaudio implements the same filterbank."""
    fmax = sample_rate / 2.0 if fmax is None else float(fmax)
    fb = torchaudio.functional.melscale_fbanks(n_fft // 2 + 1, float(fmin), fmax, n_mels, sample_rate, norm="slaney",
                                               mel_scale="slaney")
    return fb.t().contiguous()


def get_bigvgan_mel_spectrogram(waveform, n_fft=1024, n_mel_channels=100, target_sample_rate=24000, hop_length=256,
                                win_length=1024, fmin=0, fmax=None, center=False):
    """modules.py:30-72 (the `mel_spec_type: bigvgan` front-end): reflect-pad (n_fft - hop) / 2, STFT without centering,
    sqrt(re^2 + im^2 + 1e-9), Slaney mel filterbank, log(clamp 1e-5).  Runs on the waveform's device."""
    if waveform.dim() == 3:
        waveform = waveform.squeeze(1)
    assert waveform.dim() == 2
    key = ("bigvgan", str(waveform.device), n_fft, n_mel_channels, target_sample_rate, hop_length, win_length, fmin, fmax)
    if key not in _MEL_CACHE:
        _MEL_CACHE[key] = (slaney_mel_filterbank(target_sample_rate, n_fft, n_mel_channels, fmin, fmax).to(waveform.device),
                           torch.hann_window(win_length, device=waveform.device))
    mel_basis, window = _MEL_CACHE[key]
    if (waveform.is_cuda and (n_fft, hop_length, win_length) == (1024, 256, 1024) and not center
            and waveform.shape[-1] > 384):
        from lemas_tts import ops
        return ops.mel_spectrogram_bigvgan_1024(waveform.float().contiguous(), mel_basis)
    pad = (n_fft - hop_length) // 2
    x = F.pad(waveform.float().unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)
    spec = torch.stft(x, n_fft, hop_length=hop_length, win_length=win_length, window=window, center=center,
                      pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
    mag = torch.sqrt(torch.view_as_real(spec).pow(2).sum(-1) + 1e-9)
    return torch.log(torch.clamp(torch.matmul(mel_basis, mag), min=1e-5))


class MelSpec(nn.Module):
    """modules.py:104-143."""

    def __init__(self, n_fft=1024, hop_length=256, win_length=1024, n_mel_channels=100, target_sample_rate=24_000,
                 mel_spec_type="vocos"):
        super().__init__()
        assert mel_spec_type in ["vocos", "bigvgan"], "We only support two extract mel backend: vocos or bigvgan"
        self.n_fft, self.hop_length, self.win_length = n_fft, hop_length, win_length
        self.n_mel_channels, self.target_sample_rate = n_mel_channels, target_sample_rate
        self.extractor = get_vocos_mel_spectrogram if mel_spec_type == "vocos" else get_bigvgan_mel_spectrogram
        self.register_buffer("dummy", torch.tensor(0), persistent=False)

    def forward(self, wav):
        return self.extractor(waveform=wav, n_fft=self.n_fft, n_mel_channels=self.n_mel_channels,
                              target_sample_rate=self.target_sample_rate, hop_length=self.hop_length,
                              win_length=self.win_length)


# ----------------------------------------------------------------------------- parameter holders


class TimestepEmbedding(_Holder):
    """modules.py:721-731: time_mlp = Sequential(Linear(256, dim), SiLU, Linear(dim, dim))."""

    def __init__(self, dim, freq_embed_dim=256):
        super().__init__()
        self.time_mlp = nn.Sequential(nn.Linear(freq_embed_dim, dim), nn.SiLU(), nn.Linear(dim, dim))


class GRN(_Holder):
    def __init__(self, dim):
        super().__init__()
        self.gamma = nn.Parameter(torch.zeros(1, 1, dim))
        self.beta = nn.Parameter(torch.zeros(1, 1, dim))


class ConvNeXtV2Block(_Holder):
    """modules.py:241-269 parameter names."""

    def __init__(self, dim, intermediate_dim, dilation=1):
        super().__init__()
        padding = (dilation * (7 - 1)) // 2
        self.dwconv = nn.Conv1d(dim, dim, kernel_size=7, padding=padding, groups=dim, dilation=dilation)
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.pwconv1 = nn.Linear(dim, intermediate_dim)
        self.grn = GRN(intermediate_dim)
        self.pwconv2 = nn.Linear(intermediate_dim, dim)


class ConvPositionEmbedding(_Holder):
    """modules.py:167-190: conv1d = Sequential(Conv1d g16 k31, Mish, Conv1d g16 k31, Mish)."""

    def __init__(self, dim, kernel_size=31, groups=16):
        super().__init__()
        assert kernel_size % 2 != 0
        self.conv1d = nn.Sequential(
            nn.Conv1d(dim, dim, kernel_size, groups=groups, padding=kernel_size // 2), nn.Mish(),
            nn.Conv1d(dim, dim, kernel_size, groups=groups, padding=kernel_size // 2), nn.Mish())


class AdaLayerNorm(_Holder):
    """modules.py:301-315 (6 chunks) / AdaLayerNorm_Final :322-336 (2 chunks)."""

    def __init__(self, dim, chunks=6):
        super().__init__()
        self.linear = nn.Linear(dim, dim * chunks)


class RMSNormHolder(_Holder):
    def __init__(self, dim):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))


class Attention(_Holder):
    """modules.py:360-429 parameter names: to_q, to_k, to_v, to_out = [Linear, Dropout] (+ q_norm/k_norm)."""

    def __init__(self, dim, heads, dim_head, qk_norm=None):
        super().__init__()
        inner = heads * dim_head
        self.to_q, self.to_k, self.to_v = nn.Linear(dim, inner), nn.Linear(dim, inner), nn.Linear(dim, inner)
        if qk_norm == "rms_norm":
            self.q_norm, self.k_norm = RMSNormHolder(dim_head), RMSNormHolder(dim_head)
        self.to_out = nn.ModuleList([nn.Linear(inner, dim), nn.Dropout(0.0)])


class FeedForward(_Holder):
    """modules.py:342-353: ff = Sequential(Sequential(Linear, GELU(tanh)), Dropout, Linear)."""

    def __init__(self, dim, mult):
        super().__init__()
        inner = int(dim * mult)
        self.ff = nn.Sequential(nn.Sequential(nn.Linear(dim, inner), nn.GELU(approximate="tanh")), nn.Dropout(0.0),
                                nn.Linear(inner, dim))


class DiTBlock(_Holder):
    """modules.py:610-641 parameter names."""

    def __init__(self, dim, heads, dim_head, ff_mult=4, qk_norm=None):
        super().__init__()
        self.attn_norm = AdaLayerNorm(dim, 6)
        self.attn = Attention(dim, heads, dim_head, qk_norm)
        self.ff_norm = nn.LayerNorm(dim, elementwise_affine=False, eps=1e-6)
        self.ff = FeedForward(dim, ff_mult)


class AccentClassifier(_Holder):
    """modules.py:776-787 — training-only head, kept so released checkpoints strict-load."""

    def __init__(self, input_dim, hidden_dim, num_accents):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(input_dim, hidden_dim), nn.ReLU(), nn.Dropout(0.1),
                                 nn.Linear(hidden_dim, num_accents))


def precompute_freqs_cis(dim: int, end: int, theta: float = 10000.0, theta_rescale_factor=1.0):
    """modules.py:196-207: cat(cos, sin) of outer(pos, theta^(-2j/dim)) — the text abs-pos table."""
    theta *= theta_rescale_factor ** (dim / (dim - 2))
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: (dim // 2)].float() / dim))
    ang = torch.outer(torch.arange(end), freqs).float()
    return torch.cat([ang.cos(), ang.sin()], dim=-1)
