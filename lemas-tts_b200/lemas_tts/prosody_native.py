"""Host side of the native prosody path (csrc/prosody.cu): table construction, weight packing, launches.

Reference call sites: cfm.py:248-262 (per-sample resample -> extract_fbank_16k -> ProsodyEncoder),
prosody_encoder.py:30-361.  The constant tables restate torchaudio's formulas (torchaudio.functional.resample's
sinc_interp_hann kernel; torchaudio.compliance.kaldi.fbank's povey window and mel banks) so that the CUDA kernels
reproduce the reference's features; tests/test_host_cpu.py checks them against torchaudio itself.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _native as nv

f32 = torch.float32


# ------------------------------------------------------------------------------------------------ resampling
def resample_taps(orig_sr: int, new_sr: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """torchaudio.functional.resample's polyphase kernel (sinc_interp_hann): taps [new, n_taps], width, (new, orig)."""
    g = math.gcd(int(orig_sr), int(new_sr))
    orig, new = int(orig_sr) // g, int(new_sr) // g
    base_freq = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base_freq)
    idx = torch.arange(-width, width + orig, dtype=f32)[None, None] / orig
    t = torch.arange(0, -new, -1, dtype=f32)[:, None, None] / new + idx
    t *= base_freq
    t = t.clamp_(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t *= math.pi
    scale = base_freq / orig
    kernels = torch.where(t == 0, torch.tensor(1.0, dtype=f32), t.sin() / t)
    kernels *= window * scale
    return kernels[:, 0, :].contiguous(), width, new, orig


_TAPS: dict = {}


@nv.on_device
def resample(wav: torch.Tensor, orig_sr: int, new_sr: int) -> torch.Tensor:
    """wav fp32 [b, n] on the GPU -> [b, ceil(new n / orig)]   (torchaudio.functional.resample, cfm.py:254)."""
    nv.require_device()
    assert wav.is_cuda and wav.dtype == f32 and wav.dim() == 2
    if orig_sr == new_sr:
        return wav
    key = (wav.device, orig_sr, new_sr)
    if key not in _TAPS:
        taps, width, up, down = resample_taps(orig_sr, new_sr)
        _TAPS[key] = (taps.to(wav.device), width, up, down)
    taps, width, up, down = _TAPS[key]
    b, n = wav.shape
    n_out = -(-up * n // down)
    out = torch.empty(b, n_out, device=wav.device, dtype=f32)
    wav = wav if wav.stride(1) == 1 else wav.contiguous()
    nv.check(nv.load().lemas_resample_sinc(wav.data_ptr(), b, n, wav.stride(0), nv.ptr(taps), taps.shape[1], width, up,
                                           down, nv.ptr(out), n_out, n_out, nv.stream()))
    return out


# ------------------------------------------------------------------------------------------------ kaldi fbank
def kaldi_fbank_tables(n_mels: int = 80, sample_rate: int = 16000, n_fft: int = 512, win: int = 400,
                       low_freq: float = 20.0):
    """Povey window [win], mel banks [n_mels, n_fft/2+1] (kaldi mel scale, zero Nyquist column) and the non-zero bin
    range of every bank — torchaudio.compliance.kaldi.fbank defaults (prosody_encoder.py:356-360)."""
    window = torch.hann_window(win, periodic=False, dtype=f32).pow(0.85)
    num_fft_bins = n_fft // 2
    nyquist = 0.5 * sample_rate
    high_freq = nyquist
    fft_bin_width = sample_rate / n_fft

    def mel_scalar(f):
        return 1127.0 * math.log(1.0 + f / 700.0)

    mel_low, mel_high = mel_scalar(low_freq), mel_scalar(high_freq)
    delta = (mel_high - mel_low) / (n_mels + 1)
    b = torch.arange(n_mels).unsqueeze(1)
    left = mel_low + b * delta
    center = mel_low + (b + 1.0) * delta
    right = mel_low + (b + 2.0) * delta
    mel = (1127.0 * (1.0 + fft_bin_width * torch.arange(num_fft_bins) / 700.0).log()).unsqueeze(0)
    up = (mel - left) / (center - left)
    down = (right - mel) / (right - center)
    banks = torch.max(torch.zeros(1), torch.min(up, down))
    banks = torch.nn.functional.pad(banks, (0, 1), mode="constant", value=0).to(f32).contiguous()
    nz = banks > 0
    first = nz.float().argmax(1)
    last = banks.shape[1] - nz.flip(1).float().argmax(1)
    rng = torch.stack((first, last), dim=1).to(torch.int32).contiguous()
    return window, banks, rng


_FBANK: dict = {}


@nv.on_device
def kaldi_fbank_80(wav16k: torch.Tensor) -> torch.Tensor:
    """wav fp32 [b, n] (16 kHz) on the GPU -> fp32 [b, 1 + (n-400)//160, 80]   (extract_fbank_16k)."""
    nv.require_device()
    assert wav16k.is_cuda and wav16k.dtype == f32 and wav16k.dim() == 2
    if wav16k.shape[-1] < 400:  # clips shorter than one window are tiled first (prosody_encoder.py:350-354)
        wav16k = wav16k.repeat(1, 400 // wav16k.shape[-1] + 1)
    wav16k = wav16k if wav16k.stride(1) == 1 else wav16k.contiguous()
    if wav16k.device not in _FBANK:
        _FBANK[wav16k.device] = tuple(t.to(wav16k.device) for t in kaldi_fbank_tables())
    window, banks, rng = _FBANK[wav16k.device]
    b, n = wav16k.shape
    frames = 1 + (n - 400) // 160
    out = torch.empty(b, frames, 80, device=wav16k.device, dtype=f32)
    nv.check(nv.load().lemas_kaldi_fbank_16k(wav16k.data_ptr(), b, n, wav16k.stride(0), nv.ptr(window), nv.ptr(banks),
                                             nv.ptr(rng), 80, nv.ptr(out), nv.stream()))
    return out


# ------------------------------------------------------------------------------------------------ ECAPA-TDNN
def tdnn_layout(m) -> dict:
    """TDNNBlock parameters in the layout conv_cl_kernel reads: weight [cout, cin/groups, k] -> [k, cin/groups, cout]."""
    conv = m.conv
    return dict(w=conv.weight.detach().permute(2, 1, 0).contiguous(), b=conv.bias.detach(),
                ln_w=m.norm.weight.detach(), ln_b=m.norm.bias.detach(), cin=conv.in_channels, cout=conv.out_channels,
                k=conv.kernel_size[0], dil=conv.dilation[0], groups=conv.groups)


def dense_layout(conv):
    """1x1 Conv1d as a dense layer: weight [cout, cin, 1] -> [cin, cout]."""
    return conv.weight.detach()[:, :, 0].t().contiguous(), conv.bias.detach()


class PackedEcapa:
    """fp32 device copies of the encoder's parameters in the layouts csrc/prosody.cu reads, plus the ABI structs."""

    def __init__(self, enc, device):
        self.keep = []
        self.device = device
        blocks = list(enc.blocks)
        se_blocks = blocks[1:]
        c = enc.channels[0]
        ok = (len(se_blocks) >= 1 and len(se_blocks) <= 8 and enc.asp.global_context and
              all(b.shortcut is None and b.tdnn1.conv.groups == 1 and b.tdnn2.conv.groups == 1 and
                  len(b.res2net_block.blocks) == 7 and b.out_channels == c for b in se_blocks) and
              enc.channels[-1] == c * len(se_blocks))
        if not ok:
            raise RuntimeError("lemas_b200: ECAPA-TDNN variant outside the native kernel's scope (needs equal-width "
                               "SE-Res2Net blocks with res2net_scale 8, no shortcut convs, global-context pooling)")
        w = nv.ProsodyWeights()
        w.input_dim = blocks[0].conv.in_channels
        w.channels = c
        w.n_blocks = len(se_blocks)
        w.scale = 8
        w.se_channels = se_blocks[0].se_block.conv1.out_channels
        w.att_channels = enc.asp.tdnn.conv.out_channels
        w.mfa_channels = enc.channels[-1]
        w.embed_dim = enc.embed_dim
        w.block0 = self._tdnn(blocks[0])
        arr = (nv.ProsodyBlock * len(se_blocks))()
        for i, b in enumerate(se_blocks):
            arr[i].tdnn1 = self._tdnn(b.tdnn1)
            for j, r in enumerate(b.res2net_block.blocks):
                arr[i].res2[j] = self._tdnn(r)
            arr[i].tdnn2 = self._tdnn(b.tdnn2)
            arr[i].se_w1, arr[i].se_b1 = self._dense(b.se_block.conv1)
            arr[i].se_w2, arr[i].se_b2 = self._dense(b.se_block.conv2)
        self.blocks = arr
        w.blocks = C.cast(arr, C.POINTER(nv.ProsodyBlock))
        w.mfa = self._tdnn(enc.mfa)
        w.asp_tdnn = self._tdnn(enc.asp.tdnn)
        w.asp_conv_w, w.asp_conv_b = self._dense(enc.asp.conv)
        w.asp_norm_w, w.asp_norm_b = self._t(enc.asp_norm.weight), self._t(enc.asp_norm.bias)
        w.fc_w, w.fc_b = self._dense(enc.fc)
        self.w = w

    def _t(self, t):
        t = t.detach().to(device=self.device, dtype=f32).contiguous()
        self.keep.append(t)
        return t.data_ptr()

    def _tdnn(self, m):
        s = nv.ProsodyTdnn()
        lay = tdnn_layout(m)
        s.w, s.b, s.ln_w, s.ln_b = (self._t(lay[k]) for k in ("w", "b", "ln_w", "ln_b"))
        s.cin, s.cout, s.k, s.dil, s.groups = lay["cin"], lay["cout"], lay["k"], lay["dil"], lay["groups"]
        return s

    def _dense(self, conv):
        w, b = dense_layout(conv)
        return self._t(w), self._t(b)


@nv.on_device
def ecapa_encode(enc, fbank: torch.Tensor) -> torch.Tensor:
    """ECAPA_TDNN.forward(fbank [b, t, 80], padding_mask=None) on the GPU -> [b, embed_dim]."""
    nv.require_device()
    assert fbank.is_cuda and fbank.dim() == 3
    fbank = fbank.to(f32).contiguous()
    packed = getattr(enc, "_lemas_packed", None)
    key = nv.weights_key(enc)  # replaced / reloaded encoder weights must not meet a stale packed copy
    if packed is None or packed.device != fbank.device or getattr(enc, "_lemas_packed_key", None) != key:
        packed = PackedEcapa(enc, fbank.device)
        enc._lemas_packed = packed
        enc._lemas_packed_key = key
    b, t, _ = fbank.shape
    lib = nv.load()
    ws_bytes = lib.lemas_prosody_workspace_bytes(C.byref(packed.w), b, t)
    ws = torch.empty(ws_bytes, device=fbank.device, dtype=torch.uint8)
    out = torch.empty(b, packed.w.embed_dim, device=fbank.device, dtype=f32)
    nv.check(lib.lemas_prosody_encode(C.byref(packed.w), nv.ptr(fbank), b, t, nv.ptr(out), nv.ptr(ws), ws_bytes,
                                      nv.stream()))
    return out
