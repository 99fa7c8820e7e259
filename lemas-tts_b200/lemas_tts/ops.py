"""Tensor-level wrappers over the op-level C ABI (one kernel launch each).

Host plumbing only: shape checks, output allocation through torch, pointer hand-off.  Used by the
op-level parity tests and by the weight packer; the sampler itself is driven from C++ (engine.py).
"""
from __future__ import annotations

import math

import torch

from . import _native as nv

f16, f32 = torch.float16, torch.float32


def _chk(t: torch.Tensor, dtype, name: str):
    if not (t.is_cuda and t.dtype == dtype and t.is_contiguous()):
        raise ValueError(f"{name}: expected contiguous CUDA {dtype}, got {t.dtype} {t.device} contiguous={t.is_contiguous()}")


@nv.on_device
def gemm(a16: torch.Tensor, w16: torch.Tensor, *, epilogue: int, n: int | None = None, bias=None, block_n: int = 128,
         out16=None, out32=None, resid=None, gate=None, gate_bstride: int = 0, row_valid=None, seq_len: int | None = None,
         rope=None, rope_cols: int = 0, inner: int = 0, vt=None, taps: int = 1, tap_pad: int = 0,
         w_tap_stride: int = 0, group_cols: int = 0, k_per_tap: int | None = None, max_ctas: int = 0,
         tap_dilation: int = 1):
    """acc = A · Wᵀ with a fused epilogue.  a16: [rows, K] or [batches, rows, K] fp16; w16: [w_rows, ldw] fp16."""
    nv.require_device()
    _chk(a16, f16, "a16")
    _chk(w16, f16, "w16")
    if a16.dim() == 2:
        batches, rows, lda = 1, a16.shape[0], a16.shape[1]
    else:
        batches, rows, lda = a16.shape
    d = nv.GemmDesc()
    d.a, d.batches, d.rows, d.lda, d.a_cols = nv.ptr(a16), batches, rows, lda, lda
    d.w, d.w_rows, d.ldw = nv.ptr(w16), w16.shape[0], w16.shape[1]
    d.n = n if n is not None else w16.shape[0]
    d.k_per_tap = k_per_tap if k_per_tap is not None else w16.shape[1]
    d.taps, d.tap_pad, d.w_tap_stride, d.group_cols = taps, tap_pad, w_tap_stride, group_cols
    d.block_n, d.epilogue = block_n, epilogue
    d.bias = nv.ptr(bias)
    if out16 is not None:
        d.out16, d.ld16 = nv.ptr(out16), out16.shape[-1]
    if out32 is not None:
        d.out32, d.ld32 = nv.ptr(out32), out32.shape[-1]
    if resid is not None:
        d.resid, d.ldr = nv.ptr(resid), resid.shape[-1]
    d.gate, d.gate_bstride = nv.ptr(gate), gate_bstride
    d.row_valid = nv.ptr(row_valid)
    d.seq_len = seq_len if seq_len is not None else rows
    d.rope, d.rope_cols, d.inner = nv.ptr(rope), rope_cols, inner
    if vt is not None:
        d.vt, d.vt_ld = nv.ptr(vt), vt.shape[-1]
    d.max_ctas = max_ctas
    d.tap_dilation = tap_dilation
    nv.check(nv.load().lemas_gemm_f16(d, nv.stream()))


@nv.on_device
def ln_modulate(x: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, seq_len: int | None = None) -> torch.Tensor:
    """x: [rows, dim] fp32; scale/shift: [dim] or [batch, dim] fp32 -> fp16 [rows, dim]."""
    nv.require_device()
    _chk(x, f32, "x")
    rows, dim = x.shape
    out = torch.empty(rows, dim, device=x.device, dtype=f16)
    bstride = dim if scale.dim() == 2 and scale.shape[0] > 1 else 0
    nv.check(nv.load().lemas_ln_modulate(nv.ptr(x), nv.ptr(scale), nv.ptr(shift), bstride, nv.ptr(out), rows, dim,
                                         seq_len or rows, nv.stream()))
    return out


@nv.on_device
def ln_affine(x, weight, bias, eps=1e-6, want16=True, want32=False):
    nv.require_device()
    _chk(x, f32, "x")
    rows, dim = x.shape
    o16 = torch.empty(rows, dim, device=x.device, dtype=f16) if want16 else None
    o32 = torch.empty(rows, dim, device=x.device, dtype=f32) if want32 else None
    nv.check(nv.load().lemas_ln_affine(nv.ptr(x), nv.ptr(weight), nv.ptr(bias), nv.ptr(o16), nv.ptr(o32), rows, dim,
                                       eps, nv.stream()))
    return o16, o32


@nv.on_device
def attention(qk16: torch.Tensor, vt16: torch.Tensor, batch: int, seq: int, heads: int, kv_len=None) -> torch.Tensor:
    """qk16: [batch*seq, 2*heads*64] fp16 (q | k); vt16: [batch, heads, 64, vt_ld] fp16 -> [batch*seq, heads*64]."""
    nv.require_device()
    _chk(qk16, f16, "qk16")
    _chk(vt16, f16, "vt16")
    out = torch.empty(batch * seq, heads * 64, device=qk16.device, dtype=f16)
    nv.check(nv.load().lemas_attention_f16(nv.ptr(qk16), qk16.shape[-1], nv.ptr(vt16), vt16.shape[-1], nv.ptr(kv_len),
                                           nv.ptr(out), batch, seq, heads, nv.stream()))
    return out


@nv.on_device
def skinny_linear(x, w, b, act_in=False, act_out=False):
    nv.require_device()
    _chk(x, f32, "x")
    _chk(w, f32, "w")
    m, k = x.shape
    n = w.shape[0]
    y = torch.empty(m, n, device=x.device, dtype=f32)
    nv.check(nv.load().lemas_skinny_linear_f32(nv.ptr(x), nv.ptr(w), nv.ptr(b), nv.ptr(y), m, k, n, int(act_in),
                                               int(act_out), nv.stream()))
    return y


@nv.on_device
def time_sinusoid(t: torch.Tensor) -> torch.Tensor:
    nv.require_device()
    _chk(t, f32, "t")
    out = torch.empty(t.shape[0], 256, device=t.device, dtype=f32)
    nv.check(nv.load().lemas_time_sinusoid(nv.ptr(t), nv.ptr(out), t.shape[0], nv.stream()))
    return out


@nv.on_device
def cfg_euler(pred, y, x16, t: float, dt: float, cfg_strength: float, copies: int, traj=None):
    nv.require_device()
    rows, mel = y.shape[0] * y.shape[1] if y.dim() == 3 else y.shape[0], y.shape[-1]
    nv.check(nv.load().lemas_cfg_euler(nv.ptr(pred), pred.shape[-1], nv.ptr(y), nv.ptr(x16), x16.shape[-1], copies,
                                       nv.ptr(traj), rows, mel, t, dt, cfg_strength, nv.stream()))


@nv.on_device
def dwconv7_ln(x, dw_w, dw_b, ln_w, ln_b):
    """x: [b, t, dim] fp32; dw_w: [7, dim] fp32 -> fp16 [b, t, dim]."""
    nv.require_device()
    _chk(x, f32, "x")
    b, t, dim = x.shape
    out = torch.empty(b, t, dim, device=x.device, dtype=f16)
    nv.check(nv.load().lemas_dwconv7_ln(nv.ptr(x), nv.ptr(dw_w), nv.ptr(dw_b), nv.ptr(ln_w), nv.ptr(ln_b), nv.ptr(out),
                                        b, t, dim, nv.stream()))
    return out


_MEL_FB: dict = {}


def mel_filterbank(n_mels: int = 100, sample_rate: int = 24000, n_fft: int = 1024):
    """HTK mel filterbank [n_fft/2+1, n_mels], norm=None, f_min 0, f_max sr/2 — the matrix torchaudio's MelScale
    holds for the reference's MelSpectrogram (modules.py:83-93) — plus the [first, last+1) non-zero bin of each filter."""
    n_freqs = n_fft // 2 + 1
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + 0.0 / 700.0)
    m_max = 2595.0 * math.log10(1.0 + (sample_rate / 2.0) / 700.0)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = torch.max(torch.zeros(1), torch.min(down, up)).contiguous()
    nz = fb > 0
    first = torch.where(nz.any(0), nz.float().argmax(0), torch.zeros(n_mels, dtype=torch.long))
    last = torch.where(nz.any(0), n_freqs - nz.flip(0).float().argmax(0), torch.zeros(n_mels, dtype=torch.long))
    return fb, torch.stack((first, last), dim=1).to(torch.int32).contiguous()


@nv.on_device
def mel_spectrogram_1024(wav: torch.Tensor, n_mels: int = 100, sample_rate: int = 24000) -> torch.Tensor:
    """wav: fp32 [b, nw] on the GPU -> log-mel fp32 [b, n_mels, nw//256 + 1]   (modules.py:75-101, csrc/frontend.cu)."""
    nv.require_device()
    _chk(wav, f32, "wav")
    assert wav.dim() == 2
    key = (wav.device, n_mels, sample_rate)
    if key not in _MEL_FB:
        fb, rng = mel_filterbank(n_mels, sample_rate, 1024)
        _MEL_FB[key] = (fb.to(wav.device), rng.to(wav.device))
    fb, rng = _MEL_FB[key]
    b, nw = wav.shape
    mel = torch.empty(b, n_mels, nw // 256 + 1, device=wav.device, dtype=f32)
    nv.check(nv.load().lemas_mel_spectrogram_1024(nv.ptr(wav), b, nw, wav.stride(0), nv.ptr(fb), nv.ptr(rng), n_mels,
                                                  nv.ptr(mel), nv.stream()))
    return mel


@nv.on_device
def mel_spectrogram_bigvgan_1024(wav: torch.Tensor, fb_mel_by_freq: torch.Tensor) -> torch.Tensor:
    """wav: fp32 [b, nw] on the GPU; fb_mel_by_freq: Slaney filterbank [n_mels, 513] -> log-mel fp32
    [b, n_mels, (nw - 256) // 256 + 1]   (get_bigvgan_mel_spectrogram, modules.py:30-72; csrc/frontend.cu)."""
    nv.require_device()
    _chk(wav, f32, "wav")
    assert wav.dim() == 2
    n_mels = fb_mel_by_freq.shape[0]
    key = ("bigvgan", wav.device, n_mels, fb_mel_by_freq.data_ptr())
    if key not in _MEL_FB:
        fb = fb_mel_by_freq.detach().float().t().contiguous().cpu()          # [513, n_mels]
        nz = fb > 0
        first = torch.where(nz.any(0), nz.float().argmax(0), torch.zeros(n_mels, dtype=torch.long))
        last = torch.where(nz.any(0), fb.shape[0] - nz.flip(0).float().argmax(0), torch.zeros(n_mels, dtype=torch.long))
        rng = torch.stack((first, last), dim=1).to(torch.int32).contiguous()
        _MEL_FB[key] = (fb.to(wav.device), rng.to(wav.device))
    fb, rng = _MEL_FB[key]
    b, nw = wav.shape
    mel = torch.empty(b, n_mels, (nw - 256) // 256 + 1, device=wav.device, dtype=f32)
    nv.check(nv.load().lemas_mel_spectrogram_bigvgan_1024(nv.ptr(wav), b, nw, wav.stride(0), nv.ptr(fb), nv.ptr(rng), n_mels,
                                                          nv.ptr(mel), nv.stream()))
    return mel


@nv.on_device
def istft_1024(head: torch.Tensor, batch: int, t: int) -> torch.Tensor:
    """head: [batch*t, ld>=1026] fp32 rows (log-mag | phase) -> wav [batch, (t-1)*256]."""
    nv.require_device()
    _chk(head, f32, "head")
    frames = torch.empty(batch * t, 1024, device=head.device, dtype=f32)
    wav = torch.empty(batch, (t - 1) * 256, device=head.device, dtype=f32)
    nv.check(nv.load().lemas_istft_1024(nv.ptr(head), head.shape[-1], nv.ptr(frames), nv.ptr(wav), batch, t, nv.stream()))
    return wav
