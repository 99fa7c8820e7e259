"""Waveform-side pre / post-processing of `infer_batch_process` on the device (csrc/audio.cu):
reference-audio mono mix + RMS normalisation (utils_infer.py:487-493), the per-chunk un-scaling (:552-553) and the
linear cross-fade + clip of the chunk loop (:581-622).  No host synchronisation until the final waveform is copied out.
"""
from __future__ import annotations

import torch

from . import _native as nv

f32 = torch.float32


@nv.on_device
def prep_reference_audio(audio: torch.Tensor, target_rms: float):
    """audio: CUDA fp32 [channels, samples] -> (mono [1, samples] scaled to `target_rms` when quieter, stats [2] =
    (rms, target_rms) on the device)."""
    nv.require_device()
    assert audio.is_cuda and audio.dim() == 2
    audio = audio.to(f32).contiguous()
    ch, n = audio.shape
    lib = nv.load()
    ws = torch.empty(int(lib.lemas_audio_prep_workspace_bytes(n)), device=audio.device, dtype=torch.uint8)
    mono = torch.empty(1, n, device=audio.device, dtype=f32)
    stats = torch.empty(2, device=audio.device, dtype=f32)
    nv.check(lib.lemas_audio_prep(nv.ptr(audio), ch, n, audio.stride(0), float(target_rms), nv.ptr(mono), nv.ptr(stats),
                                  nv.ptr(ws), ws.numel(), nv.stream()))
    return mono, stats


@nv.on_device
def unscale_(wave: torch.Tensor, stats: torch.Tensor) -> torch.Tensor:
    """In place: wave * rms / target_rms when the reference audio had been scaled up (utils_infer.py:552-553)."""
    assert wave.is_cuda and wave.is_contiguous() and wave.dtype == f32
    nv.check(nv.load().lemas_audio_unscale(nv.ptr(wave), wave.numel(), nv.ptr(stats), nv.stream()))
    return wave


@nv.on_device
def cross_fade_concat(waves: list, cross_fade_duration: float, sample_rate: int = 24000, clip: float = 0.999):
    """utils_infer.py:581-622 on the device.  waves: 1-D CUDA fp32 chunks.  Returns the numpy array the reference
    returns: float64 once a cross-fade happened (numpy promotes against the float64 linspace weights), float32 for a
    single chunk or plain concatenation, clipped to +-clip."""
    import numpy as np

    if not waves:
        return None
    if len(waves) == 1 or cross_fade_duration <= 0:
        return torch.cat(waves).clamp_(-clip, clip).cpu().numpy()
    lib = nv.load()
    dev = waves[0].device
    acc = torch.empty(waves[0].numel(), device=dev, dtype=torch.float64)
    nv.check(lib.lemas_audio_crossfade(None, 0, nv.ptr(waves[0].contiguous()), waves[0].numel(), 0, nv.ptr(acc), 0.0,
                                       nv.stream()))
    for i, nxt in enumerate(waves[1:], start=1):
        nxt = nxt.contiguous()
        fade = min(int(cross_fade_duration * sample_rate), acc.numel(), nxt.numel())
        out = torch.empty(acc.numel() + nxt.numel() - fade, device=dev, dtype=torch.float64)
        nv.check(lib.lemas_audio_crossfade(nv.ptr(acc), acc.numel(), nv.ptr(nxt), nxt.numel(), fade, nv.ptr(out),
                                           clip if i == len(waves) - 1 else 0.0, nv.stream()))
        acc = out
    return acc.cpu().numpy()
