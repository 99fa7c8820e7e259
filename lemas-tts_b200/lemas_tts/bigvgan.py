"""BigVGAN-v2 (mel -> 24 kHz waveform) behind the surface the reference's `mel_spec_type: bigvgan` branch uses
(/root/reference/lemas_tts/infer/utils_infer.py:144-158):

    vocoder = bigvgan.BigVGAN.from_pretrained(local_path | "nvidia/bigvgan_v2_24khz_100band_256x", use_cuda_kernel=False)
    vocoder.remove_weight_norm();  vocoder = vocoder.eval().to(device)
    generated_wave = vocoder(mel)                    # utils_infer.py:550-551, [B, 100, T] -> [B, 1, T * 256]

The reference imports that class from an un-vendored submodule (third_party/BigVGAN, NVIDIA/BigVGAN).  This module
carries the parameters under the published checkpoint's keys (`bigvgan_generator.pt["generator"]`, with or without
weight normalisation; the resampling-filter buffers of the anti-aliased activations are accepted and checked) and runs
the forward in liblemas_b200.so (csrc/bigvgan.cu).  No CPU path: on anything but a Blackwell device it raises.
"""
from __future__ import annotations

import json
import os

import torch
from torch import nn

from . import _native as nv

# config.json of nvidia/bigvgan_v2_24khz_100band_256x (the model the reference names)
BIGVGAN_V2_24KHZ_100BAND_256X = dict(
    num_mels=100, upsample_rates=[4, 4, 2, 2, 2, 2], upsample_kernel_sizes=[8, 8, 4, 4, 4, 4],
    upsample_initial_channel=1536, resblock="1", resblock_kernel_sizes=[3, 7, 11],
    resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]], activation="snakebeta", snake_logscale=True,
    use_bias_at_final=False, use_tanh_at_final=False, sampling_rate=24000, n_fft=1024, hop_size=256, win_size=1024,
    fmin=0, fmax=None)


class _Snake(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.alpha = nn.Parameter(torch.zeros(ch))
        self.beta = nn.Parameter(torch.zeros(ch))


class _Act(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.act = _Snake(ch)


class _AMPBlock1(nn.Module):
    def __init__(self, ch, kernel, dilations):
        super().__init__()
        self.convs1 = nn.ModuleList([nn.Conv1d(ch, ch, kernel, dilation=d, padding=d * (kernel - 1) // 2) for d in dilations])
        self.convs2 = nn.ModuleList([nn.Conv1d(ch, ch, kernel, padding=(kernel - 1) // 2) for _ in dilations])
        self.activations = nn.ModuleList([_Act(ch) for _ in range(2 * len(dilations))])


def _fold_weight_norm(sd: dict) -> dict:
    """weight = g * v / ||v|| over all dims but 0 (torch.nn.utils.weight_norm, old and parametrized key styles)."""
    out = {}
    for k, v in sd.items():
        if k.endswith(".weight_g") or k.endswith(".parametrizations.weight.original0"):
            continue
        if k.endswith(".weight_v") or k.endswith(".parametrizations.weight.original1"):
            if k.endswith(".weight_v"):
                base, gk = k[: -len(".weight_v")], k[: -len(".weight_v")] + ".weight_g"
            else:
                base = k[: -len(".parametrizations.weight.original1")]
                gk = base + ".parametrizations.weight.original0"
            g = sd[gk].float()
            vf = v.float()
            norm = vf.flatten(1).norm(dim=1).view(-1, *([1] * (vf.dim() - 1)))
            out[base + ".weight"] = g * vf / norm
        else:
            out[k] = v
    return out


class BigVGAN(nn.Module):
    def __init__(self, h: dict | None = None, use_cuda_kernel: bool = False):
        super().__init__()
        self.h = dict(BIGVGAN_V2_24KHZ_100BAND_256X if h is None else h)
        h = self.h
        ch = int(h["upsample_initial_channel"])
        self.num_kernels = len(h["resblock_kernel_sizes"])
        self.num_upsamples = len(h["upsample_rates"])
        self.conv_pre = nn.Conv1d(int(h["num_mels"]), ch, 7, padding=3)
        self.ups = nn.ModuleList()
        self.resblocks = nn.ModuleList()
        for r, k in zip(h["upsample_rates"], h["upsample_kernel_sizes"]):
            self.ups.append(nn.ModuleList([nn.ConvTranspose1d(ch, ch // 2, k, stride=r, padding=(k - r) // 2)]))
            ch //= 2
            for kk, dd in zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"]):
                self.resblocks.append(_AMPBlock1(ch, kk, dd))
        self.activation_post = _Act(ch)
        self.conv_post = nn.Conv1d(ch, 1, 7, padding=3, bias=bool(h.get("use_bias_at_final", True)))
        self.use_tanh_at_final = bool(h.get("use_tanh_at_final", True))
        self._engine = None
        self._engine_key = None

    # ------------------------------------------------------------------------------------------------ loading
    @classmethod
    def from_pretrained(cls, model_id: str, use_cuda_kernel: bool = False, cache_dir=None, **kw) -> "BigVGAN":
        """`model_id`: a local directory with config.json + bigvgan_generator.pt (the layout of the Hugging Face repo).
        There is no network here: a hub id that is not a local directory raises FileNotFoundError, like a failed download."""
        if not os.path.isdir(model_id):
            raise FileNotFoundError(f"BigVGAN checkpoint directory not found: {model_id} (download "
                                    "nvidia/bigvgan_v2_24khz_100band_256x and pass its local path)")
        with open(os.path.join(model_id, "config.json")) as f:
            h = json.load(f)
        model = cls(h, use_cuda_kernel=use_cuda_kernel)
        ckpt = torch.load(os.path.join(model_id, "bigvgan_generator.pt"), map_location="cpu", weights_only=True)
        model.load_state_dict(ckpt["generator"] if "generator" in ckpt else ckpt)
        return model

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        sd = _fold_weight_norm(dict(state_dict))
        from .engine import bigvgan_aa_filter

        filt = bigvgan_aa_filter()
        for k in [k for k in sd if k.endswith("upsample.filter") or k.endswith("downsample.lowpass.filter")]:
            if not torch.allclose(sd[k].float().reshape(-1).cpu(), filt, atol=1e-6):
                raise ValueError(f"lemas_b200: {k} is not the 12-tap Kaiser-sinc filter the native activation is built for")
            del sd[k]
        self._engine = None
        return super().load_state_dict(sd, strict=strict, **kw)

    def remove_weight_norm(self):
        """Weight normalisation is folded into plain weights when the state dict is loaded."""
        return self

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    # ------------------------------------------------------------------------------------------------ forward
    def engine(self):
        w = self.conv_pre.weight
        key = nv.weights_key(self)
        if self._engine is None or self._engine_key != key:
            if w.device.type != "cuda":
                raise RuntimeError("CUDA error: the lemas_tts B200 build has no CPU path; move the vocoder to a "
                                   "Blackwell device (no kernel image is available for execution on the device)")
            from .engine import BigVGANEngine

            self._engine = BigVGANEngine(dict(self.state_dict()), self.h, device=w.device)
            self._engine_key = key
        return self._engine

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """mel [B, num_mels, T] -> waveform [B, 1, T * prod(upsample_rates)] fp32."""
        return self.engine().decode(x)
