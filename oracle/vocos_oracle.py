"""ORACLE — test infrastructure only.  Never imported by the product path.

CPU fp32 restatement of `Vocos.decode` for the `charactr/vocos-mel-24khz` model that the reference
loads at /root/reference/lemas_tts/infer/utils_infer.py:120-143 and calls at :549 and at
scripts/speech_edit_multilingual.py:198.

PARITY UNPINNED for this file: the arithmetic lives in the pip package `vocos`
(requirements.txt:179, unpinned; not vendored under /root/reference and not installed here), and the
reference holds no test or golden vector for it.  It is restated from the published vocos 0.1.0
sources (VocosBackbone / ConvNeXtBlock / ISTFTHead with padding="center").  The inverse STFT itself is
`torch.istft`, which *is* available here and pins the hand-written irFFT / overlap-add CUDA kernels.

State-dict keys follow `pytorch_model.bin`: backbone.embed, backbone.norm, backbone.convnext.{i}.
{dwconv,norm,pwconv1,pwconv2,gamma}, backbone.final_layer_norm, head.out, head.istft.window.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def vocos_backbone(sd, mel: torch.Tensor, num_layers: int) -> torch.Tensor:
    """VocosBackbone.forward: mel [B, C, T] -> [B, T, dim]."""
    x = F.conv1d(mel, sd["backbone.embed.weight"], sd["backbone.embed.bias"], padding=3)
    dim = x.shape[1]
    x = F.layer_norm(x.transpose(1, 2), (dim,), sd["backbone.norm.weight"], sd["backbone.norm.bias"], eps=1e-6)
    for i in range(num_layers):
        p = f"backbone.convnext.{i}."
        h = F.conv1d(x.transpose(1, 2), sd[p + "dwconv.weight"], sd[p + "dwconv.bias"], padding=3, groups=dim)
        h = F.layer_norm(h.transpose(1, 2), (dim,), sd[p + "norm.weight"], sd[p + "norm.bias"], eps=1e-6)
        h = F.gelu(F.linear(h, sd[p + "pwconv1.weight"], sd[p + "pwconv1.bias"]))
        h = F.linear(h, sd[p + "pwconv2.weight"], sd[p + "pwconv2.bias"])
        x = x + sd[p + "gamma"] * h
    return F.layer_norm(x, (dim,), sd["backbone.final_layer_norm.weight"],
                        sd["backbone.final_layer_norm.bias"], eps=1e-6)


def istft_head_spectrum(sd, h: torch.Tensor) -> torch.Tensor:
    """ISTFTHead up to the complex spectrum: [B, T, dim] -> complex [B, n_fft/2+1, T]."""
    y = F.linear(h, sd["head.out.weight"], sd["head.out.bias"]).transpose(1, 2)
    logmag, phase = y.chunk(2, dim=1)
    mag = torch.clip(torch.exp(logmag), max=1e2)
    return mag * (torch.cos(phase) + 1j * torch.sin(phase))


@torch.no_grad()
def vocos_decode(sd, mel: torch.Tensor, num_layers: int = 8, n_fft: int = 1024, hop: int = 256) -> torch.Tensor:
    """Vocos.decode: mel [B, 100, T] fp32 -> wav [B, (T-1)*hop]."""
    spec = istft_head_spectrum(sd, vocos_backbone(sd, mel.float(), num_layers))
    return torch.istft(spec, n_fft, hop, n_fft, sd["head.istft.window"], center=True)
