"""Vocos (mel -> 24 kHz waveform) behind the `vocos.Vocos` surface the reference uses
(utils_infer.py:132-143: `Vocos.from_hparams(config.yaml)`, `load_state_dict(pytorch_model.bin)`, `.eval().to(dev)`,
then `vocoder.decode(mel[B,100,T])` at utils_infer.py:549 / speech_edit_multilingual.py:198).

`vocos` is an un-vendored, unpinned pip dependency of the reference (requirements.txt:179).  This class carries the
parameters under the `charactr/vocos-mel-24khz` checkpoint keys (backbone.*, head.*; the feature_extractor buffers are
accepted and ignored — decode never touches them) and runs `decode` in liblemas_b200.so: 7-tap conv GEMM, fused
depthwise-conv+LayerNorm, tcgen05 GEMMs with GELU / layer-scale+residual epilogues, and a hand-written 1024-point
inverse FFT + overlap-add.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _native as nv


class _Block(nn.Module):
    def __init__(self, dim, inter):
        super().__init__()
        self.dwconv = nn.Conv1d(dim, dim, kernel_size=7, padding=3, groups=dim)
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.pwconv1 = nn.Linear(dim, inter)
        self.pwconv2 = nn.Linear(inter, dim)
        self.gamma = nn.Parameter(torch.full((dim,), 1.0 / 8))  # layer_scale_init_value = 1 / num_layers


class _Backbone(nn.Module):
    def __init__(self, input_channels, dim, intermediate_dim, num_layers):
        super().__init__()
        self.embed = nn.Conv1d(input_channels, dim, kernel_size=7, padding=3)
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.convnext = nn.ModuleList([_Block(dim, intermediate_dim) for _ in range(num_layers)])
        self.final_layer_norm = nn.LayerNorm(dim, eps=1e-6)


class _ISTFT(nn.Module):
    def __init__(self, n_fft):
        super().__init__()
        self.register_buffer("window", torch.hann_window(n_fft))


class _Head(nn.Module):
    def __init__(self, dim, n_fft):
        super().__init__()
        self.out = nn.Linear(dim, n_fft + 2)
        self.istft = _ISTFT(n_fft)


class Vocos(nn.Module):
    """charactr/vocos-mel-24khz: VocosBackbone(100, 512, 1536, 8) + ISTFTHead(512, n_fft 1024, hop 256, center)."""

    def __init__(self, input_channels=100, dim=512, intermediate_dim=1536, num_layers=8, n_fft=1024, hop_length=256,
                 padding="center"):
        super().__init__()
        if n_fft != 1024 or hop_length != 256 or padding != "center":
            raise ValueError("lemas_b200: the native ISTFT head is built for n_fft=1024, hop_length=256, "
                             "padding='center' (charactr/vocos-mel-24khz)")
        self.backbone = _Backbone(input_channels, dim, intermediate_dim, num_layers)
        self.head = _Head(dim, n_fft)
        self.hparams = dict(input_channels=input_channels, dim=dim, intermediate_dim=intermediate_dim,
                            num_layers=num_layers, n_fft=n_fft, hop_length=hop_length, padding=padding)
        self._engine = None
        self._engine_key = None

    @classmethod
    def from_hparams(cls, config_path: str) -> "Vocos":
        """Reads the `backbone` / `head` init_args of vocos' config.yaml (feature_extractor is not needed to decode)."""
        import yaml

        with open(config_path, "r") as f:
            cfg = yaml.safe_load(f)
        bb = (cfg.get("backbone") or {}).get("init_args", {})
        hd = (cfg.get("head") or {}).get("init_args", {})
        cls_path = (cfg.get("backbone") or {}).get("class_path", "vocos.models.VocosBackbone")
        if not cls_path.endswith("VocosBackbone") or not (cfg.get("head") or {}).get(
                "class_path", "vocos.heads.ISTFTHead").endswith("ISTFTHead"):
            raise ValueError(f"lemas_b200: unsupported vocos architecture {cls_path}")
        return cls(input_channels=bb.get("input_channels", 100), dim=bb.get("dim", 512),
                   intermediate_dim=bb.get("intermediate_dim", 1536), num_layers=bb.get("num_layers", 8),
                   n_fft=hd.get("n_fft", 1024), hop_length=hd.get("hop_length", 256), padding=hd.get("padding", "center"))

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        # pytorch_model.bin also carries the mel feature extractor's buffers; decode does not use them
        sd = {k: v for k, v in state_dict.items() if not k.startswith("feature_extractor.")}
        self._engine = None
        return super().load_state_dict(sd, strict=strict, **kw)

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def engine(self):
        w = self.head.out.weight
        key = nv.weights_key(self)
        if self._engine is None or self._engine_key != key:
            if w.device.type != "cuda":
                raise RuntimeError("CUDA error: the lemas_tts B200 build has no CPU path; move the vocoder to a "
                                   "Blackwell device (no kernel image is available for execution on the device)")
            from .engine import VocosEngine

            self._engine = VocosEngine(dict(self.state_dict()), device=w.device)
            self._engine_key = key
        return self._engine

    @torch.no_grad()
    def decode(self, features_input: torch.Tensor, **kwargs) -> torch.Tensor:
        """mel [B, 100, T] -> waveform [B, (T-1)*256] fp32."""
        return self.engine().decode(features_input)

    def forward(self, features_input: torch.Tensor, **kwargs) -> torch.Tensor:
        return self.decode(features_input)
