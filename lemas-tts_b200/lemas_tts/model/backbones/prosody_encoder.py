"""Pretssel ECAPA-TDNN prosody encoder (reference lemas_tts/model/backbones/prosody_encoder.py:30-433).

Runs ONCE per utterance, before the ODE loop (cfm.py:248-265): 16 kHz kaldi fbank (80 bins) -> ECAPA-TDNN ->
L2-normalised 512-d embedding that conditions the mel (prosody_to_mel) and the text (prosody_text_proj).  On the GPU
the whole path runs in hand-written fp32 CUDA (csrc/prosody.cu through lemas_tts.prosody_native: polyphase resampler,
kaldi fbank, ECAPA-TDNN); the nn.Modules below hold the parameters under the reference checkpoint's names
(`prosody_encoder.encoder.*`, so released weights load) and keep the reference's torch arithmetic for CPU tensors and
for the padding-mask variant the inference path never uses.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import List, Optional

import torch
import torch.nn.functional as F
import torchaudio
from torch import Tensor, nn

AUDIO_SAMPLE_RATE = 16_000


def _ln_channels(norm: nn.LayerNorm, x: Tensor) -> Tensor:
    """LayerNorm over the channel dim of a [B, C, T] tensor."""
    return norm(x.transpose(1, 2)).transpose(1, 2)


class TDNNBlock(nn.Module):
    """Conv1d (same padding, dilation, groups) -> ReLU -> LayerNorm(channels, eps 1e-12)   (:135-158)."""

    def __init__(self, in_channels, out_channels, kernel_size, dilation, groups: int = 1):
        super().__init__()
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size, dilation=dilation,
                              padding=dilation * (kernel_size - 1) // 2, groups=groups)
        self.norm = nn.LayerNorm(out_channels, eps=1e-12)

    def forward(self, x: Tensor, padding_mask: Optional[Tensor] = None) -> Tensor:
        return _ln_channels(self.norm, F.relu(self.conv(x)))


class Res2NetBlock(nn.Module):
    """Channel split into `scale` groups; group i>0 goes through its own TDNN fed with x_i (+ previous output) (:161-199)."""

    def __init__(self, in_channels, out_channels, scale=8, kernel_size=3, dilation=1):
        super().__init__()
        assert in_channels % scale == 0 and out_channels % scale == 0
        self.blocks = nn.ModuleList([TDNNBlock(in_channels // scale, out_channels // scale, kernel_size, dilation)
                                     for _ in range(scale - 1)])
        self.scale = scale

    def forward(self, x: Tensor) -> Tensor:
        parts = torch.chunk(x, self.scale, dim=1)
        out, prev = [parts[0]], None
        for i in range(1, self.scale):
            prev = self.blocks[i - 1](parts[i] if i == 1 else parts[i] + prev)
            out.append(prev)
        return torch.cat(out, dim=1)


class SEBlock(nn.Module):
    """Squeeze-excitation over time: s = sigmoid(W2 relu(W1 mean_t(x))); y = s * x   (:202-226)."""

    def __init__(self, in_channels, se_channels, out_channels):
        super().__init__()
        self.conv1 = nn.Conv1d(in_channels, se_channels, 1)
        self.conv2 = nn.Conv1d(se_channels, out_channels, 1)

    def forward(self, x: Tensor, padding_mask: Optional[Tensor] = None) -> Tensor:
        if padding_mask is not None:
            m = padding_mask.unsqueeze(1)
            s = (x * m).sum(dim=2, keepdim=True) / m.sum(dim=2, keepdim=True).clamp(min=1.0)
        else:
            s = x.mean(dim=2, keepdim=True)
        return torch.sigmoid(self.conv2(F.relu(self.conv1(s)))) * x


class AttentiveStatisticsPooling(nn.Module):
    """Attention-weighted mean and std over time, with global-context (mean, std) appended to the attention input (:229-278)."""

    def __init__(self, channels, attention_channels=128, global_context=True):
        super().__init__()
        self.eps = 1e-12
        self.global_context = global_context
        self.tdnn = TDNNBlock(channels * 3 if global_context else channels, attention_channels, 1, 1)
        self.conv = nn.Conv1d(attention_channels, channels, 1)

    @staticmethod
    def _stats(x: Tensor, w: Tensor, eps: float = 1e-12):
        mean = (w * x).sum(2)
        std = torch.sqrt((w * (x - mean.unsqueeze(2)).pow(2)).sum(2).clamp(eps))
        return mean, std

    def forward(self, x: Tensor, padding_mask: Optional[Tensor] = None) -> Tensor:
        n, _, length = x.shape
        mask = padding_mask if padding_mask is not None else torch.ones(n, length, device=x.device, dtype=x.dtype)
        mask = mask.unsqueeze(1)
        attn_in = x
        if self.global_context:
            mean, std = self._stats(x, mask / mask.sum(dim=2, keepdim=True).to(x))
            attn_in = torch.cat([x, mean.unsqueeze(2).expand(-1, -1, length), std.unsqueeze(2).expand(-1, -1, length)], 1)
        attn = self.conv(torch.tanh(self.tdnn(attn_in))).masked_fill(mask == 0, float("-inf"))
        mean, std = self._stats(x, F.softmax(attn, dim=2))
        return torch.cat((mean, std), dim=1).unsqueeze(2)


class SERes2NetBlock(nn.Module):
    """TDNN(k1) -> Res2Net -> TDNN(k1) -> SE, plus (projected) residual   (:281-334)."""

    def __init__(self, in_channels, out_channels, res2net_scale=8, se_channels=128, kernel_size=1, dilation=1, groups=1):
        super().__init__()
        self.out_channels = out_channels
        self.tdnn1 = TDNNBlock(in_channels, out_channels, 1, 1, groups)
        self.res2net_block = Res2NetBlock(out_channels, out_channels, res2net_scale, kernel_size, dilation)
        self.tdnn2 = TDNNBlock(out_channels, out_channels, 1, 1, groups)
        self.se_block = SEBlock(out_channels, se_channels, out_channels)
        self.shortcut = nn.Conv1d(in_channels, out_channels, 1) if in_channels != out_channels else None

    def forward(self, x: Tensor, padding_mask: Optional[Tensor] = None) -> Tensor:
        residual = self.shortcut(x) if self.shortcut is not None else x
        y = self.se_block(self.tdnn2(self.res2net_block(self.tdnn1(x))), padding_mask=padding_mask)
        return y + residual


class ECAPA_TDNN(nn.Module):
    """(B, T, 80) fbank -> L2-normalised (B, embed_dim)   (:30-132)."""

    def __init__(self, channels: List[int], kernel_sizes: List[int], dilations: List[int], attention_channels: int,
                 res2net_scale: int, se_channels: int, global_context: bool, groups: List[int], embed_dim: int,
                 input_dim: int):
        super().__init__()
        assert len(channels) == len(kernel_sizes) == len(dilations)
        self.channels, self.embed_dim = channels, embed_dim
        self.blocks = nn.ModuleList([TDNNBlock(input_dim, channels[0], kernel_sizes[0], dilations[0], groups[0])])
        for i in range(1, len(channels) - 1):
            self.blocks.append(SERes2NetBlock(channels[i - 1], channels[i], res2net_scale=res2net_scale,
                                              se_channels=se_channels, kernel_size=kernel_sizes[i],
                                              dilation=dilations[i], groups=groups[i]))
        self.mfa = TDNNBlock(channels[-1], channels[-1], kernel_sizes[-1], dilations[-1], groups=groups[-1])
        self.asp = AttentiveStatisticsPooling(channels[-1], attention_channels=attention_channels,
                                              global_context=global_context)
        self.asp_norm = nn.LayerNorm(channels[-1] * 2, eps=1e-12)
        self.fc = nn.Conv1d(channels[-1] * 2, embed_dim, 1)

    def forward(self, x: Tensor, padding_mask: Optional[Tensor] = None) -> Tensor:
        if x.is_cuda and padding_mask is None:
            from lemas_tts import prosody_native
            return prosody_native.ecapa_encode(self, x)
        return self.forward_torch(x, padding_mask)

    def forward_torch(self, x: Tensor, padding_mask: Optional[Tensor] = None) -> Tensor:
        """The reference's arithmetic in torch ops (CPU tensors; checker of the native path in tests)."""
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):  # fp32 convolutions, like the CPU reference
            x = x.transpose(1, 2)
            feats = []
            for layer in self.blocks:
                x = layer(x, padding_mask=padding_mask)
                feats.append(x)
            x = self.mfa(torch.cat(feats[1:], dim=1))
            x = _ln_channels(self.asp_norm, self.asp(x, padding_mask=padding_mask))
            x = self.fc(x).transpose(1, 2).squeeze(1)
        return F.normalize(x, dim=-1)


def extract_fbank_16k(audio_16k: Tensor) -> Tensor:
    """80-bin kaldi FBANK of 16 kHz audio [T] or [1, T] -> [frames, 80]; clips shorter than one 25 ms window are
    tiled first (:337-361)."""
    if audio_16k.ndim == 1:
        audio_16k = audio_16k.unsqueeze(0)
    if audio_16k.is_cuda:
        from lemas_tts import prosody_native
        return prosody_native.kaldi_fbank_80(audio_16k.float())[0]
    if audio_16k.shape[-1] < 400:
        audio_16k = audio_16k.repeat(1, 400 // audio_16k.shape[-1] + 1)
    return torchaudio.compliance.kaldi.fbank(audio_16k, num_mel_bins=80, sample_frequency=AUDIO_SAMPLE_RATE)


class ProsodyEncoder(nn.Module):
    """Builds the ECAPA-TDNN from `pretssel_cfg.json` and loads `prosody_encoder_UnitY2.pt` (:364-433)."""

    def __init__(self, cfg_path: Path, ckpt_path: Path, freeze: bool = True):
        super().__init__()
        cfg = json.loads(Path(cfg_path).read_text())
        if "model" not in cfg:
            raise ValueError(f"{cfg_path} does not contain a top-level 'model' key.")
        m = cfg["model"]
        self.encoder = ECAPA_TDNN(channels=m["prosody_channels"], kernel_sizes=m["prosody_kernel_sizes"],
                                  dilations=m["prosody_dilations"], attention_channels=m["prosody_attention_channels"],
                                  res2net_scale=m["prosody_res2net_scale"], se_channels=m["prosody_se_channels"],
                                  global_context=m["prosody_global_context"], groups=m["prosody_groups"],
                                  embed_dim=m["prosody_embed_dim"], input_dim=m["input_feat_per_channel"])
        self._load_state(self.encoder, Path(ckpt_path))
        if freeze:
            for p in self.encoder.parameters():
                p.requires_grad = False

    @staticmethod
    def _load_state(model: nn.Module, ckpt_path: Path) -> None:
        state = torch.load(ckpt_path, map_location="cpu", weights_only=True)
        prefixes = ("prosody_encoder_model.", "prosody_encoder.")
        if isinstance(state, dict) and any(isinstance(k, str) and k.startswith(prefixes) for k in state):
            state = {k.replace(prefixes[0], "", 1).replace(prefixes[1], "", 1): v for k, v in state.items()
                     if k.startswith(prefixes)}
        missing, unexpected = model.load_state_dict(state, strict=False)
        if missing or unexpected:
            raise RuntimeError(f"Error loading checkpoint {ckpt_path}: missing keys={missing}, "
                               f"unexpected keys={unexpected}")

    def forward(self, fbank: Tensor, padding_mask: Optional[Tensor] = None) -> Tensor:
        return self.encoder(fbank, padding_mask=padding_mask)
