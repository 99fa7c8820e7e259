"""CPU restatement (fp32, functional PyTorch) of the BigVGAN-v2 generator the reference can be configured with
(`mel_spec_type: bigvgan`) — TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
import this file; the product path never does.

PARITY UNPINNED.  The reference imports the generator from an un-vendored git submodule, `third_party/BigVGAN`
(NVIDIA/BigVGAN, absent from /root/reference), and loads `nvidia/bigvgan_v2_24khz_100band_256x`
(/root/reference/lemas_tts/infer/utils_infer.py:144-158: `bigvgan.BigVGAN.from_pretrained(..., use_cuda_kernel=False)`,
`remove_weight_norm()`, `.eval().to(device)`; call site utils_infer.py:550-551 `vocoder(mel)` -> [B, 1, T*256]).
There is no source, checkpoint, test or golden vector for it on this machine, so the published algorithm (BigVGAN v2,
`bigvgan.py` / `activations.py` / `alias_free_activation/torch/{act,resample,filter}.py`) is restated here from the
paper and the model card's config.json:

    num_mels 100, upsample_rates [4,4,2,2,2,2], upsample_kernel_sizes [8,8,4,4,4,4], upsample_initial_channel 1536,
    resblock "1", resblock_kernel_sizes [3,7,11], resblock_dilation_sizes [[1,3,5]]*3, activation "snakebeta",
    snake_logscale true, use_bias_at_final false, use_tanh_at_final false

  x = conv_pre(mel)                                               Conv1d(100 -> 1536, k 7, pad 3)
  for every stage i:  x = ConvTranspose1d(ch -> ch/2, k 2r, stride r, pad r/2)(x)
                      x = mean_j AMPBlock1_j(x)                   j over kernel sizes (3, 7, 11)
  x = Activation1d(SnakeBeta)(x);  x = conv_post(x) (k 7, no bias);  clamp(-1, 1)

  AMPBlock1(x): for d in (1, 3, 5):  x = x + conv2_d(act(conv1_d(act(x))))      conv1 dilated by d, conv2 dilation 1
  Activation1d = UpSample1d(2) -> SnakeBeta -> DownSample1d(2), both with a 12-tap Kaiser-windowed sinc
  (cutoff 0.25, half width 0.3), replicate padding;  SnakeBeta(x) = x + sin^2(x e^alpha) / (e^beta + 1e-9).

The mel front-end of this branch, `get_bigvgan_mel_spectrogram`, IS reference source (modules.py:30-72) and is pinned
through the verbatim import with a `librosa.filters.mel` shim (tests/test_bigvgan_cpu.py).
State-dict keys follow the published checkpoint after `remove_weight_norm()`:
conv_pre.{weight,bias}, ups.{i}.0.{weight,bias}, resblocks.{n}.convs{1,2}.{d}.{weight,bias},
resblocks.{n}.activations.{a}.act.{alpha,beta}, activation_post.act.{alpha,beta}, conv_post.weight.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def kaiser_sinc_filter1d(cutoff: float, half_width: float, kernel_size: int) -> torch.Tensor:
    """alias_free_activation/torch/filter.py: Kaiser-windowed sinc low-pass, normalised to unit DC gain."""
    even = kernel_size % 2 == 0
    half_size = kernel_size // 2
    delta_f = 4 * half_width
    A = 2.285 * (half_size - 1) * math.pi * delta_f + 7.95
    if A > 50.0:
        beta = 0.1102 * (A - 8.7)
    elif A >= 21.0:
        beta = 0.5842 * (A - 21) ** 0.4 + 0.07886 * (A - 21.0)
    else:
        beta = 0.0
    window = torch.kaiser_window(kernel_size, beta=beta, periodic=False)
    time = (torch.arange(-half_size, half_size) + 0.5) if even else (torch.arange(kernel_size) - half_size)
    if cutoff == 0:
        return torch.zeros_like(time)
    filt = 2 * cutoff * window * torch.sinc(2 * cutoff * time)
    return filt / filt.sum()


def aa_filter(ratio: int = 2, kernel_size: int = 12) -> torch.Tensor:
    """The one filter both resamplers of Activation1d use (cutoff 0.5 / ratio, half width 0.6 / ratio)."""
    return kaiser_sinc_filter1d(0.5 / ratio, 0.6 / ratio, kernel_size)


def upsample2(x: torch.Tensor, filt: torch.Tensor) -> torch.Tensor:
    """UpSample1d(ratio 2, kernel 12), x [B, C, T] -> [B, C, 2T]."""
    ratio, k = 2, filt.numel()
    C = x.shape[1]
    pad = k // ratio - 1
    pad_left = pad * ratio + (k - ratio) // 2
    pad_right = pad * ratio + (k - ratio + 1) // 2
    x = F.pad(x, (pad, pad), mode="replicate")
    x = ratio * F.conv_transpose1d(x, filt.view(1, 1, -1).expand(C, -1, -1), stride=ratio, groups=C)
    return x[..., pad_left:-pad_right]


def downsample2(x: torch.Tensor, filt: torch.Tensor) -> torch.Tensor:
    """DownSample1d(ratio 2, kernel 12) = LowPassFilter1d(stride 2), x [B, C, 2T] -> [B, C, T]."""
    k = filt.numel()
    C = x.shape[1]
    even = k % 2 == 0
    pad_left = k // 2 - int(even)
    pad_right = k // 2
    x = F.pad(x, (pad_left, pad_right), mode="replicate")
    return F.conv1d(x, filt.view(1, 1, -1).expand(C, -1, -1), stride=2, groups=C)


def snake_beta(x: torch.Tensor, alpha: torch.Tensor, beta: torch.Tensor, logscale: bool = True) -> torch.Tensor:
    a = alpha[None, :, None]
    b = beta[None, :, None]
    if logscale:
        a, b = a.exp(), b.exp()
    return x + (1.0 / (b + 1e-9)) * torch.sin(x * a).pow(2)


def activation1d(x, alpha, beta, filt, logscale=True):
    return downsample2(snake_beta(upsample2(x, filt), alpha, beta, logscale), filt)


def amp_block1(sd, prefix: str, x: torch.Tensor, kernel: int, dilations, filt, logscale=True) -> torch.Tensor:
    for i, d in enumerate(dilations):
        a1, a2 = f"{prefix}activations.{2 * i}.act.", f"{prefix}activations.{2 * i + 1}.act."
        xt = activation1d(x, sd[a1 + "alpha"], sd[a1 + "beta"], filt, logscale)
        xt = F.conv1d(xt, sd[f"{prefix}convs1.{i}.weight"], sd[f"{prefix}convs1.{i}.bias"], dilation=d,
                      padding=d * (kernel - 1) // 2)
        xt = activation1d(xt, sd[a2 + "alpha"], sd[a2 + "beta"], filt, logscale)
        xt = F.conv1d(xt, sd[f"{prefix}convs2.{i}.weight"], sd[f"{prefix}convs2.{i}.bias"], padding=(kernel - 1) // 2)
        x = xt + x
    return x


def bigvgan_forward(sd: dict, mel: torch.Tensor, upsample_rates=(4, 4, 2, 2, 2, 2),
                    upsample_kernel_sizes=(8, 8, 4, 4, 4, 4), resblock_kernel_sizes=(3, 7, 11),
                    resblock_dilation_sizes=((1, 3, 5), (1, 3, 5), (1, 3, 5)), snake_logscale=True,
                    use_tanh_at_final=False) -> torch.Tensor:
    """BigVGAN.forward: mel [B, num_mels, T] fp32 -> waveform [B, 1, T * prod(upsample_rates)]."""
    filt = aa_filter()
    nk = len(resblock_kernel_sizes)
    x = F.conv1d(mel, sd["conv_pre.weight"], sd["conv_pre.bias"], padding=3)
    for i, (r, k) in enumerate(zip(upsample_rates, upsample_kernel_sizes)):
        x = F.conv_transpose1d(x, sd[f"ups.{i}.0.weight"], sd[f"ups.{i}.0.bias"], stride=r, padding=(k - r) // 2)
        xs = None
        for j in range(nk):
            y = amp_block1(sd, f"resblocks.{i * nk + j}.", x, resblock_kernel_sizes[j], resblock_dilation_sizes[j], filt,
                           snake_logscale)
            xs = y if xs is None else xs + y
        x = xs / nk
    x = activation1d(x, sd["activation_post.act.alpha"], sd["activation_post.act.beta"], filt, snake_logscale)
    x = F.conv1d(x, sd["conv_post.weight"], sd.get("conv_post.bias"), padding=3)
    return torch.tanh(x) if use_tanh_at_final else torch.clamp(x, min=-1.0, max=1.0)


def slaney_mel_filterbank(sr: int, n_fft: int, n_mels: int, fmin: float = 0.0, fmax=None) -> torch.Tensor:
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) with its defaults (htk=False, norm='slaney'): [n_mels,
    n_fft // 2 + 1].  Restated (librosa is absent): Slaney's auditory-toolbox mel scale (linear below 1 kHz, log above),
    triangular filters normalised to unit area."""
    fmax = sr / 2.0 if fmax is None else fmax

    def hz_to_mel(f):
        f = torch.as_tensor(f, dtype=torch.float64)
        f_sp = 200.0 / 3
        mel = f / f_sp
        min_log_hz, logstep = 1000.0, math.log(6.4) / 27.0
        min_log_mel = min_log_hz / f_sp
        return torch.where(f >= min_log_hz, min_log_mel + torch.log(f.clamp_min(1e-10) / min_log_hz) / logstep, mel)

    def mel_to_hz(m):
        f_sp = 200.0 / 3
        min_log_hz, logstep = 1000.0, math.log(6.4) / 27.0
        min_log_mel = min_log_hz / f_sp
        return torch.where(m >= min_log_mel, min_log_hz * torch.exp(logstep * (m - min_log_mel)), f_sp * m)

    fftfreqs = torch.linspace(0, sr / 2.0, n_fft // 2 + 1, dtype=torch.float64)
    mel_f = mel_to_hz(torch.linspace(float(hz_to_mel(fmin)), float(hz_to_mel(fmax)), n_mels + 2, dtype=torch.float64))
    fdiff = mel_f[1:] - mel_f[:-1]
    ramps = mel_f[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    weights = torch.clamp(torch.minimum(lower, upper), min=0.0)
    enorm = 2.0 / (mel_f[2: n_mels + 2] - mel_f[:n_mels])
    return (weights * enorm[:, None]).float()
