"""CFM — the flow-matching sampler behind the reference's `CFM.sample` surface (reference
lemas_tts/model/cfm.py:85-473), with the ODE loop running in liblemas_b200.so.

What stays here is the once-per-call prologue of `sample` (mel of the reference audio, masks, duration fix-up,
padding, noise, the sway-sampled t grid — cfm.py:228-339, 430-453) as device tensor plumbing.  The loop itself
(cfm.py:382-456: 2 DiT forwards per step, time-weighted CFG, clamp, Euler update) is one native call:
cond/uncond rows are co-batched, every AdaLN modulation of every step is produced before the loop, the
step-invariant half of the input projection is hoisted, and the CFG+clamp+Euler update is a fused kernel.

Training (`CFM.forward`, cfm.py:506-702) is out of scope: the reference ships no trainer.
"""
from __future__ import annotations

import math
from pathlib import Path
from typing import Callable

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn.utils.rnn import pad_sequence

from .modules import AccentClassifier, MelSpec
from .utils import default, exists, lens_to_mask, list_str_to_idx, list_str_to_tensor


def compute_sway_max(steps: int, t_start: float = 0.0, min_ratio: float | None = None, safety_factor: float = 0.5,
                     eps: float = torch.finfo(torch.float32).eps) -> float:
    """cfm.py:343-373: largest sway exponent for which the first dt^p stays above `min_ratio`."""
    assert 0.0 <= t_start < 1.0
    dt = (1.0 - t_start) / max(1, steps)
    if min_ratio is None:
        min_ratio = max(1e-9, 1e2 * float(eps))
    p_max = 11.0 if dt >= 0.9 else math.log(min_ratio) / math.log(dt)
    return max(0.0, p_max - 1.0) * float(safety_factor)


def sway_time_grid(steps: int, sway_sampling_coef, t_start: float = 0.0) -> torch.Tensor:
    """cfm.py:445-453, fp32 on the host: linspace(t_start, 1, steps+1) ** (1 + min(sway_max, coef))."""
    t = torch.linspace(t_start, 1, int(steps + 1), dtype=torch.float32)
    smax = torch.tensor(compute_sway_max(steps, t_start=t_start, min_ratio=1e-9, safety_factor=0.7),
                        dtype=torch.float32)
    if sway_sampling_coef is not None:
        return t ** (1 + min(smax, sway_sampling_coef))
    return t ** (1 + smax)


def clip_and_shuffle(mel: torch.Tensor, mel_len: int, sample_rate=24000, hop_length=256, ratio=None) -> torch.Tensor:
    """cfm.py:39-83 (accent-GRL path, `ref_ratio < 1`, batch 1): crop a random window of the reference mel [d, T]
    (ratio*T frames, or 25-75 % of T), cut it into one-second pieces, shuffle them, then append randomly chosen
    pieces until T frames are covered again.  Same python-RNG draw order as the reference, so a seeded run agrees."""
    import random as _random

    fps = int(sample_rate / hop_length)
    total = mel_len
    seg = int(total * ratio) if ratio else _random.randint(int(0.25 * total), int(0.75 * total))
    start = _random.randint(0, max(0, total - seg))
    window = mel[:, start:start + seg]
    pieces = [window[:, i:i + fps] for i in range(0, window.size(1), fps)]
    _random.shuffle(pieces)
    parts, have = list(pieces), window.size(1)
    if have < total:
        extra = 0
        while extra < total:
            pick = _random.choice(pieces)
            parts.append(pick)
            extra += pick.size(1)
    out = torch.cat(parts, dim=1)[:, :total]
    assert out.shape == mel.shape, f"shuffled_mel.shape != mel.shape: {out.shape} != {mel.shape}"
    return out


class CFM(nn.Module):
    def __init__(self, transformer: nn.Module, sigma=0.0, odeint_kwargs: dict = dict(method="euler"),
                 audio_drop_prob=0.3, text_drop_prob=0.1, num_channels=None, mel_spec_module: nn.Module | None = None,
                 mel_spec_kwargs: dict = dict(), frac_lengths_mask: tuple[float, float] = (0.7, 1.0),
                 vocab_char_map: dict[str, int] | None = None, use_ctc_loss: bool = False, use_spk_enc: bool = False,
                 use_prosody_encoder: bool = False, prosody_cfg_path: str | None = None,
                 prosody_ckpt_path: str | None = None):
        super().__init__()
        if odeint_kwargs.get("method", "euler") != "euler":
            raise ValueError("lemas_b200: the native sampler integrates with fixed-grid Euler (the only method the "
                             "reference entry points use, utils_infer.py:76)")
        if use_spk_enc:
            raise ValueError("lemas_b200: use_spk_enc is False in both shipped configs and is unsupported")
        self.frac_lengths_mask = frac_lengths_mask
        self.mel_spec = default(mel_spec_module, MelSpec(**mel_spec_kwargs))
        self.num_channels = default(num_channels, self.mel_spec.n_mel_channels)
        self.audio_drop_prob, self.text_drop_prob = audio_drop_prob, text_drop_prob
        self.transformer = transformer
        self.dim = transformer.dim
        self.sigma = sigma
        self.odeint_kwargs = odeint_kwargs
        self.vocab_char_map = vocab_char_map
        # Ragged batches (B > 1, cfm.py:336-339): skip the transformer-block rows that can never reach a valid row
        # (lemas_sample_args.flags, include/lemas_b200.h).  Valid rows of `out` are unchanged; PADDED rows then differ
        # from the reference's leaky values, so it is opt-in (LEMAS_SKIP_PADDED_ROWS=1 or set the attribute).
        import os

        self.skip_padded_rows = os.environ.get("LEMAS_SKIP_PADDED_ROWS", "0") == "1"
        # LayerNorm folded into the surrounding GEMMs (LEMAS_SAMPLE_FOLD_LAYERNORM): same parity, 64 instead of 1 440
        # norm launches per utterance, measured 1.2 % slower on C2 -> off by default
        self.fold_layernorm = os.environ.get("LEMAS_FUSED_LN", "0") == "1"
        # two-GPU latency mode: a lemas_tts.parallel.CfgSplit (this process then runs one CFG variant per step)
        self.cfg_split = None
        self.use_prosody_encoder = bool(use_prosody_encoder and prosody_cfg_path and prosody_ckpt_path)
        if self.use_prosody_encoder:
            from .backbones.prosody_encoder import ProsodyEncoder

            self.prosody_encoder = ProsodyEncoder(Path(prosody_cfg_path), Path(prosody_ckpt_path), freeze=True)
            self.prosody_to_mel = nn.Linear(512, self.num_channels)
            self.prosody_dropout = nn.Dropout(p=0.2)
        else:
            self.prosody_encoder = None
        self.use_spk_enc = False
        self.use_ctc_loss = use_ctc_loss  # training-only head; its weights are dropped at load (utils_infer.py:232-235)
        self.accent_classifier = AccentClassifier(input_dim=self.num_channels, hidden_dim=self.dim, num_accents=12)
        # dtype of what `sample` returns; load_checkpoint() sets it to what the reference would run in.
        self.register_buffer("_anchor", torch.zeros(1), persistent=False)

    @property
    def device(self):
        return self._anchor.device

    def forward(self, *a, **k):
        raise NotImplementedError("lemas_b200 is inference-only: CFM.forward is the training loss (cfm.py:506-702) and "
                                  "the reference ships no trainer")

    # ------------------------------------------------------------------------------------------------
    def _prosody_embeds(self, raw_audio: torch.Tensor, device) -> torch.Tensor:
        """cfm.py:248-262: per-sample 24k -> 16k resample, kaldi fbank, ECAPA-TDNN -> [B, 512]."""
        from lemas_tts import _native as nv
        from lemas_tts import prosody_native as pn

        nv.require_device()   # no CPU path: raises "CUDA error: ..." like every other stage of sample()
        src_sr = self.mel_spec.target_sample_rate
        raw_audio = raw_audio.to(device=device, dtype=torch.float32)
        # The reference encodes one utterance at a time over the whole (padded) row raw_audio[b]; every stage is
        # per-sample, so the rows of the batch go through the native kernels together with the same result:
        # polyphase resampler -> kaldi fbank -> ECAPA-TDNN (csrc/prosody.cu).
        audio_16k = pn.resample(raw_audio.contiguous(), src_sr, 16_000)
        return self.prosody_encoder(pn.kaldi_fbank_80(audio_16k), padding_mask=None)

    @torch.no_grad()
    def sample(self, cond, text, duration, *, lens=None, steps=32, cfg_strength=1.0, sway_sampling_coef=None,
               seed: int | None = None, max_duration=4096, vocoder: Callable | None = None, no_ref_audio=False,
               duplicate_test=False, t_inter=0.1, edit_mask=None, use_acc_grl=True, use_prosody_encoder=True,
               ref_ratio=1, noise: torch.Tensor | None = None, return_trajectory: bool = True):
        """cfm.py:206-473.  Returns (out [b, n, mel], trajectory [steps+1, b, n, mel]).

        Two keyword extensions, both default to the reference behaviour: `noise` injects y0 (the reference draws it
        on its own device, cfm.py:434, so CPU-vs-GPU parity needs the same draw), and `return_trajectory=False`
        skips materialising the (steps+1)-state stack that every shipped caller discards (utils_infer.py:531,543);
        the second return value is then just the final state [1, b, n, mel].
        """
        self.eval()
        device = self.device
        raw_audio = None
        if cond.ndim == 2:  # raw wave -> mel
            raw_audio = cond.clone()
            cond = self.mel_spec(cond.to(device)).permute(0, 2, 1)
            assert cond.shape[-1] == self.num_channels
        cond = cond.to(device=device, dtype=torch.float32)
        cond_mean = cond.mean(dim=1, keepdim=True)
        batch, cond_seq_len = cond.shape[:2]
        if not exists(lens):
            lens = torch.full((batch,), cond_seq_len, device=device, dtype=torch.long)
        lens = lens.to(device)

        prosody_mel_cond = prosody_embeds = None
        if self.prosody_encoder is not None and raw_audio is not None and use_prosody_encoder:
            prosody_embeds = self._prosody_embeds(raw_audio.to(device), device)
            prosody_mel_cond = prosody_embeds[:, None, :].expand(-1, cond_seq_len, -1)

        cond_grl = None
        if use_acc_grl:  # captured BEFORE the prosody projection is added (cfm.py:279 vs :318)
            if ref_ratio is not None and ref_ratio < 1:
                rand_mel = clip_and_shuffle(cond.permute(0, 2, 1).squeeze(0), cond.shape[1], ratio=ref_ratio)
                rand_mel = rand_mel.unsqueeze(0).permute(0, 2, 1)
                assert rand_mel.shape == cond.shape, f"Shape diff: rand_mel.shape: {rand_mel.shape}, cond.shape: {cond.shape}"
                cond_grl = rand_mel
            else:
                cond_grl = cond

        if isinstance(text, list):
            if exists(self.vocab_char_map):
                text = list_str_to_idx(text, self.vocab_char_map).to(device)
            else:
                text = list_str_to_tensor(text).to(device)
            assert text.shape[0] == batch
        text = text.to(device)

        cond_mask = lens_to_mask(lens)
        if edit_mask is not None:
            cond_mask = cond_mask & edit_mask.to(device)

        if isinstance(duration, int):
            duration = torch.full((batch,), duration, device=device, dtype=torch.long)
        duration = duration.to(device)
        duration = torch.maximum(torch.maximum((text != -1).sum(dim=-1), lens) + 1, duration)
        duration = duration.clamp(max=max_duration)
        dur_host = duration.tolist()  # the one host sync of the prologue (cfm.py:305)
        max_duration = max(dur_host)

        if duplicate_test:
            test_cond = F.pad(cond, (0, 0, cond_seq_len, max_duration - 2 * cond_seq_len), value=0.0)
        cond = F.pad(cond, (0, 0, 0, max_duration - cond_seq_len), value=0.0)
        if prosody_mel_cond is not None:  # the Linear bias leaks into the zero padding, like the reference
            prosody_mel_cond = F.pad(prosody_mel_cond, (0, 0, 0, max_duration - cond_seq_len), value=0.0)
            cond = cond + F.linear(prosody_mel_cond, self.prosody_to_mel.weight.float(),
                                   self.prosody_to_mel.bias.float())
        if no_ref_audio:
            random_cond = torch.randn_like(cond) * 0.1 + cond_mean
            cond = random_cond / random_cond.mean(dim=1, keepdim=True) * cond_mean

        cond_mask = F.pad(cond_mask, (0, max_duration - cond_mask.shape[-1]), value=False).unsqueeze(-1)
        if use_acc_grl:
            cond_grl = F.pad(cond_grl, (0, 0, 0, max_duration - cond_seq_len), value=0.0)
        step_cond = torch.where(cond_mask, cond_grl if use_acc_grl else cond, torch.zeros_like(cond)).contiguous()

        kv_len = duration.to(torch.int32).contiguous() if batch > 1 else None  # mask of cfm.py:336-339

        prosody_text_cond = None
        if prosody_embeds is not None:
            prosody_text_cond = prosody_embeds[:, None, :].expand(-1, text.shape[1], -1)

        # y0 (cfm.py:430-435)
        if noise is not None:
            y = noise.to(device=device, dtype=torch.float32)
            if y.shape != (batch, max_duration, self.num_channels):
                raise ValueError(f"noise: expected {(batch, max_duration, self.num_channels)}, got {tuple(y.shape)}")
            y = y.clone().contiguous()
        else:
            y0 = []
            for dur in dur_host:
                if exists(seed):
                    torch.manual_seed(seed)
                y0.append(torch.randn(dur, self.num_channels, device=device, dtype=torch.float32))
            y = pad_sequence(y0, padding_value=0, batch_first=True).contiguous()

        t_start = 0.0
        if duplicate_test:
            t_start = t_inter
            y = ((1 - t_start) * y + t_start * test_cond).contiguous()
            steps = int(steps * (1 - t_start))
        t = sway_time_grid(steps, sway_sampling_coef, t_start)

        tr = self.transformer
        tr.clear_cache()
        text_c, text_u = tr.text_embeds(text, max_duration, prosody_text_cond, cache=True)
        engine = tr.engine()
        traj = None
        if return_trajectory:
            traj = torch.empty(steps + 1, batch, max_duration, self.num_channels, device=device, dtype=torch.float32)
        split = self.cfg_split if cfg_strength >= 1e-5 else None
        if split is not None and split.variant == 1:  # this process runs the unconditional forward (cfm.py:403-417)
            engine.sample_loop(y, torch.zeros_like(step_cond), text_u, None, t, cfg_strength, kv_len=kv_len,
                               trajectory=traj, skip_padded_rows=self.skip_padded_rows and not duplicate_test,
                               fold_layernorm=self.fold_layernorm, split=split)
        else:
            engine.sample_loop(y, step_cond, text_c, text_u if (cfg_strength >= 1e-5 and split is None) else None, t,
                               cfg_strength, kv_len=kv_len, trajectory=traj,
                               skip_padded_rows=self.skip_padded_rows and not duplicate_test,
                               fold_layernorm=self.fold_layernorm, split=split)
        tr.clear_cache()

        out = torch.where(cond_mask, cond, y)
        if no_ref_audio:
            out_mean = out[:, cond_seq_len:, :].mean(dim=1, keepdim=True)
            out[:, cond_seq_len:, :] = out[:, cond_seq_len:, :] - (out_mean - cond_mean)
        if exists(vocoder):
            out = vocoder(out.permute(0, 2, 1))
        return out, (traj if return_trajectory else y.unsqueeze(0))
