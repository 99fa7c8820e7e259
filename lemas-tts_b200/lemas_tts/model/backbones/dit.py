"""DiT backbone behind the reference's constructor / forward / clear_cache surface (reference
lemas_tts/model/backbones/dit.py:105-254), executing in liblemas_b200.so.

The module tree only carries parameters under the reference's state-dict keys.  On first use after a (re)load the
weights are packed for the native engine (lemas_tts.engine.DiTEngine) and every DiT FLOP — input projection,
conv position embedding, 22 AdaLN-zero blocks, final norm and projection — runs in the sm_100a kernels.

The text embedding (dit.py:51-81: embedding + abs-pos + ConvNeXtV2/GRN blocks) runs once per `sample()` for the
conditional and unconditional copies of the text; on the device it is `lemas_text_embedding` (csrc/text.cu).  The
torch statement of the same arithmetic in `TextEmbedding.forward` is what the CPU tests compare with the oracle.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from ... import _native as nv

from ..modules import (AdaLayerNorm, ConvNeXtV2Block, ConvPositionEmbedding, DiTBlock, TimestepEmbedding,
                       precompute_freqs_cis)


class TextEmbedding(nn.Module):
    """dit.py:34-81."""

    def __init__(self, text_num_embeds, text_dim, mask_padding=True, conv_layers=0, conv_mult=2):
        super().__init__()
        self.text_embed = nn.Embedding(text_num_embeds + 1, text_dim)  # 0 is the filler token
        self.mask_padding = mask_padding
        self.extra_modeling = conv_layers > 0
        if self.extra_modeling:
            self.precompute_max_pos = 4096
            self.register_buffer("freqs_cis", precompute_freqs_cis(text_dim, self.precompute_max_pos),
                                 persistent=False)
            self.text_blocks = nn.Sequential(*[ConvNeXtV2Block(text_dim, text_dim * conv_mult)
                                               for _ in range(conv_layers)])

    @staticmethod
    def _block(blk: ConvNeXtV2Block, x: torch.Tensor) -> torch.Tensor:
        """modules.py:259-269 with GRN over the sequence dim (modules.py:231-234); fp32, no TF32 paths."""
        B, N, Dm = x.shape
        w = blk.dwconv.weight.float()[:, 0]  # [dim, 7]
        xp = F.pad(x, (0, 0, 3, 3))
        h = blk.dwconv.bias.float().expand(B, N, Dm).clone()
        for k in range(7):
            h = h + xp[:, k:k + N] * w[:, k]
        h = F.layer_norm(h, (Dm,), blk.norm.weight.float(), blk.norm.bias.float(), eps=1e-6)
        h = F.gelu(F.linear(h, blk.pwconv1.weight.float(), blk.pwconv1.bias.float()))
        gx = torch.norm(h, p=2, dim=1, keepdim=True)
        nx = gx / (gx.mean(dim=-1, keepdim=True) + 1e-6)
        h = blk.grn.gamma.float() * (h * nx) + blk.grn.beta.float() + h
        h = F.linear(h, blk.pwconv2.weight.float(), blk.pwconv2.bias.float())
        return x + h

    def forward_pair(self, text: torch.Tensor, seq_len: int):
        """Conditional and unconditional embeddings in ONE batched pass (rows [0,B) keep their ids, rows [B,2B) have
        them dropped): every op of the block stack is per-sample, so this equals two separate calls."""
        B = text.shape[0]
        if text.is_cuda:  # product path: liblemas_b200.so (csrc/text.cu); the torch code below is the CPU reference
            ids = (torch.cat((text, text), 0) + 1)[:, :seq_len]
            ids = F.pad(ids, (0, seq_len - ids.shape[1]), value=0)
            drop = torch.zeros(2 * B, dtype=torch.uint8, device=text.device)
            drop[B:] = 1
            both = self.native().embed(ids, drop)
            return both[:B], both[B:]
        both = self.forward(torch.cat((text, text), 0), seq_len, drop_rows=B)
        return both[:B].contiguous(), both[B:].contiguous()

    def native(self):
        """Weights packed for lemas_text_embedding (once per weight version / device)."""
        w = self.text_embed.weight
        key = nv.weights_key(self)
        if getattr(self, "_native", None) is None or self._native_key != key:
            from ...engine import TextEngine

            self._native = TextEngine(dict(self.state_dict()), text_dim=w.shape[1],
                                      conv_layers=len(self.text_blocks) if self.extra_modeling else 0,
                                      mask_padding=self.mask_padding, device=w.device, prefix="")
            self._native_key = key
        return self._native

    def forward(self, text: torch.Tensor, seq_len: int, drop_text: bool = False, drop_rows: int | None = None) -> torch.Tensor:
        text = (text + 1)[:, :seq_len]
        text = F.pad(text, (0, seq_len - text.shape[1]), value=0)
        text_mask = text == 0  # taken before the ids are dropped (dit.py:56-60)
        if drop_text:
            text = torch.zeros_like(text)
        elif drop_rows is not None:  # rows >= drop_rows are the unconditional copies
            text = text.clone()
            text[drop_rows:] = 0
        h = F.embedding(text, self.text_embed.weight.float())
        if self.extra_modeling:
            pos = torch.arange(seq_len, device=text.device).clamp(max=self.precompute_max_pos - 1)
            h = h + self.freqs_cis[pos].float()[None]
            if self.mask_padding:
                fill = text_mask[..., None]
                h = h.masked_fill(fill, 0.0)
                for blk in self.text_blocks:
                    h = self._block(blk, h).masked_fill(fill, 0.0)
            else:
                for blk in self.text_blocks:
                    h = self._block(blk, h)
        return h


class InputEmbedding(nn.Module):
    """dit.py:87-99 parameter names; computed by the engine (split projection + 31-tap grouped conv GEMMs)."""

    def __init__(self, mel_dim, text_dim, out_dim):
        super().__init__()
        self.proj = nn.Linear(mel_dim * 2 + text_dim, out_dim)
        self.conv_pos_embed = ConvPositionEmbedding(dim=out_dim)


class RotaryEmbedding(nn.Module):
    """x-transformers RotaryEmbedding(dim_head): only the persistent `inv_freq` buffer (a checkpoint key)."""

    def __init__(self, dim, base=10000):
        super().__init__()
        self.register_buffer("inv_freq", 1.0 / (base ** (torch.arange(0, dim, 2).float() / dim)))


class DiT(nn.Module):
    def __init__(self, *, dim, depth=8, heads=8, dim_head=64, dropout=0.1, ff_mult=4, mel_dim=100,
                 text_num_embeds=256, text_dim=None, text_mask_padding=True, qk_norm=None, conv_layers=0,
                 pe_attn_head=None, long_skip_connection=False, checkpoint_activations=False,
                 use_prosody_encoder=False):
        super().__init__()
        if dim_head != 64:
            raise ValueError("lemas_b200: the sm_100a attention kernel is built for dim_head=64 (both shipped configs)")
        if long_skip_connection:
            raise ValueError("lemas_b200: long_skip_connection is not used by any shipped config and is unsupported")
        if text_dim is None:
            text_dim = mel_dim
        self.time_embed = TimestepEmbedding(dim)
        self.text_embed = TextEmbedding(text_num_embeds, text_dim, mask_padding=text_mask_padding,
                                        conv_layers=conv_layers)
        self.use_prosody_encoder = use_prosody_encoder
        self.prosody_text_proj = nn.Linear(512, text_dim) if use_prosody_encoder else None
        self.text_cond, self.text_uncond = None, None  # text cache (dit.py:140)
        self.input_embed = InputEmbedding(mel_dim, text_dim, dim)
        self.rotary_embed = RotaryEmbedding(dim_head)
        self.dim, self.depth, self.heads, self.ff_mult = dim, depth, heads, ff_mult
        self.mel_dim, self.text_dim = mel_dim, text_dim
        self.qk_norm, self.pe_attn_head = qk_norm, pe_attn_head
        self.transformer_blocks = nn.ModuleList([DiTBlock(dim=dim, heads=heads, dim_head=dim_head, ff_mult=ff_mult,
                                                          qk_norm=qk_norm) for _ in range(depth)])
        self.long_skip_connection = None
        self.norm_out = AdaLayerNorm(dim, 2)
        self.proj_out = nn.Linear(dim, mel_dim)
        self.checkpoint_activations = checkpoint_activations  # training-time option; no effect at inference
        self._engine = None
        self._engine_key = None
        self.initialize_weights()

    def initialize_weights(self):
        """dit.py:171-181: AdaLN and output layers start at zero."""
        for block in self.transformer_blocks:
            nn.init.constant_(block.attn_norm.linear.weight, 0)
            nn.init.constant_(block.attn_norm.linear.bias, 0)
        nn.init.constant_(self.norm_out.linear.weight, 0)
        nn.init.constant_(self.norm_out.linear.bias, 0)
        nn.init.constant_(self.proj_out.weight, 0)
        nn.init.constant_(self.proj_out.bias, 0)

    def clear_cache(self):
        self.text_cond, self.text_uncond = None, None

    # ------------------------------------------------------------------------------------------------ engine
    def _apply(self, fn, *a, **k):  # .to()/.half()/.cuda() move parameters: repack on next use
        self._engine = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._engine = None
        return super().load_state_dict(*a, **k)

    def _load_from_state_dict(self, *a, **k):
        self._engine = None
        return super()._load_from_state_dict(*a, **k)

    def engine(self):
        """Pack the current parameters for liblemas_b200.so (once per weight version / device)."""
        dev = self.proj_out.weight.device
        key = nv.weights_key(self)
        if self._engine is None or self._engine_key != key:
            if dev.type != "cuda":
                raise RuntimeError("CUDA error: the lemas_tts B200 build has no CPU path; move the model to a "
                                   "Blackwell device (no kernel image is available for execution on the device)")
            from ...engine import DiTEngine

            import os
            from pathlib import Path

            sd = {k: v for k, v in self.state_dict().items()}
            # LEMAS_PACKED_CACHE=<dir>: the packed (fp16, re-laid-out) weights are written once per checkpoint content
            # and uploaded as they are on later starts instead of being re-derived from the fp32 tensors
            cache = os.environ.get("LEMAS_PACKED_CACHE")
            blob = None
            if cache and self.qk_norm is None:
                only = {k: v for k, v in sd.items() if not k.startswith("text_embed.") and "prosody" not in k}
                blob = Path(cache) / f"dit_{nv.state_fingerprint(only)}.safetensors"
            if blob is not None and blob.is_file():
                self._engine = DiTEngine.from_blob(blob, device=dev)
            else:
                self._engine = DiTEngine(sd, dim=self.dim, depth=self.depth, heads=self.heads, ff_mult=self.ff_mult,
                                         text_dim=self.text_dim, mel_dim=self.mel_dim, pe_attn_head=self.pe_attn_head,
                                         qk_norm=self.qk_norm, device=dev, prefix="")
                if blob is not None:
                    blob.parent.mkdir(parents=True, exist_ok=True)
                    tmp = blob.with_suffix(f".tmp{os.getpid()}")
                    self._engine.export_blob(tmp)
                    os.replace(tmp, blob)
            self._engine_key = key
        return self._engine

    def text_embeds(self, text, seq_len, prosody_text=None, cache=True):
        """Both cached text embeddings (dit.py:212-233), prosody projection already added."""
        if not cache or self.text_cond is None:
            tc, tu = self.text_embed.forward_pair(text, seq_len)
            if prosody_text is not None and self.use_prosody_encoder:
                pt = F.linear(prosody_text.float(), self.prosody_text_proj.weight.float(),
                              self.prosody_text_proj.bias.float())
                if pt.size(1) < seq_len:
                    pt = F.pad(pt, (0, 0, 0, seq_len - pt.size(1)))
                elif pt.size(1) > seq_len:
                    pt = pt[:, :seq_len]
                tc, tu = tc + pt, tu + pt
            tc, tu = tc.contiguous(), tu.contiguous()
            if not cache:
                return tc, tu
            self.text_cond, self.text_uncond = tc, tu
        return self.text_cond, self.text_uncond

    @torch.no_grad()
    def forward(self, x, cond, text, time, drop_audio_cond, drop_text, mask=None, cache=False, prosody_text=None):
        """dit.py:194-254.  One native forward; `time` is a 0-dim tensor (or [b] of equal values)."""
        B, N, _ = x.shape
        time = torch.as_tensor(time)
        if time.ndim == 1 and B > 1 and not bool((time == time[0]).all()):
            outs = [self.forward(x[b:b + 1], cond[b:b + 1], text[b:b + 1], time[b], drop_audio_cond, drop_text,
                                 None if mask is None else mask[b:b + 1], False,
                                 None if prosody_text is None else prosody_text[b:b + 1]) for b in range(B)]
            return torch.cat(outs, 0)
        t = float(time.reshape(-1)[0])
        tc, tu = self.text_embeds(text, N, prosody_text, cache=False)
        te = tu if drop_text else tc
        if cache:  # keep the reference's cache contract observable (dit.py:212-220)
            if drop_text and self.text_uncond is None:
                self.text_uncond = tu
            if not drop_text and self.text_cond is None:
                self.text_cond = tc
        c = torch.zeros_like(cond, dtype=torch.float32) if drop_audio_cond else cond.float()
        kv_len = None
        if mask is not None:
            kv_len = mask.sum(dim=-1).to(torch.int32).contiguous()
        pred, _ = self.engine().forward_pair(x.float().contiguous(), c.contiguous(), te, te, t, kv_len)
        return pred.to(x.dtype)
