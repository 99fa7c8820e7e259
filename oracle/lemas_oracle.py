"""ORACLE — test infrastructure only.  Never imported by the product path.

CPU fp32 restatement, in plain functional PyTorch, of the arithmetic of the LEMAS-TTS acoustic
hot path (flow-matching DiT sampler).  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import this module, and only as the checker.

Every function cites the reference file:line (relative to /root/reference) that it follows.
Third-party arithmetic that is NOT vendored under /root/reference is restated from the
published algorithm of the pinned package:

* torchdiffeq==0.2.4 `odeint(method="euler")`  (requirements.txt:167; call cfm.py:456)
* x-transformers>=1.31.14 `RotaryEmbedding` / `apply_rotary_pos_emb`
  (requirements.txt:180; calls dit.py:143,236 and modules.py:476-480)

Pinning: `tests/test_oracle_golden.py` checks this file against golden vectors produced by
importing the *verbatim* reference modules (oracle/verbatim.py + oracle/gen_golden.py, run in the
build container where /root/reference exists) — see tests/golden/MANIFEST.json.

State dicts use the reference checkpoint key layout.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------- small pieces


def time_embedding(sd, t: torch.Tensor) -> torch.Tensor:
    """TimestepEmbedding (modules.py:721-731) over SinusPositionEmbedding(256) (modules.py:149-161).

    t: [B] -> [B, dim].  sin/cos of 1000*t*w_k, w_k = exp(-k ln(1e4)/127), k<128; cat(sin, cos).
    """
    half = 128
    w = torch.exp(torch.arange(half).float() * -(math.log(10000) / (half - 1)))
    arg = 1000 * t[:, None] * w[None, :]
    h = torch.cat((arg.sin(), arg.cos()), dim=-1).to(t.dtype)
    p = "transformer.time_embed.time_mlp."
    h = F.linear(h, sd[p + "0.weight"], sd[p + "0.bias"])
    h = F.silu(h)
    return F.linear(h, sd[p + "2.weight"], sd[p + "2.bias"])


def text_abs_pos_table(text_dim: int, max_pos: int = 4096) -> torch.Tensor:
    """precompute_freqs_cis (modules.py:196-207): cat(cos, sin) of outer(pos, 1e4^(-2j/dim))."""
    inv = 1.0 / (10000.0 ** (torch.arange(0, text_dim, 2)[: text_dim // 2].float() / text_dim))
    ang = torch.outer(torch.arange(max_pos), inv).float()
    return torch.cat([ang.cos(), ang.sin()], dim=-1)


def convnext_v2_block(sd, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """ConvNeXtV2Block (modules.py:241-269) with GRN over the *sequence* dim (modules.py:225-234)."""
    h = F.conv1d(x.transpose(1, 2), sd[prefix + "dwconv.weight"], sd[prefix + "dwconv.bias"],
                 padding=3, groups=x.shape[-1]).transpose(1, 2)
    h = F.layer_norm(h, (h.shape[-1],), sd[prefix + "norm.weight"], sd[prefix + "norm.bias"], eps=1e-6)
    h = F.linear(h, sd[prefix + "pwconv1.weight"], sd[prefix + "pwconv1.bias"])
    h = F.gelu(h)
    gx = torch.norm(h, p=2, dim=1, keepdim=True)
    nx = gx / (gx.mean(dim=-1, keepdim=True) + 1e-6)
    h = sd[prefix + "grn.gamma"] * (h * nx) + sd[prefix + "grn.beta"] + h
    h = F.linear(h, sd[prefix + "pwconv2.weight"], sd[prefix + "pwconv2.bias"])
    return x + h


def text_embedding(sd, arch, text: torch.Tensor, seq_len: int, drop_text: bool) -> torch.Tensor:
    """TextEmbedding.forward (dit.py:51-81).  text: [B, nt] int64 with -1 padding -> [B, N, text_dim].

    Note the filler mask is taken *before* the ids are zeroed for the unconditional pass (dit.py:56-60).
    """
    ids = (text + 1)[:, :seq_len]
    ids = F.pad(ids, (0, seq_len - ids.shape[1]), value=0)
    filler = ids == 0
    if drop_text:
        ids = torch.zeros_like(ids)
    h = F.embedding(ids, sd["transformer.text_embed.text_embed.weight"])
    if arch.conv_layers > 0:
        table = text_abs_pos_table(arch.text_dim)
        pos = torch.arange(seq_len).clamp(max=4095)  # get_pos_embed_indices, modules.py:210-219
        h = h + table[pos][None]
        if arch.text_mask_padding:
            h = h.masked_fill(filler[..., None], 0.0)
        for i in range(arch.conv_layers):
            h = convnext_v2_block(sd, f"transformer.text_embed.text_blocks.{i}.", h)
            if arch.text_mask_padding:
                h = h.masked_fill(filler[..., None], 0.0)
    return h


def mish(x):
    return x * torch.tanh(F.softplus(x))


def input_embedding(sd, x, cond, text_embed, drop_audio_cond: bool) -> torch.Tensor:
    """InputEmbedding.forward (dit.py:87-99) + ConvPositionEmbedding (modules.py:167-190, no mask)."""
    if drop_audio_cond:
        cond = torch.zeros_like(cond)
    p = "transformer.input_embed."
    h = F.linear(torch.cat((x, cond, text_embed), dim=-1), sd[p + "proj.weight"], sd[p + "proj.bias"])
    c = h.transpose(1, 2)
    for j in (0, 2):
        c = mish(F.conv1d(c, sd[f"{p}conv_pos_embed.conv1d.{j}.weight"],
                          sd[f"{p}conv_pos_embed.conv1d.{j}.bias"], padding=15, groups=16))
    return c.transpose(1, 2) + h


def rotary_table(inv_freq: torch.Tensor, seq_len: int) -> torch.Tensor:
    """RotaryEmbedding.forward_from_seq_len (x-transformers): angles[n, 2j] = angles[n, 2j+1] = n*inv_freq[j]."""
    ang = torch.outer(torch.arange(seq_len).to(inv_freq.dtype), inv_freq)
    return torch.repeat_interleave(ang, 2, dim=-1)  # [N, dim_head]


def apply_rotary(t: torch.Tensor, ang: torch.Tensor) -> torch.Tensor:
    """apply_rotary_pos_emb (x-transformers): t*cos + rotate_half(t)*sin on adjacent pairs (-x1, x0)."""
    pairs = t.unflatten(-1, (-1, 2))
    rot = torch.stack((-pairs[..., 1], pairs[..., 0]), dim=-1).flatten(-2)
    return t * ang.cos() + rot * ang.sin()


def rms_norm_head(x, weight, eps=1e-6):
    """RMSNorm on dim_head (modules.py:275-294), only with qk_norm='rms_norm'."""
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * weight


def attention(sd, arch, prefix: str, x: torch.Tensor, mask, ang) -> torch.Tensor:
    """Attention + AttnProcessor.__call__ (modules.py:442-503)."""
    B, N, _ = x.shape
    H, dh = arch.heads, arch.dim_head
    q = F.linear(x, sd[prefix + "to_q.weight"], sd[prefix + "to_q.bias"]).view(B, N, H, dh).transpose(1, 2)
    k = F.linear(x, sd[prefix + "to_k.weight"], sd[prefix + "to_k.bias"]).view(B, N, H, dh).transpose(1, 2)
    v = F.linear(x, sd[prefix + "to_v.weight"], sd[prefix + "to_v.bias"]).view(B, N, H, dh).transpose(1, 2)
    if arch.qk_norm == "rms_norm":
        q = rms_norm_head(q, sd[prefix + "q_norm.weight"])
        k = rms_norm_head(k, sd[prefix + "k_norm.weight"])
    pn = H if arch.pe_attn_head is None else arch.pe_attn_head
    q = torch.cat((apply_rotary(q[:, :pn], ang), q[:, pn:]), dim=1)
    k = torch.cat((apply_rotary(k[:, :pn], ang), k[:, pn:]), dim=1)
    s = (q @ k.transpose(-1, -2)) * (1.0 / math.sqrt(dh))
    if mask is not None:
        s = s.masked_fill(~mask[:, None, None, :], float("-inf"))
    o = (s.softmax(dim=-1) @ v).transpose(1, 2).reshape(B, N, H * dh)
    o = F.linear(o, sd[prefix + "to_out.0.weight"], sd[prefix + "to_out.0.bias"])
    if mask is not None:
        o = o.masked_fill(~mask[..., None], 0.0)
    return o


def layer_norm_plain(x):
    return F.layer_norm(x, (x.shape[-1],), eps=1e-6)


def dit_block(sd, arch, i: int, x, t_emb, mask, ang):
    """DiTBlock.forward (modules.py:627-641) with AdaLayerNorm (modules.py:301-315)."""
    p = f"transformer.transformer_blocks.{i}."
    emb = F.linear(F.silu(t_emb), sd[p + "attn_norm.linear.weight"], sd[p + "attn_norm.linear.bias"])
    shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = emb.chunk(6, dim=1)
    n1 = layer_norm_plain(x) * (1 + scale_msa[:, None]) + shift_msa[:, None]
    x = x + gate_msa[:, None] * attention(sd, arch, p + "attn.", n1, mask, ang)
    n2 = layer_norm_plain(x) * (1 + scale_mlp[:, None]) + shift_mlp[:, None]
    h = F.linear(n2, sd[p + "ff.ff.0.0.weight"], sd[p + "ff.ff.0.0.bias"])
    h = F.gelu(h, approximate="tanh")
    h = F.linear(h, sd[p + "ff.ff.2.weight"], sd[p + "ff.ff.2.bias"])
    return x + gate_mlp[:, None] * h


def dit_forward(sd, arch, x, cond, text_embed, time, mask=None, drop_audio_cond=False,
                prosody_text=None, return_hidden=False):
    """DiT.forward (dit.py:194-254) given an already computed text embedding (the reference caches it).

    x, cond: [B, N, mel]; text_embed: [B, N, text_dim]; time: 0-dim or [B]; mask: [B, N] bool | None.
    """
    B, N, _ = x.shape
    if time.ndim == 0:
        time = time.repeat(B)
    t_emb = time_embedding(sd, time)
    if prosody_text is not None and arch.use_prosody_encoder:  # dit.py:225-233
        pt = F.linear(prosody_text, sd["transformer.prosody_text_proj.weight"],
                      sd["transformer.prosody_text_proj.bias"])
        if pt.shape[1] < N:
            pt = F.pad(pt, (0, 0, 0, N - pt.shape[1]))
        else:
            pt = pt[:, :N]
        text_embed = text_embed + pt
    h = input_embedding(sd, x, cond, text_embed, drop_audio_cond)
    ang = rotary_table(sd["transformer.rotary_embed.inv_freq"], N)
    for i in range(arch.depth):
        h = dit_block(sd, arch, i, h, t_emb, mask, ang)
    # AdaLayerNorm_Final (modules.py:322-336): chunk order is (scale, shift)
    emb = F.linear(F.silu(t_emb), sd["transformer.norm_out.linear.weight"], sd["transformer.norm_out.linear.bias"])
    scale, shift = emb.chunk(2, dim=1)
    hn = layer_norm_plain(h) * (1 + scale)[:, None] + shift[:, None]
    out = F.linear(hn, sd["transformer.proj_out.weight"], sd["transformer.proj_out.bias"])
    return (out, h) if return_hidden else out


# --------------------------------------------------------------------------- sampler


def sway_max(steps: int, t_start: float = 0.0, min_ratio: float = 1e-9, safety: float = 0.7) -> float:
    """compute_sway_max (cfm.py:343-373) as called at cfm.py:447."""
    dt = (1.0 - t_start) / max(1, steps)
    p_max = 11.0 if dt >= 0.9 else math.log(min_ratio) / math.log(dt)
    return max(0.0, p_max - 1.0) * safety


def time_grid(steps: int, sway_coef) -> torch.Tensor:
    """t grid of cfm.py:445-453 (fp32): linspace(0,1,steps+1) ** (1 + min(sway_max, coef))."""
    t = torch.linspace(0, 1, int(steps + 1), dtype=torch.float32)
    smax = torch.tensor(sway_max(steps), dtype=torch.float32)
    if sway_coef is not None:
        return t ** (1 + min(smax, sway_coef))
    return t ** (1 + smax)


def lens_to_mask(lens: torch.Tensor, length: int | None = None) -> torch.Tensor:
    """model/utils.py:42-47."""
    length = int(lens.amax()) if length is None else length
    return torch.arange(length)[None, :] < lens[:, None]


@torch.no_grad()
def cfm_sample(sd, arch, cond: torch.Tensor, text: torch.Tensor, duration, *, lens=None, steps=32,
               cfg_strength=1.0, sway_sampling_coef=None, noise: torch.Tensor | None = None, seed=None,
               max_duration=4096, edit_mask=None, use_acc_grl=True, prosody_embeds=None,
               return_trajectory=True):
    """CFM.sample (cfm.py:206-473) for mel `cond` [B, Tc, mel] and int `text` [B, nt] (-1 padded).

    Differences from the reference signature, all test plumbing: `noise` lets the caller inject y0
    (CUDA and CPU generators differ, cfm.py:434); `prosody_embeds` [B,512] stands in for the
    ProsodyEncoder output of cfm.py:248-265 (restated separately in prosody_oracle.py).
    """
    B, Tc, _ = cond.shape
    cond = cond.float()
    if lens is None:
        lens = torch.full((B,), Tc, dtype=torch.long)
    cond_mask = lens_to_mask(lens)
    if edit_mask is not None:
        cond_mask = cond_mask & edit_mask
    if isinstance(duration, int):
        duration = torch.full((B,), duration, dtype=torch.long)
    duration = torch.maximum(torch.maximum((text != -1).sum(dim=-1), lens) + 1, duration)
    duration = duration.clamp(max=max_duration)
    N = int(duration.amax())

    cond_grl = cond  # captured before the prosody projection is added (cfm.py:279 vs :318)
    cond = F.pad(cond, (0, 0, 0, N - Tc))
    prosody_text = None
    if prosody_embeds is not None:
        pm = F.pad(prosody_embeds[:, None, :].expand(-1, Tc, -1), (0, 0, 0, N - Tc))
        cond = cond + F.linear(pm, sd["prosody_to_mel.weight"], sd["prosody_to_mel.bias"])  # bias leaks into pad
        prosody_text = prosody_embeds[:, None, :].expand(-1, text.shape[1], -1)
    cond_mask = F.pad(cond_mask, (0, N - cond_mask.shape[-1]), value=False)[..., None]
    cond_grl = F.pad(cond_grl, (0, 0, 0, N - Tc))
    mask = lens_to_mask(duration) if B > 1 else None

    step_cond = torch.where(cond_mask, cond_grl if use_acc_grl else cond, torch.zeros_like(cond))
    text_c = text_embedding(sd, arch, text, N, drop_text=False)
    text_u = text_embedding(sd, arch, text, N, drop_text=True)

    def fn(t, x):  # cfm.py:382-425
        pred = dit_forward(sd, arch, x, step_cond, text_c, t, mask, False, prosody_text)
        if cfg_strength < 1e-5:
            return pred
        null = dit_forward(sd, arch, x, step_cond, text_u, t, mask, True, prosody_text)
        return (pred + (pred - null) * (cfg_strength * (1 - t) ** 2)).clamp(-20, 20)

    if noise is None:  # cfm.py:430-435
        rows = []
        for d in duration:
            if seed is not None:
                torch.manual_seed(seed)
            rows.append(torch.randn(int(d), arch.mel_dim))
        noise = torch.nn.utils.rnn.pad_sequence(rows, padding_value=0, batch_first=True)
    y = noise.float().clone()
    t = time_grid(steps, sway_sampling_coef)
    traj = [y]
    for i in range(steps):  # torchdiffeq fixed-grid Euler: y += (t1 - t0) * f(t0, y)
        y = y + (t[i + 1] - t[i]) * fn(t[i], y)
        if return_trajectory:
            traj.append(y)
    out = torch.where(cond_mask, cond, y)
    return out, (torch.stack(traj) if return_trajectory else y)


# --------------------------------------------------------------------------- mel front-end


def mel_spectrogram(wav: torch.Tensor, n_fft=1024, hop=256, n_mels=100, sr=24000) -> torch.Tensor:
    """get_vocos_mel_spectrogram (modules.py:75-101): torchaudio MelSpectrogram(power=1, htk, no norm) -> log(clamp 1e-5)."""
    import torchaudio

    tf = torchaudio.transforms.MelSpectrogram(sample_rate=sr, n_fft=n_fft, win_length=n_fft, hop_length=hop,
                                              n_mels=n_mels, power=1, center=True, normalized=False, norm=None)
    return tf(wav).clamp(min=1e-5).log()
