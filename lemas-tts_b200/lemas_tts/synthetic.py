"""Deterministic synthetic weights and inputs (no checkpoints or datasets exist offline).

The reference ships no weights in-tree and zero-initialises every AdaLN / output projection
(/root/reference/lemas_tts/model/backbones/dit.py:171-181), which would make the DiT output
identically zero and any parity test vacuous.  This factory therefore draws *every* tensor of
the reference's checkpoint key layout (`transformer.*`, `accent_classifier.*`,
optionally `prosody_*`) from a seeded CPU generator, with gains picked so activations stay O(1)
through 22 blocks and the +-20 clamp of cfm.py:424 does not saturate.

Used by tests, bench.py and the golden-vector generator.  CPU `torch.Generator` streams are
bit-reproducible across machines, so fixtures minted in the build container and tensors made
on the GPU box agree exactly.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict

import torch


@dataclass
class DiTArch:
    """Mirror of `model.arch` in configs/*.yaml plus the two ctor args load_model() adds."""

    dim: int = 1024
    depth: int = 22
    heads: int = 16
    dim_head: int = 64
    ff_mult: int = 2
    text_dim: int = 512
    conv_layers: int = 4
    mel_dim: int = 100
    text_num_embeds: int = 898
    text_mask_padding: bool = True
    qk_norm: str | None = None
    pe_attn_head: int | None = None
    use_prosody_encoder: bool = False

    def to_kwargs(self) -> dict:
        return asdict(self)


FULL_ARCH = DiTArch()
# Reduced architecture for fast CPU oracle runs; same kernels (dim_head=64, dims multiple of 64).
TINY_ARCH = DiTArch(dim=256, depth=2, heads=4, ff_mult=2, text_dim=128, conv_layers=2, text_num_embeds=60)


def _gen(seed: int) -> torch.Generator:
    return torch.Generator().manual_seed(int(seed))


def _normal(g, shape, std):
    return torch.randn(*shape, generator=g, dtype=torch.float32) * std


def _linear(sd, g, name, n_out, n_in, gain=1.0, bias_std=0.02):
    sd[f"{name}.weight"] = _normal(g, (n_out, n_in), gain / math.sqrt(n_in))
    sd[f"{name}.bias"] = _normal(g, (n_out,), bias_std)


def make_dit_state_dict(arch: DiTArch = FULL_ARCH, seed: int = 0) -> dict[str, torch.Tensor]:
    """State dict in the reference key layout (what `load_checkpoint` strict-loads into CFM)."""
    g = _gen(seed)
    sd: dict[str, torch.Tensor] = {}
    D, Dt, M = arch.dim, arch.text_dim, arch.mel_dim
    inner = arch.heads * arch.dim_head
    F = D * arch.ff_mult
    p = "transformer."

    _linear(sd, g, p + "time_embed.time_mlp.0", D, 256)
    _linear(sd, g, p + "time_embed.time_mlp.2", D, D)

    sd[p + "text_embed.text_embed.weight"] = _normal(g, (arch.text_num_embeds + 1, Dt), 1.0)
    for i in range(arch.conv_layers):
        q = f"{p}text_embed.text_blocks.{i}."
        sd[q + "dwconv.weight"] = _normal(g, (Dt, 1, 7), 1.0 / math.sqrt(7))
        sd[q + "dwconv.bias"] = _normal(g, (Dt,), 0.02)
        sd[q + "norm.weight"] = 1.0 + _normal(g, (Dt,), 0.1)
        sd[q + "norm.bias"] = _normal(g, (Dt,), 0.05)
        _linear(sd, g, q + "pwconv1", 2 * Dt, Dt)
        sd[q + "grn.gamma"] = _normal(g, (1, 1, 2 * Dt), 0.2)
        sd[q + "grn.beta"] = _normal(g, (1, 1, 2 * Dt), 0.05)
        _linear(sd, g, q + "pwconv2", Dt, 2 * Dt, gain=0.5)

    if arch.use_prosody_encoder:
        _linear(sd, g, p + "prosody_text_proj", Dt, 512)

    _linear(sd, g, p + "input_embed.proj", D, 2 * M + Dt, gain=0.7)
    for j in (0, 2):
        q = f"{p}input_embed.conv_pos_embed.conv1d.{j}."
        sd[q + "weight"] = _normal(g, (D, D // 16, 31), 1.0 / math.sqrt(31 * D // 16))
        sd[q + "bias"] = _normal(g, (D,), 0.02)

    sd[p + "rotary_embed.inv_freq"] = 1.0 / (
        10000.0 ** (torch.arange(0, arch.dim_head, 2).float() / arch.dim_head)
    )

    for i in range(arch.depth):
        q = f"{p}transformer_blocks.{i}."
        # AdaLN-zero linear: small but non-zero so gates/scales/shifts are exercised.
        _linear(sd, g, q + "attn_norm.linear", 6 * D, D, gain=0.35, bias_std=0.05)
        _linear(sd, g, q + "attn.to_q", inner, D)
        _linear(sd, g, q + "attn.to_k", inner, D)
        _linear(sd, g, q + "attn.to_v", inner, D)
        _linear(sd, g, q + "attn.to_out.0", D, inner)
        _linear(sd, g, q + "ff.ff.0.0", F, D)
        _linear(sd, g, q + "ff.ff.2", D, F)
        if arch.qk_norm == "rms_norm":
            sd[q + "attn.q_norm.weight"] = 1.0 + _normal(g, (arch.dim_head,), 0.1)
            sd[q + "attn.k_norm.weight"] = 1.0 + _normal(g, (arch.dim_head,), 0.1)

    _linear(sd, g, p + "norm_out.linear", 2 * D, D, gain=0.35, bias_std=0.05)
    _linear(sd, g, p + "proj_out", M, D, gain=1.0)

    # training-only head; present in every released checkpoint, so kept in the key set
    _linear(sd, g, "accent_classifier.net.0", D, M)
    _linear(sd, g, "accent_classifier.net.3", 12, D)
    if arch.use_prosody_encoder:
        _linear(sd, g, "prosody_to_mel", M, 512, gain=0.3)
    return sd


@dataclass
class VocosArch:
    """charactr/vocos-mel-24khz hyper-parameters (config.yaml of the pip package `vocos`)."""

    input_channels: int = 100
    dim: int = 512
    intermediate_dim: int = 1536
    num_layers: int = 8
    n_fft: int = 1024
    hop_length: int = 256


FULL_VOCOS = VocosArch()
TINY_VOCOS = VocosArch(dim=128, intermediate_dim=256, num_layers=2)


def make_vocos_state_dict(arch: VocosArch = FULL_VOCOS, seed: int = 7) -> dict[str, torch.Tensor]:
    """State dict in the `pytorch_model.bin` key layout of vocos-mel-24khz (backbone.* / head.*)."""
    g = _gen(seed)
    sd: dict[str, torch.Tensor] = {}
    C, D, H = arch.input_channels, arch.dim, arch.intermediate_dim
    sd["backbone.embed.weight"] = _normal(g, (D, C, 7), 1.0 / math.sqrt(7 * C))
    sd["backbone.embed.bias"] = _normal(g, (D,), 0.02)
    sd["backbone.norm.weight"] = 1.0 + _normal(g, (D,), 0.1)
    sd["backbone.norm.bias"] = _normal(g, (D,), 0.05)
    for i in range(arch.num_layers):
        q = f"backbone.convnext.{i}."
        sd[q + "dwconv.weight"] = _normal(g, (D, 1, 7), 1.0 / math.sqrt(7))
        sd[q + "dwconv.bias"] = _normal(g, (D,), 0.02)
        sd[q + "norm.weight"] = 1.0 + _normal(g, (D,), 0.1)
        sd[q + "norm.bias"] = _normal(g, (D,), 0.05)
        _linear(sd, g, q + "pwconv1", H, D)
        _linear(sd, g, q + "pwconv2", D, H)
        sd[q + "gamma"] = 0.25 + _normal(g, (D,), 0.05)
    sd["backbone.final_layer_norm.weight"] = 1.0 + _normal(g, (D,), 0.1)
    sd["backbone.final_layer_norm.bias"] = _normal(g, (D,), 0.05)
    # log-magnitude half kept small so exp() stays well below the clip at 1e2 for most bins
    _linear(sd, g, "head.out", arch.n_fft + 2, D, gain=0.5)
    sd["head.istft.window"] = torch.hann_window(arch.n_fft)
    return sd


@dataclass
class BigVGANArch:
    """config.json of nvidia/bigvgan_v2_24khz_100band_256x (the model utils_infer.py:150-155 names)."""

    num_mels: int = 100
    upsample_rates: tuple = (4, 4, 2, 2, 2, 2)
    upsample_kernel_sizes: tuple = (8, 8, 4, 4, 4, 4)
    upsample_initial_channel: int = 1536
    resblock_kernel_sizes: tuple = (3, 7, 11)
    resblock_dilation_sizes: tuple = ((1, 3, 5), (1, 3, 5), (1, 3, 5))
    snake_logscale: bool = True
    use_bias_at_final: bool = False
    use_tanh_at_final: bool = False

    def to_config(self) -> dict:
        return dict(num_mels=self.num_mels, upsample_rates=list(self.upsample_rates),
                    upsample_kernel_sizes=list(self.upsample_kernel_sizes),
                    upsample_initial_channel=self.upsample_initial_channel, resblock="1",
                    resblock_kernel_sizes=list(self.resblock_kernel_sizes),
                    resblock_dilation_sizes=[list(d) for d in self.resblock_dilation_sizes], activation="snakebeta",
                    snake_logscale=self.snake_logscale, use_bias_at_final=self.use_bias_at_final,
                    use_tanh_at_final=self.use_tanh_at_final, sampling_rate=24000, n_fft=1024, hop_size=256,
                    win_size=1024, fmin=0, fmax=None)


FULL_BIGVGAN = BigVGANArch()
# 128 -> 64 -> 32 channels (the last stage exercises the channel padding to 64), total up-sampling 8x
TINY_BIGVGAN = BigVGANArch(upsample_rates=(4, 2), upsample_kernel_sizes=(8, 4), upsample_initial_channel=128)


def make_bigvgan_state_dict(arch: BigVGANArch = FULL_BIGVGAN, seed: int = 17) -> dict[str, torch.Tensor]:
    """State dict in the key layout of `bigvgan_generator.pt["generator"]` after remove_weight_norm()."""
    g = _gen(seed)
    sd: dict[str, torch.Tensor] = {}
    ch = arch.upsample_initial_channel
    sd["conv_pre.weight"] = _normal(g, (ch, arch.num_mels, 7), 0.25 / math.sqrt(7 * arch.num_mels))
    sd["conv_pre.bias"] = _normal(g, (ch,), 0.02)
    n = 0
    for i, (r, k) in enumerate(zip(arch.upsample_rates, arch.upsample_kernel_sizes)):
        sd[f"ups.{i}.0.weight"] = _normal(g, (ch, ch // 2, k), 1.0 / math.sqrt(2 * ch))
        sd[f"ups.{i}.0.bias"] = _normal(g, (ch // 2,), 0.02)
        ch //= 2
        for kk in arch.resblock_kernel_sizes:
            q = f"resblocks.{n}."
            for d in range(3):
                sd[f"{q}convs1.{d}.weight"] = _normal(g, (ch, ch, kk), 1.0 / math.sqrt(ch * kk))
                sd[f"{q}convs1.{d}.bias"] = _normal(g, (ch,), 0.02)
                sd[f"{q}convs2.{d}.weight"] = _normal(g, (ch, ch, kk), 0.4 / math.sqrt(ch * kk))
                sd[f"{q}convs2.{d}.bias"] = _normal(g, (ch,), 0.02)
            for a in range(6):
                sd[f"{q}activations.{a}.act.alpha"] = _normal(g, (ch,), 0.3)
                sd[f"{q}activations.{a}.act.beta"] = _normal(g, (ch,), 0.3)
            n += 1
    sd["activation_post.act.alpha"] = _normal(g, (ch,), 0.3)
    sd["activation_post.act.beta"] = _normal(g, (ch,), 0.3)
    sd["conv_post.weight"] = _normal(g, (1, ch, 7), 0.25 / math.sqrt(7 * ch))
    if arch.use_bias_at_final:
        sd["conv_post.bias"] = _normal(g, (1,), 0.02)
    return sd


# ----------------------------------------------------------------------------- inputs


def synthetic_ref_mel(batch: int, frames: int, n_mels: int = 100, seed: int = 0) -> torch.Tensor:
    """log-mel-like reference: clamp(N(-4, 2^2), min=ln 1e-5)  (SURVEY.md §8d)."""
    g = _gen(1000 + seed)
    mel = torch.randn(batch, frames, n_mels, generator=g) * 2.0 - 4.0
    return mel.clamp_(min=math.log(1e-5))


def synthetic_ref_audio(batch: int, samples: int, seed: int = 0) -> torch.Tensor:
    g = _gen(2000 + seed)
    return 0.1 * torch.randn(batch, samples, generator=g)


def synthetic_text_ids(batch: int, n_tokens: int, vocab: int = 898, seed: int = 0,
                       lengths: list[int] | None = None) -> torch.Tensor:
    """Token ids uniform in [0, vocab); rows shorter than n_tokens are padded with -1."""
    g = _gen(3000 + seed)
    ids = torch.randint(0, vocab, (batch, n_tokens), generator=g, dtype=torch.long)
    if lengths is not None:
        for b, n in enumerate(lengths):
            ids[b, n:] = -1
    return ids


def synthetic_noise(durations: list[int], n_mels: int = 100, seed: int = 0) -> torch.Tensor:
    """y0 as CFM.sample builds it (cfm.py:430-435): per-row randn(dur, n_mels), zero padded."""
    g = _gen(4000 + seed)
    n = max(durations)
    y0 = torch.zeros(len(durations), n, n_mels)
    for b, d in enumerate(durations):
        y0[b, :d] = torch.randn(d, n_mels, generator=g)
    return y0


@dataclass
class BenchConfig:
    name: str
    batch: int
    ref_frames: int
    total_frames: int  # N
    n_text: int
    steps: int
    cfg_strength: float
    sway_coef: float | None
    seed: int


# BASELINE.json configs made concrete.  C3's ragged shapes are drawn in bench/tests.
CONFIGS = {
    "C1": BenchConfig("C1", 1, 376, 940, 150, 16, 2.0, 5.0, 0),
    "C2": BenchConfig("C2", 1, 937, 2187, 350, 32, 2.0, 5.0, 0),
    "C3": BenchConfig("C3", 32, 938, 0, 450, 32, 2.0, 3.0, 1),   # ragged: see c3_batch(); total_frames = max duration
    "C4": BenchConfig("C4", 256, 256, 768, 120, 32, 2.0, 3.0, 2),
    "C5": BenchConfig("C5", 1, 2813, 2814, 450, 64, 5.0, 3.0, 3),
}


def c3_batch(batch: int = 32, seed: int = 1, samples: int = 240_000, vocab: int = 898):
    """Config C3 (BASELINE.json configs[2]): mixed-length utterances with raw 10 s reference audio,
    lens ~ U{281..937} frames, ref tokens ~ U{45..150}, gen tokens ~ U{100..300},
    duration_b = lens_b + int(lens_b / ref_b * gen_b) (utils_infer.py:520-527), clamped to 4096 (cfm.py:304)."""
    g = _gen(5000 + seed)
    lens = torch.randint(281, 938, (batch,), generator=g)
    ref_t = torch.randint(45, 151, (batch,), generator=g)
    gen_t = torch.randint(100, 301, (batch,), generator=g)
    dur = torch.tensor([min(4096, int(l) + int(int(l) / int(r) * int(n))) for l, r, n in zip(lens, ref_t, gen_t)])
    n_tok = (ref_t + gen_t).tolist()
    text = synthetic_text_ids(batch, max(n_tok), vocab, seed=seed, lengths=n_tok)
    audio = synthetic_ref_audio(batch, samples, seed=seed)
    return dict(audio=audio, text=text, lens=lens, duration=dur)


# ----------------------------------------------------------------------------- prosody encoder (config C3)

# `pretssel_cfg.json` keys the reference reads (prosody_encoder.py:387-400); values = Pretssel defaults.
PROSODY_CFG = dict(prosody_channels=[512, 512, 512, 512, 1536], prosody_kernel_sizes=[5, 3, 3, 3, 1],
                   prosody_dilations=[1, 2, 3, 4, 1], prosody_attention_channels=128, prosody_res2net_scale=8,
                   prosody_se_channels=128, prosody_global_context=True, prosody_groups=[1, 1, 1, 1, 3],
                   prosody_embed_dim=512, input_feat_per_channel=80)
# reduced width for fast CPU tests (embed dim stays 512: prosody_to_mel / prosody_text_proj are Linear(512, .))
TINY_PROSODY_CFG = dict(PROSODY_CFG, prosody_channels=[64, 64, 64, 64, 192], prosody_attention_channels=32,
                        prosody_se_channels=32)


def prosody_param_shapes(cfg: dict) -> dict[str, tuple]:
    """Parameter names and shapes of the reference ECAPA_TDNN (prosody_encoder.py:30-132) for a pretssel cfg."""
    ch, ks, gr = cfg["prosody_channels"], cfg["prosody_kernel_sizes"], cfg["prosody_groups"]
    scale, se, att = cfg["prosody_res2net_scale"], cfg["prosody_se_channels"], cfg["prosody_attention_channels"]
    shapes: dict[str, tuple] = {}

    def tdnn(prefix, cin, cout, k, g=1):
        shapes[prefix + "conv.weight"] = (cout, cin // g, k)
        shapes[prefix + "conv.bias"] = (cout,)
        shapes[prefix + "norm.weight"] = (cout,)
        shapes[prefix + "norm.bias"] = (cout,)

    def conv(prefix, cin, cout):
        shapes[prefix + "weight"] = (cout, cin, 1)
        shapes[prefix + "bias"] = (cout,)

    tdnn("blocks.0.", cfg["input_feat_per_channel"], ch[0], ks[0], gr[0])
    for i in range(1, len(ch) - 1):
        q = f"blocks.{i}."
        tdnn(q + "tdnn1.", ch[i - 1], ch[i], 1, gr[i])
        for j in range(scale - 1):
            tdnn(q + f"res2net_block.blocks.{j}.", ch[i] // scale, ch[i] // scale, ks[i])
        tdnn(q + "tdnn2.", ch[i], ch[i], 1, gr[i])
        conv(q + "se_block.conv1.", ch[i], se)
        conv(q + "se_block.conv2.", se, ch[i])
        if ch[i - 1] != ch[i]:
            conv(q + "shortcut.", ch[i - 1], ch[i])
    tdnn("mfa.", ch[-1], ch[-1], ks[-1], gr[-1])
    tdnn("asp.tdnn.", ch[-1] * (3 if cfg["prosody_global_context"] else 1), att, 1)
    conv("asp.conv.", att, ch[-1])
    shapes["asp_norm.weight"] = (2 * ch[-1],)
    shapes["asp_norm.bias"] = (2 * ch[-1],)
    conv("fc.", 2 * ch[-1], cfg["prosody_embed_dim"])
    return shapes


def make_prosody_state_dict(cfg: dict = PROSODY_CFG, seed: int = 13) -> dict[str, torch.Tensor]:
    """ECAPA-TDNN weights under the reference's module names.  Every tensor is drawn from its own generator seeded by
    (seed, key), so the values do not depend on enumeration order."""
    import zlib

    sd = {}
    for key, shape in prosody_param_shapes(cfg).items():
        g = _gen(seed * 1_000_003 + zlib.crc32(key.encode()))
        if key.endswith("norm.weight"):
            sd[key] = 1.0 + _normal(g, shape, 0.1)
        elif key.endswith("bias"):
            sd[key] = _normal(g, shape, 0.05)
        else:  # conv weights [out, in/groups, k]
            sd[key] = _normal(g, shape, 1.0 / math.sqrt(shape[1] * shape[2]))
    return sd


def write_prosody_assets(directory, cfg: dict = PROSODY_CFG, seed: int = 13):
    """pretssel_cfg.json + prosody_encoder checkpoint (keys prefixed `prosody_encoder.` like the released file)."""
    import json
    from pathlib import Path

    d = Path(directory)
    d.mkdir(parents=True, exist_ok=True)
    cfg_path, ckpt_path = d / "pretssel_cfg.json", d / "prosody_encoder_UnitY2.pt"
    cfg_path.write_text(json.dumps({"model": cfg}))
    torch.save({"prosody_encoder." + k: v for k, v in make_prosody_state_dict(cfg, seed).items()}, ckpt_path)
    return cfg_path, ckpt_path
