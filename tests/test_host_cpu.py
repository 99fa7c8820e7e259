"""CPU checks of the host side that mirrors the reference interface: checkpoint key layout, time grid, text embedding
(torch plumbing) against the oracle / golden vectors, tokenizer and chunking helpers, cross-fade."""
import numpy as np
import pytest
import torch

import golden_cases as gc
from lemas_tts import synthetic as syn


def _cfm(arch):
    from lemas_tts.model.backbones.dit import DiT
    from lemas_tts.model.cfm import CFM

    return CFM(transformer=DiT(**arch.to_kwargs()), mel_spec_kwargs=dict(mel_spec_type="vocos"))


def test_state_dict_key_set_is_the_reference_layout():
    model = _cfm(syn.FULL_ARCH)
    sd = syn.make_dit_state_dict(syn.FULL_ARCH, seed=0)
    assert len(sd) == 368  # 368 tensors for the grl model
    model.load_state_dict(sd, strict=True)
    assert abs(sum(p.numel() for p in model.parameters()) - 336.37e6) < 0.01e6
    with pytest.raises(RuntimeError):
        bad = dict(sd)
        bad.pop("transformer.proj_out.bias")
        model.load_state_dict(bad, strict=True)


def test_time_grid_bit_equal_to_reference():
    from lemas_tts.model.cfm import sway_time_grid

    grids = torch.load(gc.GOLDEN / "time_grids.pt", weights_only=True)
    for key, want in grids.items():
        steps, coef = key.split("_")
        got = sway_time_grid(int(steps), None if coef == "None" else float(coef))
        assert torch.equal(got, want), key


@pytest.mark.parametrize("drop", [False, True])
def test_text_embedding_matches_oracle(drop):
    from oracle import lemas_oracle as orc

    arch = syn.TINY_ARCH
    sd = syn.make_dit_state_dict(arch, seed=11)
    model = _cfm(arch)
    model.load_state_dict(sd, strict=True)
    text = syn.synthetic_text_ids(3, 40, arch.text_num_embeds, seed=2, lengths=[40, 22, 35])
    got = model.transformer.text_embed(text, 131, drop_text=drop)
    want = orc.text_embedding(sd, arch, text, 131, drop_text=drop)
    assert (got - want).abs().max() < 2e-5


def test_load_checkpoint_formats(tmp_path):
    from safetensors.torch import save_file

    from lemas_tts.infer.utils_infer import load_checkpoint

    arch = syn.TINY_ARCH
    sd = syn.make_dit_state_dict(arch, seed=3)
    # safetensors, EMA layout with the bookkeeping and legacy keys the reference strips (utils_infer.py:224-235)
    ema = {"ema_model." + k: v for k, v in sd.items()}
    ema["initted"] = torch.tensor(1.0)
    ema["step"] = torch.tensor(5.0)
    ema["ema_model.mel_spec.mel_stft.spectrogram.window"] = torch.zeros(4)
    ema["ema_model.ctc.proj.0.weight"] = torch.zeros(4)
    p1 = tmp_path / "model.safetensors"
    save_file({k: v.contiguous() for k, v in ema.items()}, str(p1))
    m1 = load_checkpoint(_cfm(arch), str(p1), "cpu", use_ema=True)
    # .pt, non-EMA layout, stored in fp16 like a released CUDA checkpoint
    p2 = tmp_path / "model.pt"
    torch.save({"model_state_dict": {k: v.half() if v.is_floating_point() else v for k, v in sd.items()}}, p2)
    m2 = load_checkpoint(_cfm(arch), str(p2), "cpu", use_ema=False)
    k = "transformer.transformer_blocks.1.attn.to_q.weight"
    assert torch.equal(m1.state_dict()[k], sd[k])
    assert torch.equal(m2.state_dict()[k], sd[k].half().float())


def test_tokenizer_and_text_helpers(tmp_path):
    from lemas_tts.infer.utils_infer import chunk_text, cross_fade_concat
    from lemas_tts.model.utils import get_tokenizer, lens_to_mask, list_str_to_idx

    vocab = tmp_path / "vocab.txt"
    vocab.write_text(" \na\nb\n(en)\n")
    m, n = get_tokenizer(str(vocab), "custom")
    assert n == 4 and m["(en)"] == 3
    ids = list_str_to_idx([["(en)", "a", "zzz"], ["b"]], m)
    assert ids.tolist() == [[3, 1, 0], [2, -1, -1]]
    assert lens_to_mask(torch.tensor([1, 3])).tolist() == [[True, False, False], [True, True, True]]
    assert all(len(c.encode()) <= 40 for c in chunk_text("One two three. Four five six, seven! Eight nine ten?", 40))
    a, b = np.ones(10000), np.zeros(10000)
    x = cross_fade_concat([a, b], 0.15)
    assert len(x) == 20000 - 3600 and x[0] == 1 and x[-1] == 0 and 0.49 < x[10000 - 1800] < 0.51


def test_process_phone_list():
    from lemas_tts.api import process_phone_list

    got = process_phone_list(["(en)", "h", "_", ",", "_", "(zh)", "n", "."])
    assert got == ["(en)h", ",", "(zh)n", "."]


def test_vocos_from_hparams(tmp_path):
    from lemas_tts.vocoder import Vocos

    cfg = tmp_path / "config.yaml"
    cfg.write_text("""
feature_extractor:
  class_path: vocos.feature_extractors.MelSpectrogramFeatures
  init_args: {sample_rate: 24000, n_fft: 1024, hop_length: 256, n_mels: 100, padding: center}
backbone:
  class_path: vocos.models.VocosBackbone
  init_args: {input_channels: 100, dim: 512, intermediate_dim: 1536, num_layers: 8}
head:
  class_path: vocos.heads.ISTFTHead
  init_args: {dim: 512, n_fft: 1024, hop_length: 256, padding: center}
""")
    v = Vocos.from_hparams(str(cfg))
    v.load_state_dict(syn.make_vocos_state_dict(), strict=True)
    assert sum(p.numel() for p in v.parameters()) == 13_531_650 or sum(p.numel() for p in v.parameters()) > 13e6


def test_prosody_encoder_matches_reference(tmp_path):
    """ECAPA-TDNN restatement (torch, host side) vs embeddings minted from the verbatim reference module."""
    import torchaudio

    from lemas_tts.model.backbones.prosody_encoder import ProsodyEncoder, extract_fbank_16k

    case = gc.PROSODY_CASE
    gold = gc.load(case["name"])
    _, audio, _, _, _ = gc.prosody_inputs(case)
    cfg_path, ckpt_path = syn.write_prosody_assets(tmp_path, syn.TINY_PROSODY_CFG, seed=case["pseed"])
    enc = ProsodyEncoder(cfg_path, ckpt_path).eval()
    with torch.no_grad():
        for b in range(case["batch"]):
            a16 = torchaudio.functional.resample(audio[b:b + 1], 24000, 16000).squeeze(0)
            emb = enc(extract_fbank_16k(a16).unsqueeze(0))[0]
            assert (emb - gold["embeds"][b]).abs().max() < 1e-5
            assert abs(float(emb.norm()) - 1.0) < 1e-5
    full = syn.make_prosody_state_dict(syn.PROSODY_CFG)
    assert abs(sum(v.numel() for v in full.values()) - 5.6e6) < 0.1e6  # Pretssel-sized encoder


def test_text_embedding_pair_equals_two_passes():
    arch = syn.TINY_ARCH
    model = _cfm(arch)
    model.load_state_dict(syn.make_dit_state_dict(arch, seed=11), strict=True)
    text = syn.synthetic_text_ids(3, 40, arch.text_num_embeds, seed=2, lengths=[40, 22, 35])
    te = model.transformer.text_embed
    tc, tu = te.forward_pair(text, 131)
    assert torch.equal(tc, te(text, 131, drop_text=False)) and torch.equal(tu, te(text, 131, drop_text=True))


def test_prosody_tables_match_torchaudio():
    """The constant tables of the native prosody path (lemas_tts/prosody_native.py) against torchaudio's own:
    resampling taps (cfm.py:254), povey window and kaldi mel banks (prosody_encoder.py:356-360)."""
    import torchaudio
    from torchaudio.compliance import kaldi
    from torchaudio.functional.functional import _get_sinc_resample_kernel

    from lemas_tts import prosody_native as pn

    for orig, new in [(24000, 16000), (22050, 16000), (16000, 24000)]:
        taps, width, up, down = pn.resample_taps(orig, new)
        import math
        g = math.gcd(orig, new)
        want, w2 = _get_sinc_resample_kernel(orig, new, g, dtype=torch.float32)
        assert width == w2 and (up, down) == (new // g, orig // g)
        assert torch.equal(taps, want[:, 0, :])
    # the polyphase form the kernel evaluates == torchaudio.functional.resample
    taps, width, up, down = pn.resample_taps(24000, 16000)
    x = torch.randn(2, 5001, generator=torch.Generator().manual_seed(0))
    xp = torch.nn.functional.pad(x, (width, width + down))
    n_out = -(-up * x.shape[1] // down)
    y = torch.stack([sum(taps[i % up, k] * xp[:, (i // up) * down + k] for k in range(taps.shape[1]))
                     for i in range(0, n_out, 97)], dim=1)
    assert (y - torchaudio.functional.resample(x, 24000, 16000)[:, ::97]).abs().max() < 1e-5

    window, banks, rng = pn.kaldi_fbank_tables()
    want_banks, _ = kaldi.get_mel_banks(80, 512, 16000.0, 20.0, 0.0, 100.0, -500.0, 1.0)
    assert torch.equal(banks[:, :256], want_banks) and banks[:, 256].abs().max() == 0
    assert torch.equal(window, kaldi._feature_window_function("povey", 400, 0.42, torch.device("cpu"), torch.float32))
    for m in range(80):
        nz = torch.nonzero(banks[m]).flatten()
        assert nz.min() >= rng[m, 0] and nz.max() < rng[m, 1]


def test_mel_filterbank_matches_torchaudio():
    import torchaudio

    from lemas_tts import ops

    fb, rng = ops.mel_filterbank(100, 24000, 1024)
    want = torchaudio.functional.melscale_fbanks(513, 0.0, 12000.0, 100, 24000, norm=None, mel_scale="htk")
    assert torch.equal(fb, want)
    for m in range(100):
        nz = torch.nonzero(want[:, m]).flatten()
        assert nz.min() >= rng[m, 0] and nz.max() < rng[m, 1]


def test_native_ecapa_dataflow_restated_on_cpu(tmp_path):
    """csrc/prosody.cu's data flow (lemas_prosody_encode) restated step by step in torch on the PACKED layouts the
    kernels read — channels-last activations, [k][cin/groups][cout] conv weights, Res2Net groups and the MFA
    concatenation as column slices of one buffer — against the reference's module arithmetic.  Pins the packing and
    the wiring without a GPU; the kernels themselves are checked in tests/test_prosody_gpu.py."""
    from lemas_tts import prosody_native as pn
    from lemas_tts.model.backbones.prosody_encoder import ProsodyEncoder

    cfg_path, ckpt_path = syn.write_prosody_assets(tmp_path, syn.TINY_PROSODY_CFG, seed=3)
    enc = ProsodyEncoder(cfg_path, ckpt_path).eval().encoder

    def conv_cl(x, lay, add=None, act="relu"):          # x: [b, t, cin] channels-last, zero "same" padding
        if add is not None:
            x = x + add
        b, t, _ = x.shape
        k, dil, g = lay["k"], lay["dil"], lay["groups"]
        cin_g, cout_g = lay["cin"] // g, lay["cout"] // g
        y = torch.zeros(b, t, lay["cout"])
        for tap in range(k):
            shift = (tap - (k - 1) // 2) * dil
            xs = torch.zeros_like(x)
            lo, hi = max(0, -shift), min(t, t - shift)
            if hi > lo:
                xs[:, lo:hi] = x[:, lo + shift:hi + shift]
            for gi in range(g):
                y[:, :, gi * cout_g:(gi + 1) * cout_g] += xs[:, :, gi * cin_g:(gi + 1) * cin_g] @ \
                    lay["w"][tap][:, gi * cout_g:(gi + 1) * cout_g]
        y = y + lay["b"]
        return {"relu": torch.relu, "none": lambda v: v}[act](y)

    def tdnn(x, m, add=None, after=None):
        lay = pn.tdnn_layout(m)
        y = torch.nn.functional.layer_norm(conv_cl(x, lay, add), (lay["cout"],), lay["ln_w"], lay["ln_b"], 1e-12)
        return after(y) if after else y

    def dense(x, conv, act):
        w, b = pn.dense_layout(conv)
        return act(x @ w + b)

    g = torch.Generator().manual_seed(5)
    fbank = torch.randn(2, 41, 80, generator=g) * 2 + 3
    with torch.no_grad():
        blocks = list(enc.blocks)
        c = enc.channels[0]
        cg = c // 8
        x = tdnn(fbank, blocks[0])
        cat = torch.zeros(2, 41, c * (len(blocks) - 1))
        for bi, blk in enumerate(blocks[1:]):
            t1 = tdnn(x, blk.tdnn1)
            y = torch.zeros_like(t1)
            y[:, :, :cg] = t1[:, :, :cg]
            for i in range(1, 8):
                add = y[:, :, (i - 1) * cg:i * cg] if i >= 2 else None
                y[:, :, i * cg:(i + 1) * cg] = tdnn(t1[:, :, i * cg:(i + 1) * cg], blk.res2net_block.blocks[i - 1], add)
            t2 = tdnn(y, blk.tdnn2)
            s = dense(dense(t2.mean(1), blk.se_block.conv1, torch.relu), blk.se_block.conv2, torch.sigmoid)
            cat[:, :, bi * c:(bi + 1) * c] = s[:, None, :] * t2 + x
            x = cat[:, :, bi * c:(bi + 1) * c]
        mfa = tdnn(cat, enc.mfa)
        mean = mfa.mean(1)
        std = ((mfa - mean[:, None]) ** 2).mean(1).clamp(1e-12).sqrt()
        concat = torch.cat([mfa, mean[:, None].expand_as(mfa), std[:, None].expand_as(mfa)], dim=2)
        a1 = tdnn(concat, enc.asp.tdnn, after=torch.tanh)
        logits = dense(a1, enc.asp.conv, lambda v: v)
        attn = torch.softmax(logits, dim=1)
        pm = (attn * mfa).sum(1)
        ps = (attn * (mfa - pm[:, None]) ** 2).sum(1).clamp(1e-12).sqrt()
        pooled = torch.nn.functional.layer_norm(torch.cat([pm, ps], 1), (2 * mfa.shape[2],), enc.asp_norm.weight,
                                                enc.asp_norm.bias, 1e-12)
        emb = torch.nn.functional.normalize(dense(pooled, enc.fc, lambda v: v), dim=-1)
        want = enc.forward_torch(fbank)
    assert (emb - want).abs().max() < 2e-5
