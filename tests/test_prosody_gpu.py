"""GPU parity of the native prosody path (csrc/prosody.cu) through the C ABI.

Checkers (CPU fp32): torchaudio.functional.resample and torchaudio.compliance.kaldi.fbank — the very functions the
reference calls (cfm.py:254, prosody_encoder.py:356) — and the ECAPA-TDNN in the reference's torch arithmetic
(`ECAPA_TDNN.forward_torch`, pinned to the verbatim reference through tests/golden/sample_tiny_prosody.pt).
Tolerances: resampler 1e-5 abs; log-fbank 5e-4 abs (fp32 FFT of |x|^2 ~ 1e2..1e5); L2-normalised embedding 2e-5 abs.
"""
import tempfile
from pathlib import Path

import pytest
import torch
import torchaudio

from lemas_tts import synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("orig,new,n", [(24000, 16000, 240000), (24000, 16000, 5001), (22050, 16000, 7000),
                                        (16000, 24000, 3000)])
def test_resample_matches_torchaudio(orig, new, n):
    from lemas_tts import prosody_native as pn

    wav = syn.synthetic_ref_audio(2, n, seed=n % 89)
    want = torchaudio.functional.resample(wav, orig, new)
    got = pn.resample(wav.cuda(), orig, new).cpu()
    assert got.shape == want.shape
    assert (got - want).abs().max().item() < 1e-5


@pytest.mark.parametrize("n", [160000, 16001, 400, 30123])
def test_kaldi_fbank_matches_torchaudio(n):
    from lemas_tts import prosody_native as pn

    wav = syn.synthetic_ref_audio(2, n, seed=n % 83)
    want = torch.stack([torchaudio.compliance.kaldi.fbank(w[None], num_mel_bins=80, sample_frequency=16000) for w in wav])
    got = pn.kaldi_fbank_80(wav.cuda()).cpu()
    assert got.shape == want.shape == (2, 1 + (n - 400) // 160, 80)
    assert (got - want).abs().max().item() < 5e-4


def test_kaldi_fbank_tiles_short_clips():
    from lemas_tts.model.backbones.prosody_encoder import extract_fbank_16k

    wav = syn.synthetic_ref_audio(1, 150, seed=5)[0]
    want = extract_fbank_16k(wav)
    got = extract_fbank_16k(wav.cuda()).cpu()
    assert got.shape == want.shape and (got - want).abs().max().item() < 5e-4


@pytest.mark.parametrize("cfg_name,batch,frames", [("tiny", 3, 57), ("tiny", 1, 200), ("full", 2, 333), ("full", 1, 998)])
def test_ecapa_matches_torch_reference(cfg_name, batch, frames):
    from lemas_tts.model.backbones.prosody_encoder import ProsodyEncoder

    cfg = syn.TINY_PROSODY_CFG if cfg_name == "tiny" else syn.PROSODY_CFG
    with tempfile.TemporaryDirectory() as tmp:
        cfg_path, ckpt_path = syn.write_prosody_assets(Path(tmp), cfg, seed=21)
        enc = ProsodyEncoder(cfg_path, ckpt_path).eval()
    g = torch.Generator().manual_seed(frames)
    fbank = torch.randn(batch, frames, 80, generator=g) * 2.0 + 3.0     # log-energies of that order
    with torch.no_grad():
        want = enc.encoder.forward_torch(fbank)
        got = enc.cuda()(fbank.cuda()).cpu()
    assert got.shape == want.shape == (batch, 512)
    assert torch.allclose(got.norm(dim=-1), torch.ones(batch), atol=1e-5)
    assert (got - want).abs().max().item() < 2e-5


def test_prosody_embeds_end_to_end_matches_cpu_chain():
    """CFM._prosody_embeds on the GPU (native resampler -> fbank -> ECAPA) against the reference chain on the CPU."""
    from lemas_tts.model.backbones.prosody_encoder import ProsodyEncoder, extract_fbank_16k

    with tempfile.TemporaryDirectory() as tmp:
        cfg_path, ckpt_path = syn.write_prosody_assets(Path(tmp), syn.TINY_PROSODY_CFG, seed=4)
        enc = ProsodyEncoder(cfg_path, ckpt_path).eval()
    raw = syn.synthetic_ref_audio(2, 48000, seed=17)
    with torch.no_grad():
        want = torch.stack([enc.encoder.forward_torch(
            extract_fbank_16k(torchaudio.functional.resample(r[None], 24000, 16000)[0])[None])[0] for r in raw])
        from lemas_tts import prosody_native as pn
        got = enc.cuda()(pn.kaldi_fbank_80(pn.resample(raw.cuda(), 24000, 16000))).cpu()
    assert (got - want).abs().max().item() < 5e-5


def test_native_embeddings_match_verbatim_reference_golden(tmp_path):
    """The whole native chain (resampler -> kaldi fbank -> ECAPA-TDNN) against embeddings minted by the VERBATIM
    reference modules (oracle/gen_golden.py, tests/golden/sample_tiny_prosody.pt)."""
    import golden_cases as gc

    from lemas_tts import prosody_native as pn
    from lemas_tts.model.backbones.prosody_encoder import ProsodyEncoder

    case = gc.PROSODY_CASE
    gold = gc.load(case["name"])
    _, audio, _, _, _ = gc.prosody_inputs(case)
    cfg_path, ckpt_path = syn.write_prosody_assets(tmp_path, syn.TINY_PROSODY_CFG, seed=case["pseed"])
    enc = ProsodyEncoder(cfg_path, ckpt_path).eval().cuda()
    with torch.no_grad():
        got = enc(pn.kaldi_fbank_80(pn.resample(audio.cuda().float().contiguous(), 24000, 16000))).cpu()
    assert got.shape == gold["embeds"].shape
    assert (got - gold["embeds"]).abs().max().item() < 2e-5
