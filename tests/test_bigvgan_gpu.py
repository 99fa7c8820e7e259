"""BigVGAN-v2 generator on hardware: `lemas_bigvgan_decode` (csrc/bigvgan.cu: tap-GEMM convolutions on
tcgen05, 3-tap-GEMM transposed convolutions, fused anti-aliased SnakeBeta) against oracle/bigvgan_oracle.py — the fp32
CPU restatement of the published algorithm (PARITY UNPINNED: the reference's generator is an un-vendored submodule).

Bar: waveform SNR >= 40 dB against the fp32 oracle on the same mel and weights (fp16 tensor-core operands, fp32
accumulation and residual streams; the same bar as the Vocos tests)."""
import json

import pytest
import torch

from lemas_tts import synthetic as syn
from oracle import bigvgan_oracle as bo

pytestmark = pytest.mark.gpu

SNR_DB = 40.0


def _snr(got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    return 10 * torch.log10(ref.pow(2).mean() / (got - ref).pow(2).mean().clamp_min(1e-30)).item()


def _oracle(arch, sd, mel):
    return bo.bigvgan_forward(sd, mel, arch.upsample_rates, arch.upsample_kernel_sizes, arch.resblock_kernel_sizes,
                              arch.resblock_dilation_sizes, arch.snake_logscale, arch.use_tanh_at_final)


@pytest.mark.parametrize("B,T", [(1, 1), (2, 50), (3, 97), (1, 300)])
def test_tiny_generator_matches_oracle(B, T):
    from lemas_tts.bigvgan import BigVGAN

    arch = syn.TINY_BIGVGAN
    sd = syn.make_bigvgan_state_dict(arch, seed=17)
    voc = BigVGAN(arch.to_config())
    voc.load_state_dict(sd, strict=True)
    voc = voc.eval().to("cuda")
    mel = syn.synthetic_ref_mel(B, T, 100, seed=T).permute(0, 2, 1).contiguous()
    wav = voc(mel.cuda())
    ref = _oracle(arch, sd, mel)
    assert wav.shape == ref.shape == (B, 1, T * 8)
    snr = _snr(wav, ref)
    print(f"tiny BigVGAN B={B} T={T}: SNR {snr:.1f} dB")
    assert snr >= SNR_DB
    again = voc(mel.cuda())
    assert torch.equal(wav, again), "decode is not deterministic"


def test_full_generator_matches_oracle():
    """nvidia/bigvgan_v2_24khz_100band_256x architecture (112 M parameters, 1536 -> 24 channels, 256x), seeded weights."""
    from lemas_tts.bigvgan import BigVGAN

    arch = syn.FULL_BIGVGAN
    sd = syn.make_bigvgan_state_dict(arch, seed=17)
    voc = BigVGAN()
    voc.load_state_dict(sd, strict=True)
    voc = voc.eval().to("cuda")
    mel = syn.synthetic_ref_mel(2, 40, 100, seed=5).permute(0, 2, 1).contiguous()
    wav = voc(mel.cuda())
    ref = _oracle(arch, sd, mel)
    assert wav.shape == ref.shape == (2, 1, 40 * 256)
    snr = _snr(wav, ref)
    print(f"full BigVGAN B=2 T=40: SNR {snr:.1f} dB, clamped samples {(ref.abs() >= 1).float().mean().item():.3f}")
    assert snr >= SNR_DB


def test_load_vocoder_bigvgan_branch(tmp_path):
    """utils_infer.py:144-158 / :550-551: load_vocoder('bigvgan', is_local=True, local_path) -> vocoder(mel)."""
    from lemas_tts.infer.utils_infer import load_vocoder

    arch = syn.TINY_BIGVGAN
    sd = syn.make_bigvgan_state_dict(arch, seed=3)
    wn = {}
    for k, v in sd.items():   # the published checkpoint carries weight-normalised convolutions
        if k.endswith(".weight") and v.dim() == 3:
            wn[k[:-7] + ".weight_v"] = v * 2.0
            wn[k[:-7] + ".weight_g"] = v.flatten(1).norm(dim=1).view(-1, 1, 1)
        else:
            wn[k] = v
    (tmp_path / "config.json").write_text(json.dumps(arch.to_config()))
    torch.save({"generator": wn}, tmp_path / "bigvgan_generator.pt")
    voc = load_vocoder("bigvgan", is_local=True, local_path=str(tmp_path), device="cuda")
    mel = syn.synthetic_ref_mel(1, 33, 100, seed=2).permute(0, 2, 1).contiguous()
    wav = voc(mel.cuda())
    assert wav.shape == (1, 1, 33 * 8) and torch.isfinite(wav).all()
    assert _snr(wav, _oracle(arch, sd, mel)) >= SNR_DB
    with pytest.raises(FileNotFoundError):
        load_vocoder("bigvgan", is_local=False, device="cuda")   # hub id, no network: fails like a failed download
