"""GPU parity of Vocos.decode (liblemas_b200.so) against the CPU fp32 restatement (oracle/vocos_oracle.py).

Bar (north_star: "perceptually identical waveforms"): SNR >= 40 dB against the fp32 oracle on the same mel and
log-mel L1 of the re-analysed audio <= 0.05; fp16 GEMM operands, fp32 accumulation/residual/iSTFT.
"""
import math

import pytest
import torch

from lemas_tts import synthetic as syn

pytestmark = pytest.mark.gpu


def _snr_db(got, ref):
    noise = (got - ref).pow(2).mean().item()
    sig = ref.pow(2).mean().item()
    return 10 * math.log10(sig / max(noise, 1e-30))


@pytest.mark.parametrize("arch_name,B,T", [("TINY_VOCOS", 2, 50), ("FULL_VOCOS", 1, 300), ("FULL_VOCOS", 3, 97),
                                           ("FULL_VOCOS", 1, 2)])
def test_decode_matches_oracle(arch_name, B, T):
    from lemas_tts.vocoder import Vocos
    from oracle import vocos_oracle as vo
    from oracle.lemas_oracle import mel_spectrogram

    arch = getattr(syn, arch_name)
    sd = syn.make_vocos_state_dict(arch, seed=7)
    mel = syn.synthetic_ref_mel(B, T, arch.input_channels, seed=40 + T).transpose(1, 2).contiguous()  # [B, 100, T]
    ref = vo.vocos_decode(sd, mel, num_layers=arch.num_layers)
    voc = Vocos(input_channels=arch.input_channels, dim=arch.dim, intermediate_dim=arch.intermediate_dim,
                num_layers=arch.num_layers)
    voc.load_state_dict(sd, strict=True)
    voc = voc.cuda().eval()
    got = voc.decode(mel.cuda())
    torch.cuda.synchronize()
    got = got.cpu()
    assert got.shape == ref.shape == (B, (T - 1) * 256)
    snr = _snr_db(got, ref)
    print(f"vocos {arch_name} B={B} T={T}: SNR {snr:.1f} dB, max abs {(got - ref).abs().max():.3e} (peak {ref.abs().max():.3f})")
    assert snr >= 40.0, f"SNR {snr:.1f} dB"
    if T >= 8:
        l1 = (mel_spectrogram(got) - mel_spectrogram(ref)).abs().mean().item()
        assert l1 <= 0.05, f"log-mel L1 of re-analysed audio {l1:.4f}"


def test_decode_checkpoint_with_feature_extractor_keys_loads():
    from lemas_tts.vocoder import Vocos

    sd = syn.make_vocos_state_dict(syn.FULL_VOCOS, seed=7)
    sd["feature_extractor.mel_spec.spectrogram.window"] = torch.hann_window(1024)
    sd["feature_extractor.mel_spec.mel_scale.fb"] = torch.zeros(513, 100)
    voc = Vocos()
    voc.load_state_dict(sd)
    wav = voc.cuda().decode(torch.zeros(1, 100, 5, device="cuda"))
    assert wav.shape == (1, 1024) and torch.isfinite(wav).all()
