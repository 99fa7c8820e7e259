"""Waveform-side pre / post-processing on the device (csrc/audio.cu) against the reference's host
arithmetic (/root/reference/lemas_tts/infer/utils_infer.py:487-493, 552-553, 581-622).

Bars: the cross-fade + clip is BIT-EXACT against numpy's float64 evaluation (same formula, one rounding per operation);
the RMS is a reduction whose summation order differs from torch's, so RMS / scaled audio are held to 1e-6 relative."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _host_prep(audio, target_rms=0.1):
    if audio.shape[0] > 1:
        audio = torch.mean(audio, dim=0, keepdim=True)
    rms = torch.sqrt(torch.mean(torch.square(audio)))
    if rms < target_rms:
        audio = audio * target_rms / rms
    return audio, rms


@pytest.mark.parametrize("channels,gain,n", [(1, 0.02, 36001), (2, 0.03, 240000), (2, 0.5, 50000), (1, 0.1, 4096)])
def test_prep_reference_audio_matches_host(channels, gain, n):
    from lemas_tts import audio_native

    g = torch.Generator().manual_seed(n)
    audio = torch.randn(channels, n, generator=g) * gain
    ref, rms = _host_prep(audio)
    mono, stats = audio_native.prep_reference_audio(audio.cuda(), 0.1)
    assert mono.shape == (1, n)
    assert abs(stats[0].item() - rms.item()) <= 1e-6 * rms.item() and stats[1].item() == pytest.approx(0.1)
    assert torch.allclose(mono.cpu(), ref, rtol=2e-6, atol=1e-9)
    # un-scaling restores the mono mix (utils_infer.py:552-553 applies it to the generated wave)
    back = audio_native.unscale_(mono.clone(), stats).cpu()
    mix = audio.mean(0, keepdim=True) if channels > 1 else audio
    assert torch.allclose(back, mix, rtol=1e-5, atol=1e-8)


@pytest.mark.parametrize("lens,fade", [([30000, 20000], 0.15), ([9000, 2500, 40000], 0.15), ([5000], 0.15),
                                        ([4000, 4000], 0.0), ([3000, 1000, 1], 0.15)])
def test_cross_fade_concat_is_bit_exact(lens, fade):
    from lemas_tts import audio_native
    from lemas_tts.infer.utils_infer import cross_fade_concat

    g = torch.Generator().manual_seed(sum(lens))
    waves = [torch.randn(n, generator=g) * 0.6 for n in lens]   # some samples beyond the +-0.999 clip
    want = np.clip(cross_fade_concat([w.numpy() for w in waves], fade), -0.999, 0.999)
    got = audio_native.cross_fade_concat([w.cuda() for w in waves], fade, 24000, clip=0.999)
    assert got.dtype == want.dtype and got.shape == want.shape
    assert np.array_equal(got, want)
