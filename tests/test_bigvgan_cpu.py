"""BigVGAN branch, CPU side: the oracle restatement against independent formulations, the host-side
weight packing against torch's own convolutions, and the mel front-end against the verbatim reference golden.

The generator's source is an un-vendored submodule of the reference (third_party/BigVGAN): oracle/bigvgan_oracle.py
restates the published algorithm — PARITY UNPINNED for the generator; the mel front-end (modules.py:30-72) is pinned."""
import math

import pytest
import torch
import torch.nn.functional as F

import golden_cases as gc
from lemas_tts import synthetic as syn
from oracle import bigvgan_oracle as bo


def test_slaney_filterbank_restatement_matches_torchaudio():
    import torchaudio

    got = bo.slaney_mel_filterbank(24000, 1024, 100, 0.0, None)
    want = torchaudio.functional.melscale_fbanks(513, 0.0, 12000.0, 100, 24000, norm="slaney", mel_scale="slaney").t()
    assert got.shape == (100, 513)
    assert (got - want).abs().max() <= 1e-6 * want.abs().max() * 10
    # unit-area triangles: every filter integrates to ~1 over frequency (Slaney normalisation)
    area = got.sum(1) * (12000.0 / 512)
    assert torch.allclose(area[50:-1], torch.ones_like(area[50:-1]), atol=0.02)   # wide filters: well sampled


def test_bigvgan_mel_frontend_matches_reference_golden():
    from lemas_tts.model.modules import MelSpec

    want = torch.load(gc.GOLDEN / "bigvgan_mel.pt", weights_only=True)["mel"]
    wav = syn.synthetic_ref_audio(2, 24000 + 77, seed=21)
    got = MelSpec(mel_spec_type="bigvgan")(wav)
    assert got.shape == want.shape == (2, 100, (24000 + 77) // 256)
    assert (got - want).abs().max() < 2e-4


def test_aa_filter_is_the_published_kaiser_sinc():
    from lemas_tts.engine import bigvgan_aa_filter

    f = bo.aa_filter()
    assert f.shape == (12,) and abs(float(f.sum()) - 1.0) < 1e-6
    assert torch.allclose(f, f.flip(0), atol=1e-7)            # linear phase
    assert torch.equal(f, bo.kaiser_sinc_filter1d(0.25, 0.3, 12))
    assert torch.allclose(bigvgan_aa_filter(), f, atol=1e-7)  # the copy the product hands to the kernel
    # half-band low-pass: passes DC, rejects the 2x-rate Nyquist
    alt = torch.tensor([(-1.0) ** i for i in range(12)])
    assert abs(float((f * alt).sum())) < 1e-3


def _activation_clamped_index_form(x, ea, ib, f):
    """The formulation csrc/bigvgan.cu:snake_aa_kernel implements (one channel, python loops):
       u[n] = 2 sum_j x[clamp(j)] f[n + 5 - 2j];  s = u + ib sin^2(u ea);  y[t] = sum_k s[clamp(2t + k - 5)] f[k]."""
    T = x.numel()
    u = torch.zeros(2 * T, dtype=torch.float64)
    for n in range(2 * T):
        a = n // 2
        js = range(a - 3, a + 3) if n % 2 == 0 else range(a - 2, a + 4)
        acc = 0.0
        for j in js:
            k = n + 5 - 2 * j
            assert 0 <= k < 12
            acc += float(x[min(max(j, 0), T - 1)]) * float(f[k])
        u[n] = 2.0 * acc
    s = u + ib * torch.sin(u * ea) ** 2
    y = torch.zeros(T, dtype=torch.float64)
    for t in range(T):
        y[t] = sum(float(s[min(max(2 * t + k - 5, 0), 2 * T - 1)]) * float(f[k]) for k in range(12))
    return y


@pytest.mark.parametrize("T", [1, 2, 7, 40])
def test_activation1d_oracle_equals_the_clamped_index_form_the_kernel_uses(T):
    g = torch.Generator().manual_seed(T)
    x = torch.randn(1, 1, T, generator=g, dtype=torch.float64) * 2
    alpha, beta = torch.tensor([0.3], dtype=torch.float64), torch.tensor([-0.2], dtype=torch.float64)
    f = bo.aa_filter().double()
    want = bo.activation1d(x, alpha, beta, f)[0, 0]
    got = _activation_clamped_index_form(x[0, 0], math.exp(0.3), 1.0 / (math.exp(-0.2) + 1e-9), f)
    assert torch.allclose(got, want, atol=1e-12)


@pytest.mark.parametrize("r", [2, 4])
def test_upsample_packing_reproduces_conv_transpose(r):
    """ConvTranspose1d(k = 2r, stride r, pad r/2) as the 3-tap GEMM of csrc/bigvgan.cu, emulated with matmuls."""
    from lemas_tts.engine import bigvgan_pack_upsample

    g = torch.Generator().manual_seed(r)
    cin, cout, T = 24, 12, 19
    cin_pad, cpad = 64, 64
    up = torch.randn(cin, cout, 2 * r, generator=g)
    x = torch.randn(1, cin, T, generator=g)
    want = F.conv_transpose1d(x, up, stride=r, padding=r // 2)[0].t()            # [T*r, cout]
    P = bigvgan_pack_upsample(up, r, cin_pad, cpad).reshape(3, r * cpad, cin_pad)
    rows = torch.zeros(T, cin_pad)
    rows[:, :cin] = x[0].t()
    out = torch.zeros(T, r * cpad)
    for tap, delta in enumerate((-1, 0, 1)):
        shifted = torch.zeros_like(rows)                                         # row m reads input row m + delta
        lo, hi = max(0, -delta), min(T, T - delta)
        shifted[lo:hi] = rows[lo + delta: hi + delta]
        out += shifted @ P[tap].t()
    got = out.reshape(T * r, cpad)
    assert torch.allclose(got[:, :cout], want, atol=1e-5)
    assert torch.count_nonzero(got[:, cout:]) == 0


def test_conv_packing_reproduces_dilated_conv():
    from lemas_tts.engine import bigvgan_pack_conv

    g = torch.Generator().manual_seed(3)
    c, k, d, T = 20, 7, 3, 33
    w = torch.randn(c, c, k, generator=g)
    x = torch.randn(1, c, T, generator=g)
    want = F.conv1d(x, w, dilation=d, padding=d * (k - 1) // 2)[0].t()
    P = bigvgan_pack_conv(w, 64, 64).reshape(k, 64, 64)
    rows = torch.zeros(T, 64)
    rows[:, :c] = x[0].t()
    out = torch.zeros(T, 64)
    for tap in range(k):
        delta = (tap - (k - 1) // 2) * d
        shifted = torch.zeros_like(rows)
        lo, hi = max(0, -delta), min(T, T - delta)
        if hi > lo:
            shifted[lo:hi] = rows[lo + delta: hi + delta]
        out += shifted @ P[tap].t()
    assert torch.allclose(out[:, :c], want, atol=1e-4)


def test_weight_norm_checkpoints_load_like_plain_ones():
    from lemas_tts.bigvgan import BigVGAN

    arch = syn.TINY_BIGVGAN
    sd = syn.make_bigvgan_state_dict(arch)
    plain = BigVGAN(arch.to_config())
    plain.load_state_dict(sd, strict=True)
    wn = {}
    g = torch.Generator().manual_seed(1)
    for k, v in sd.items():
        if k.endswith(".weight") and v.dim() == 3:
            scale = torch.rand(v.shape[0], 1, 1, generator=g) + 0.5
            wn[k[:-7] + ".weight_v"] = v * scale                       # any positive rescaling of v ...
            wn[k[:-7] + ".weight_g"] = v.flatten(1).norm(dim=1).view(-1, 1, 1)   # ... with g = ||w|| gives w back
        else:
            wn[k] = v
    flt = bo.aa_filter()
    wn["resblocks.0.activations.0.upsample.filter"] = flt.view(1, 1, -1)
    wn["resblocks.0.activations.0.downsample.lowpass.filter"] = flt.view(1, 1, -1)
    folded = BigVGAN(arch.to_config())
    folded.load_state_dict(wn, strict=True)
    for (ka, a), (kb, b) in zip(plain.state_dict().items(), folded.state_dict().items()):
        assert ka == kb and torch.allclose(a, b, atol=1e-6), ka
    with pytest.raises(RuntimeError, match="CUDA error"):
        plain(torch.zeros(1, 100, 8))                                   # no CPU path
    bad = dict(wn)
    bad["resblocks.0.activations.0.upsample.filter"] = torch.ones(1, 1, 12) / 12
    with pytest.raises(ValueError, match="Kaiser-sinc"):
        BigVGAN(arch.to_config()).load_state_dict(bad)


def test_oracle_shapes_and_full_architecture_size():
    arch = syn.FULL_BIGVGAN
    sd = syn.make_bigvgan_state_dict(arch)
    assert abs(sum(v.numel() for v in sd.values()) / 1e6 - 112.4) < 0.1       # bigvgan_v2_24khz_100band_256x: 112 M
    tiny = syn.TINY_BIGVGAN
    y = bo.bigvgan_forward(syn.make_bigvgan_state_dict(tiny), torch.randn(2, 100, 9), tiny.upsample_rates,
                           tiny.upsample_kernel_sizes, tiny.resblock_kernel_sizes, tiny.resblock_dilation_sizes)
    assert y.shape == (2, 1, 9 * 8) and float(y.abs().max()) <= 1.0
