"""GPU parity of the HBM-bound kernels (csrc/elementwise.cu) through the C ABI, against PyTorch fp32."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


@pytest.mark.parametrize("rows,dim,B", [(300, 1024, 1), (262, 256, 2), (77, 512, 1)])
def test_ln_modulate(rows, dim, B):
    from lemas_tts import ops

    x = _rand((rows, dim), 1, 3.0) + 0.5
    scale, shift = _rand((B, dim), 2, 0.3), _rand((B, dim), 3, 0.3)
    seq = rows // B
    got = ops.ln_modulate(x, scale, shift, seq_len=seq)
    b_idx = torch.arange(rows, device="cuda") // seq
    ref = F.layer_norm(x, (dim,), eps=1e-6) * (1 + scale[b_idx]) + shift[b_idx]
    assert (got.float() - ref).abs().max() < 4e-3  # fp16 output rounding at |x| <= 8


def test_ln_affine():
    from lemas_tts import ops

    x = _rand((190, 512), 4, 2.0)
    w, b = 1 + _rand((512,), 5, 0.1), _rand((512,), 6, 0.1)
    o16, o32 = ops.ln_affine(x, w, b, want16=True, want32=True)
    ref = F.layer_norm(x, (512,), w, b, eps=1e-6)
    assert (o32 - ref).abs().max() < 2e-5
    assert (o16.float() - ref).abs().max() < 4e-3


@pytest.mark.parametrize("m,k,n", [(32, 1024, 6144 * 2 + 7), (16, 256, 1024), (64, 1024, 2048), (3, 256, 130)])
def test_skinny_linear(m, k, n):
    from lemas_tts import ops

    x, w, b = _rand((m, k), 7), _rand((n, k), 8, 1 / math.sqrt(k)), _rand((n,), 9)
    got = ops.skinny_linear(x, w, b, act_in=True, act_out=False)
    ref = F.linear(F.silu(x).double(), w.double(), b.double()).float()
    assert (got - ref).abs().max() < 2e-5
    got2 = ops.skinny_linear(x, w, b, act_in=False, act_out=True)
    assert (got2 - F.silu(F.linear(x.double(), w.double(), b.double())).float()).abs().max() < 2e-5


def test_time_sinusoid_matches_oracle_formula():
    from lemas_tts import ops

    t = torch.linspace(0, 1, 33)[:-1] ** 4.4856
    got = ops.time_sinusoid(t.cuda())
    w = torch.exp(torch.arange(128).float() * -(math.log(10000) / 127))
    arg = 1000 * t[:, None] * w[None]
    ref = torch.cat((arg.sin(), arg.cos()), -1)
    assert (got.cpu() - ref).abs().max() < 2e-4  # fp32 argument up to 1000: 1 ulp of the argument is 6e-5


def test_cfg_euler():
    from lemas_tts import ops

    rows, mel = 500, 100
    pred = _rand((2, rows, 128), 10, 8.0)
    y = _rand((rows, mel), 11)
    x16 = torch.zeros(2, rows, 128, device="cuda", dtype=torch.float16)
    t, dt, cfg = 0.3, 0.05, 2.0
    pc, pu = pred[0, :, :mel], pred[1, :, :mel]
    f = (pc + (pc - pu) * (cfg * (1 - t) ** 2)).clamp(-20, 20)
    ref = y + dt * f
    traj = torch.zeros(rows, mel, device="cuda")
    ops.cfg_euler(pred, y, x16, t, dt, cfg, copies=2, traj=traj)
    assert (y - ref).abs().max() < 1e-5 and torch.equal(traj, y)
    assert (x16[0, :, :mel].float() - y).abs().max() < 4e-3 and torch.equal(x16[0], x16[1])
    assert (x16[..., mel:] == 0).all()
    y2 = _rand((rows, mel), 11)
    ops.cfg_euler(pred, y2, x16, t, dt, 0.0, copies=1)  # cfg < 1e-5: plain Euler on the cond prediction, no clamp
    assert (y2 - (_rand((rows, mel), 11) + dt * pc)).abs().max() < 1e-5


def test_dwconv7_ln():
    from lemas_tts import ops

    B, T, D = 2, 150, 512
    x = _rand((B, T, D), 12)
    w, b = _rand((D, 1, 7), 13, 0.4), _rand((D,), 14, 0.1)
    lw, lb = 1 + _rand((D,), 15, 0.1), _rand((D,), 16, 0.1)
    got = ops.dwconv7_ln(x, w[:, 0].t().contiguous(), b, lw, lb)
    h = F.conv1d(x.transpose(1, 2), w, b, padding=3, groups=D).transpose(1, 2)
    ref = F.layer_norm(h, (D,), lw, lb, eps=1e-6)
    assert (got.float() - ref).abs().max() < 4e-3


@pytest.mark.parametrize("B,T", [(1, 2), (2, 37), (1, 1250)])
def test_istft_matches_torch(B, T):
    from lemas_tts import ops

    head = torch.zeros(B * T, 1152, device="cuda")
    head[:, :513] = _rand((B * T, 513), 17, 1.5)
    head[:, 0] += 4.0  # some bins hit the clip at 1e2
    head[:, 513:1026] = _rand((B * T, 513), 18, 3.0)
    got = ops.istft_1024(head, B, T)
    mag = torch.clip(torch.exp(head[:, :513]), max=1e2)
    spec = (mag * (torch.cos(head[:, 513:1026]) + 1j * torch.sin(head[:, 513:1026]))).view(B, T, 513).transpose(1, 2)
    ref = torch.istft(spec, 1024, 256, 1024, torch.hann_window(1024).cuda(), center=True)
    assert got.shape == ref.shape
    scale = ref.abs().max().item()
    assert (got - ref).abs().max().item() < 2e-5 * max(scale, 1.0)


@pytest.mark.parametrize("arch_name,B,N,lens", [("TINY_ARCH", 3, 131, [40, 22, 35]), ("FULL_ARCH", 2, 300, [150, 97]),
                                                 ("FULL_ARCH", 1, 70, [90])])
def test_text_embedding_matches_oracle(arch_name, B, N, lens):
    """csrc/text.cu (gather + ConvNeXt-V2 / GRN blocks on the tcgen05 GEMMs) vs the CPU oracle, both CFG variants.
    fp16 GEMM operands: |err| <= 2e-2 on outputs of magnitude O(1-10)."""
    from lemas_tts import synthetic as syn
    from lemas_tts.model.backbones.dit import DiT
    from oracle import lemas_oracle as orc

    arch = getattr(syn, arch_name)
    sd = syn.make_dit_state_dict(arch, seed=11)
    dit = DiT(**arch.to_kwargs())
    dit.load_state_dict({k[len("transformer."):]: v for k, v in sd.items() if k.startswith("transformer.")}, strict=True)
    dit = dit.cuda()
    text = syn.synthetic_text_ids(B, max(lens), arch.text_num_embeds, seed=2, lengths=lens)
    tc, tu = dit.text_embed.forward_pair(text.cuda(), N)
    torch.cuda.synchronize()
    for got, drop in ((tc, False), (tu, True)):
        want = orc.text_embedding(sd, arch, text, N, drop_text=drop)
        err = (got.cpu() - want).abs().max().item()
        scale = want.abs().max().item()
        assert err <= 2e-2 * max(1.0, scale / 4), f"drop={drop}: max err {err:.3e} (scale {scale:.2f})"
        filler = (torch.nn.functional.pad((text + 1)[:, :N], (0, max(0, N - text.shape[1]))) == 0)
        assert (got.cpu()[filler] == 0).all(), "filler rows must be exactly zero"
