"""GPU parity of the tcgen05 GEMM (csrc/gemm.cu) and its fused epilogues, through the C ABI.

Checker: plain PyTorch fp32 (TF32 off) on the SAME fp16-rounded operands, so the tolerance measures the
kernel (fp32 accumulation order + one output rounding), not operand quantisation:
  fp32 outputs: |err| <= 2e-3 * (1 + |ref|)      fp16 outputs: |err| <= 2e-3 * (1 + |ref|) + fp16 ulp
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def _rand(shape, seed, scale=1.0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to("cuda").to(dtype)


def _close(got, ref, tol=2e-3, what=""):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    bound = tol * (1 + ref.abs())
    bad = err > bound
    assert not bad.any(), f"{what}: max err {err.max().item():.3e} at {bad.nonzero()[0].tolist()} ref {ref[bad][0].item():.4f} got {got[bad][0].item():.4f}; {int(bad.sum())} bad of {bad.numel()}"


@pytest.mark.parametrize("block_n", [64, 128, 256])
@pytest.mark.parametrize("rows,k,n,max_ctas", [(300, 256, 512, 0), (1000, 1024, 768, 3), (128, 64, 256, 0)])
def test_bias_f16(block_n, rows, k, n, max_ctas):
    from lemas_tts import ops, _native as nv

    a = _rand((rows, k), 1, dtype=torch.float16)
    w = _rand((n, k), 2, 1 / math.sqrt(k), torch.float16)
    bias = _rand((n,), 3)
    out = torch.zeros(rows, n, device="cuda", dtype=torch.float16)
    ops.gemm(a, w, epilogue=nv.EPI_BIAS_F16, bias=bias, block_n=block_n, out16=out, max_ctas=max_ctas)
    torch.cuda.synchronize()
    _close(out, a.float() @ w.float().T + bias, what=f"bias_f16 bn={block_n}")


def test_bias_f32_partial_n():
    from lemas_tts import ops, _native as nv

    rows, k, n = 333, 512, 100
    a = _rand((rows, k), 4, dtype=torch.float16)
    w = torch.zeros(128, k, device="cuda", dtype=torch.float16)
    w[:n] = _rand((n, k), 5, 1 / math.sqrt(k), torch.float16)
    bias = _rand((n,), 6)
    out = torch.full((rows, 128), 7.0, device="cuda")
    ops.gemm(a, w, epilogue=nv.EPI_BIAS_F32, n=n, bias=bias, block_n=128, out32=out)
    torch.cuda.synchronize()
    _close(out[:, :n], a.float() @ w[:n].float().T + bias, what="bias_f32")
    assert (out[:, n:] == 7.0).all(), "columns >= n must not be written"


@pytest.mark.parametrize("epi_name", ["gelu_tanh", "gelu_erf", "mish"])
def test_activation_epilogues(epi_name):
    from lemas_tts import ops, _native as nv

    rows, k, n = 520, 256, 512
    a = _rand((rows, k), 7, 2.0, torch.float16)
    w = _rand((n, k), 8, 1 / math.sqrt(k), torch.float16)
    bias = _rand((n,), 9)
    out = torch.zeros(rows, n, device="cuda", dtype=torch.float16)
    epi = {"gelu_tanh": nv.EPI_GELU_TANH_F16, "gelu_erf": nv.EPI_GELU_ERF_F16, "mish": nv.EPI_MISH_F16}[epi_name]
    ops.gemm(a, w, epilogue=epi, bias=bias, block_n=256, out16=out)
    torch.cuda.synchronize()
    pre = a.float() @ w.float().T + bias
    ref = {"gelu_tanh": lambda z: F.gelu(z, approximate="tanh"), "gelu_erf": F.gelu, "mish": F.mish}[epi_name](pre)
    _close(out, ref, what=epi_name)


def test_gate_resid_rowmask():
    from lemas_tts import ops, _native as nv

    B, N, k, n = 3, 150, 256, 256
    a = _rand((B * N, k), 10, dtype=torch.float16)
    w = _rand((n, k), 11, 1 / math.sqrt(k), torch.float16)
    bias = _rand((n,), 12)
    gate = _rand((B, n), 13)
    x = _rand((B * N, n), 14)
    valid = torch.tensor([150, 70, 1], device="cuda", dtype=torch.int32)
    ref = a.float() @ w.float().T + bias
    keep = (torch.arange(N, device="cuda")[None] < valid[:, None]).reshape(-1, 1)
    ref = x + gate.repeat_interleave(N, 0) * torch.where(keep, ref, torch.zeros_like(ref))
    out = x.clone()
    ops.gemm(a, w, epilogue=nv.EPI_GATE_RESID_F32, bias=bias, block_n=128, out32=out, resid=out, gate=gate,
             gate_bstride=n, row_valid=valid, seq_len=N)
    torch.cuda.synchronize()
    _close(out, ref, what="gate_resid")


def test_add_f32_f16():
    from lemas_tts import ops, _native as nv

    rows, k, n = 400, 128, 256
    a = _rand((rows, k), 15, dtype=torch.float16)
    w = _rand((n, k), 16, 1 / math.sqrt(k), torch.float16)
    add = _rand((rows, n), 17)
    o32 = torch.zeros(rows, n, device="cuda")
    o16 = torch.zeros(rows, n, device="cuda", dtype=torch.float16)
    ops.gemm(a, w, epilogue=nv.EPI_ADD_F32_F16, block_n=256, out32=o32, out16=o16, resid=add)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().T + add
    _close(o32, ref, what="add32")
    _close(o16, ref, what="add16")


@pytest.mark.parametrize("B,N,heads,rope_heads", [(2, 200, 4, 4), (1, 130, 16, 16), (2, 64, 4, 2)])
def test_qkv_rope(B, N, heads, rope_heads):
    from lemas_tts import ops, _native as nv
    from oracle import lemas_oracle as orc

    D, inner = 256, heads * 64
    a = _rand((B * N, D), 18, dtype=torch.float16)
    w = _rand((3 * inner, D), 19, 1 / math.sqrt(D), torch.float16)
    bias = _rand((3 * inner,), 20)
    inv_freq = 1.0 / (10000.0 ** (torch.arange(0, 64, 2).float() / 64))
    ang = torch.outer(torch.arange(N).float(), inv_freq)
    rope = torch.stack((ang.cos(), ang.sin()), dim=-1).cuda().contiguous()  # [N, 32, 2]
    npad = (N + 63) // 64 * 64
    qk = torch.zeros(B * N, 2 * inner, device="cuda", dtype=torch.float16)
    vt = torch.zeros(B, heads, 64, npad, device="cuda", dtype=torch.float16)
    ops.gemm(a, w, epilogue=nv.EPI_QKV_ROPE, bias=bias, block_n=256, out16=qk, rope=rope, rope_cols=rope_heads * 64,
             inner=inner, vt=vt, seq_len=N)
    torch.cuda.synchronize()
    pre = (a.float() @ w.float().T + bias).cpu()
    q, k, v = pre.split(inner, dim=-1)
    q = q.view(B, N, heads, 64).transpose(1, 2)
    k = k.view(B, N, heads, 64).transpose(1, 2)
    full = orc.rotary_table(inv_freq, N)
    q = torch.cat((orc.apply_rotary(q[:, :rope_heads], full), q[:, rope_heads:]), dim=1)
    k = torch.cat((orc.apply_rotary(k[:, :rope_heads], full), k[:, rope_heads:]), dim=1)
    got_q = qk[:, :inner].view(B, N, heads, 64).transpose(1, 2).cpu()
    got_k = qk[:, inner:].view(B, N, heads, 64).transpose(1, 2).cpu()
    _close(got_q, q, what="q rope")
    _close(got_k, k, what="k rope")
    v_ref = v.view(B, N, heads, 64).permute(0, 2, 3, 1)  # [B, H, 64, N]
    _close(vt[..., :N].cpu(), v_ref, what="v transposed")


def test_grouped_conv31_mish():
    """modules.py:171-176: Conv1d(1024, 1024, 31, groups=16, padding=15) + Mish as a 31-tap GEMM."""
    from lemas_tts import ops, _native as nv

    B, N, D = 2, 300, 1024
    x = _rand((B, N, D), 21, dtype=torch.float16)
    wt = _rand((D, 64, 31), 22, 1 / math.sqrt(31 * 64))
    bias = _rand((D,), 23)
    w_tap = wt.permute(2, 0, 1).contiguous().to(torch.float16).view(31 * D, 64)
    out = torch.zeros(B * N, D, device="cuda", dtype=torch.float16)
    ops.gemm(x, w_tap, epilogue=nv.EPI_MISH_F16, n=D, bias=bias, block_n=64, out16=out, taps=31, tap_pad=15,
             w_tap_stride=D, group_cols=64, k_per_tap=64, seq_len=N)
    torch.cuda.synchronize()
    ref = F.mish(F.conv1d(x.float().transpose(1, 2), wt.to(torch.float16).float(), bias, padding=15, groups=16))
    _close(out.view(B, N, D), ref.transpose(1, 2), what="grouped conv")


def test_dense_conv7_bias_f32():
    """vocos backbone.embed: Conv1d(100, 512, 7, padding=3) as a 7-tap GEMM over a 128-padded channel dim."""
    from lemas_tts import ops, _native as nv

    B, T, Cin, Cout = 2, 211, 100, 512
    x = torch.zeros(B, T, 128, device="cuda", dtype=torch.float16)
    x[..., :Cin] = _rand((B, T, Cin), 24, dtype=torch.float16)
    wt = _rand((Cout, Cin, 7), 25, 1 / math.sqrt(7 * Cin))
    bias = _rand((Cout,), 26)
    w_tap = torch.zeros(7, Cout, 128, device="cuda", dtype=torch.float16)
    w_tap[..., :Cin] = wt.permute(2, 0, 1).to(torch.float16)
    out = torch.zeros(B * T, Cout, device="cuda")
    ops.gemm(x, w_tap.view(7 * Cout, 128), epilogue=nv.EPI_BIAS_F32, n=Cout, bias=bias, block_n=256, out32=out, taps=7,
             tap_pad=3, w_tap_stride=Cout, k_per_tap=128, seq_len=T)
    torch.cuda.synchronize()
    ref = F.conv1d(x[..., :Cin].float().transpose(1, 2), wt.to(torch.float16).float(), bias, padding=3)
    _close(out.view(B, T, Cout), ref.transpose(1, 2), what="dense conv")


@pytest.mark.parametrize("k,dil,C,bn", [(3, 5, 64, 64), (7, 3, 192, 64), (11, 5, 128, 128), (11, 1, 256, 256)])
def test_dilated_conv_gate_resid(k, dil, C, bn):
    """BigVGAN AMPBlock1 convolutions (Conv1d(C, C, k, dilation d, padding d (k-1)/2)) as tap GEMMs with
    lemas_gemm_desc.tap_dilation; residual epilogue with gate = NULL (= 1), two sequences (zero padding at both ends)."""
    from lemas_tts import ops, _native as nv

    B, T = 2, 333
    x = _rand((B, T, C), 40 + k, dtype=torch.float16)
    wt = _rand((C, C, k), 41 + k, 1 / math.sqrt(k * C))
    bias = _rand((C,), 42 + k)
    res = _rand((B * T, C), 43 + k)
    w_tap = wt.permute(2, 0, 1).contiguous().to(torch.float16).view(k * C, C)
    out = torch.zeros(B * T, C, device="cuda")
    ops.gemm(x, w_tap, epilogue=nv.EPI_GATE_RESID_F32, n=C, bias=bias, block_n=bn, out32=out, resid=res, taps=k,
             tap_pad=(k - 1) // 2, tap_dilation=dil, w_tap_stride=C, k_per_tap=C, seq_len=T)
    torch.cuda.synchronize()
    ref = F.conv1d(x.float().transpose(1, 2), wt.to(torch.float16).float(), bias, dilation=dil, padding=dil * (k - 1) // 2)
    _close(out.view(B, T, C), ref.transpose(1, 2) + res.view(B, T, C), what=f"dilated conv k={k} d={dil}")


def test_mish_resid_f32():
    from lemas_tts import ops, _native as nv

    rows, k, n = 260, 128, 128
    a = _rand((rows, k), 27, dtype=torch.float16)
    w = _rand((n, k), 28, 1 / math.sqrt(k), torch.float16)
    bias = _rand((n,), 29)
    res = _rand((rows, n), 30)
    out = torch.zeros(rows, n, device="cuda")
    ops.gemm(a, w, epilogue=nv.EPI_MISH_RESID_F32, bias=bias, block_n=128, out32=out, resid=res)
    torch.cuda.synchronize()
    _close(out, F.mish(a.float() @ w.float().T + bias) + res, what="mish_resid")


def test_invalid_arguments_fail_loudly():
    from lemas_tts import ops, _native as nv

    a = torch.zeros(128, 100, device="cuda", dtype=torch.float16)  # K not a multiple of 64
    w = torch.zeros(128, 100, device="cuda", dtype=torch.float16)
    with pytest.raises(RuntimeError):
        ops.gemm(a, w, epilogue=nv.EPI_BIAS_F16, out16=torch.zeros(128, 128, device="cuda", dtype=torch.float16))


# ------------------------------------------------------------------ CTA-pair kernel (csrc/gemm2.cu, block_n=256, n % 256 == 0)


@pytest.mark.parametrize("rows,k,n,max_ctas", [(4374, 1024, 1024, 0), (700, 2048, 1024, 0), (100, 64, 256, 0),
                                               (300, 512, 512, 2), (1000, 1024, 768, 5), (257, 128, 256, 0)])
def test_pair_gate_resid_and_bias_f32(rows, k, n, max_ctas):
    from lemas_tts import ops, _native as nv

    B = 2 if rows % 2 == 0 else 1
    N = rows // B
    a = _rand((rows, k), 40, dtype=torch.float16)
    w = _rand((n, k), 41, 1 / math.sqrt(k), torch.float16)
    bias, gate, x = _rand((n,), 42), _rand((B, n), 43), _rand((rows, n), 44)
    valid = torch.tensor([N, max(1, N // 3)][:B], device="cuda", dtype=torch.int32)
    pre = a.float() @ w.float().T + bias
    keep = (torch.arange(N, device="cuda")[None] < valid[:, None]).reshape(-1, 1)
    ref = x + gate.repeat_interleave(N, 0) * torch.where(keep, pre, torch.zeros_like(pre))
    out = x.clone()
    ops.gemm(a, w, epilogue=nv.EPI_GATE_RESID_F32, bias=bias, block_n=256, out32=out, resid=out, gate=gate,
             gate_bstride=n, row_valid=valid, seq_len=N, max_ctas=max_ctas)
    torch.cuda.synchronize()
    _close(out, ref, what="pair gate_resid")
    out2 = torch.full((rows, n), 3.0, device="cuda")
    ops.gemm(a, w, epilogue=nv.EPI_BIAS_F32, bias=bias, block_n=256, out32=out2, max_ctas=max_ctas)
    torch.cuda.synchronize()
    _close(out2, pre, what="pair bias_f32")
    # shared gate row (gate_bstride == 0), no row mask: the sampler's FF2 call
    out3 = x.clone()
    ops.gemm(a, w, epilogue=nv.EPI_GATE_RESID_F32, bias=bias, block_n=256, out32=out3, resid=out3, gate=gate[0],
             seq_len=N, max_ctas=max_ctas)
    torch.cuda.synchronize()
    _close(out3, x + gate[0] * pre, what="pair gate_resid shared gate")


def test_pair_matches_single_cta_kernel_bitwise_inputs():
    """Same operands through both kernels (block_n 128 -> gemm.cu, 256 -> gemm2.cu): results agree to fp32 rounding."""
    from lemas_tts import ops, _native as nv

    rows, k, n = 1500, 1024, 2048
    a = _rand((rows, k), 50, dtype=torch.float16)
    w = _rand((n, k), 51, 1 / math.sqrt(k), torch.float16)
    bias = _rand((n,), 52)
    o1 = torch.zeros(rows, n, device="cuda", dtype=torch.float16)
    o2 = torch.zeros_like(o1)
    ops.gemm(a, w, epilogue=nv.EPI_GELU_TANH_F16, bias=bias, block_n=128, out16=o1)
    ops.gemm(a, w, epilogue=nv.EPI_GELU_TANH_F16, bias=bias, block_n=256, out16=o2)
    torch.cuda.synchronize()
    assert (o1.float() - o2.float()).abs().max() <= 2e-3
    _close(o2, F.gelu(a.float() @ w.float().T + bias, approximate="tanh"), what="pair gelu")
