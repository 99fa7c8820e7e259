"""GPU parity of the tcgen05 attention kernels through the C ABI: csrc/attention.cu (v3: long uniform sequences) and
csrc/attention9.cu (v9: batches of short or ragged sequences); `lemas_attention_f16` chooses by shape, the tests force
each kernel on every shape.

Checker: fp32 softmax(QK^T/8 + keymask)V in PyTorch on the same fp16 q/k/v.  P is rounded to fp16 before the PV
MMA, so the bar is 4e-3 absolute on outputs of magnitude O(1) (documented in DESIGN.md).
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(q, k, v, kv_len):
    B, H, N, _ = q.shape
    s = (q.float() @ k.float().transpose(-1, -2)) / 8.0
    if kv_len is not None:
        mask = torch.arange(N, device=q.device)[None] < kv_len[:, None]
        s = s.masked_fill(~mask[:, None, None, :], float("-inf"))
    return s.softmax(-1) @ v.float()


@pytest.fixture(params=[0, 40, 100], ids=["v3", "v9", "auto"])
def kernel_variant(request):
    """Force one production kernel (0 = v3, 40 = v9) or leave the choice to the library (100)."""
    from lemas_tts import _native as nv

    lib = nv.load()
    lib.lemas_debug_attention_variant(request.param)
    yield request.param
    lib.lemas_debug_attention_variant(100)


@pytest.mark.parametrize("B,N,H,ragged", [(1, 97, 4, False), (1, 128, 2, False), (2, 300, 4, True),
                                           (1, 2187, 16, False), (3, 640, 2, True), (2, 129, 1, True),
                                           (64, 300, 16, True)])
def test_attention_matches_fp32(B, N, H, ragged, kernel_variant):
    from lemas_tts import ops

    g = torch.Generator().manual_seed(N + H)
    q = (torch.randn(B, H, N, 64, generator=g) * 1.5).cuda().half()
    k = (torch.randn(B, H, N, 64, generator=g) * 1.5).cuda().half()
    v = torch.randn(B, H, N, 64, generator=g).cuda().half()
    kv_len = None
    if ragged:
        kv_len = torch.tensor(([N, max(1, N // 3), 5] * ((B + 2) // 3))[:B], device="cuda", dtype=torch.int32)
    inner = H * 64
    qk = torch.cat((q.transpose(1, 2).reshape(B * N, inner), k.transpose(1, 2).reshape(B * N, inner)), dim=1).contiguous()
    npad = (N + 63) // 64 * 64
    vt = torch.full((B, H, 64, npad), float("nan"), device="cuda", dtype=torch.float16)  # padding must never be read
    vt[..., :N] = v.transpose(-1, -2)
    out = ops.attention(qk, vt, B, N, H, kv_len)
    torch.cuda.synchronize()
    ref = _ref(q, k, v, kv_len).transpose(1, 2).reshape(B * N, inner)
    got = out.float()
    if kv_len is not None:  # rows of all-padding query tiles are skipped by design; compare valid rows only
        rows = (torch.arange(N, device="cuda")[None] < kv_len[:, None]).reshape(-1)
        got, ref = got[rows], ref[rows]
    err = (got - ref).abs().max().item()
    assert math.isfinite(err) and err < 4e-3, f"max abs err {err}"


@pytest.mark.timeout(120)
def test_attention_multiwave_back_to_back_launches(kernel_variant):
    """Regression (round 1): 2 x 16 x 18 = 576 CTAs (two waves on 148 SMs x 2) launched back to back, last query tile
    with three all-padding warps per key half.  Their lanes polled the barrier independently, lane 0 ran ahead and the
    parity wait of the others was lapped -> the kernel hung.  Must finish, and every launch must give the same bits."""
    from lemas_tts import ops

    B, N, H = 2, 2187, 16
    g = torch.Generator().manual_seed(7)
    qk = (torch.randn(B * N, 2 * H * 64, generator=g) * 1.5).cuda().half()
    npad = (N + 63) // 64 * 64
    vt = torch.randn(B, H, 64, npad, generator=g).cuda().half()
    first = ops.attention(qk, vt, B, N, H, None).clone()
    for _ in range(30):
        out = ops.attention(qk, vt, B, N, H, None)
    torch.cuda.synchronize()
    assert torch.equal(out, first)
    kv_len = torch.tensor([N, 700], device="cuda", dtype=torch.int32)   # ragged: early-exit tiles + partial blocks
    a = ops.attention(qk, vt, B, N, H, kv_len).clone()
    for _ in range(10):
        b = ops.attention(qk, vt, B, N, H, kv_len)
    torch.cuda.synchronize()
    rows = (torch.arange(N, device="cuda")[None] < kv_len[:, None]).reshape(-1)
    assert torch.equal(a[rows], b[rows])
