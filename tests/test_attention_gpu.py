"""GPU parity of the tcgen05 attention kernel (csrc/attention.cu) through the C ABI.

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


@pytest.mark.parametrize("B,N,H,ragged", [(1, 97, 4, False), (1, 128, 2, False), (2, 300, 4, True),
                                           (1, 2187, 16, False), (3, 640, 2, True), (2, 129, 1, True)])
def test_attention_matches_fp32(B, N, H, ragged):
    from lemas_tts import ops

    g = torch.Generator().manual_seed(N + H)
    q = (torch.randn(B, H, N, 64, generator=g) * 1.5).cuda().half()
    k = (torch.randn(B, H, N, 64, generator=g) * 1.5).cuda().half()
    v = torch.randn(B, H, N, 64, generator=g).cuda().half()
    kv_len = None
    if ragged:
        kv_len = torch.tensor([N, max(1, N // 3), 5][:B], device="cuda", dtype=torch.int32)
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
