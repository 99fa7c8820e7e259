"""Run the attention kernel a few times on realistic inputs (for `ncu -s 2 -c 1`)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "lemas-tts_b200")]
import torch
from lemas_tts import _native as nv
seq = int(sys.argv[1]) if len(sys.argv) > 1 else 2187
B2, H, D = 2, 16, 1024
M = seq * B2
g = torch.Generator(device="cuda").manual_seed(0)
npad = (seq + 63) // 64 * 64
qk = torch.randn(M, 2 * D, device="cuda", generator=g).half()
vt = torch.randn(B2, H, 64, npad, device="cuda", generator=g).half()
out = torch.empty(M, D, device="cuda", dtype=torch.float16)
for _ in range(4):
    nv.check(nv.load().lemas_attention_f16(nv.ptr(qk), 2 * D, nv.ptr(vt), npad, None, nv.ptr(out), B2, seq, H, nv.stream()))
torch.cuda.synchronize()
