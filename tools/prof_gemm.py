"""Run ONE GEMM flavour a few times (for `ncu -s 2 -c 1`): python tools/prof_gemm.py out|ff2|ff1|qkv|bias"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "lemas-tts_b200")]
import torch
from lemas_tts import _native as nv, ops
kind = sys.argv[1]
seq, B2, D, F, H = 2187, 2, 1024, 2048, 16
M = seq * B2
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
r16 = lambda *s: (torch.randn(*s, device=dev, generator=g) * 0.5).half()
r32 = lambda *s: torch.randn(*s, device=dev, generator=g)
a_d, a_f = r16(M, D), r16(M, F)
x, gate = r32(M, D), r32(D)
npad = (seq + 63) // 64 * 64
vt = torch.zeros(B2, H, 64, npad, device=dev, dtype=torch.float16)
ang = torch.outer(torch.arange(seq).float(), 1.0 / (10000.0 ** (torch.arange(0, 64, 2).float() / 64)))
rope = torch.stack((ang.cos(), ang.sin()), -1).to(dev).contiguous()
qk = torch.empty(M, 2 * D, device=dev, dtype=torch.float16)
o16 = torch.empty(M, F, device=dev, dtype=torch.float16)
w = {"out": r16(D, D), "ff2": r16(D, F), "ff1": r16(F, D) * 0.05, "qkv": r16(3 * D, D) * 0.05, "bias": r16(F, D)}[kind]
b = r32(w.shape[0])
for _ in range(4):
    if kind == "out":
        ops.gemm(a_d, w, epilogue=nv.EPI_GATE_RESID_F32, bias=b, block_n=256, out32=x, resid=x, gate=gate, seq_len=seq)
    elif kind == "ff2":
        ops.gemm(a_f, w, epilogue=nv.EPI_GATE_RESID_F32, bias=b, block_n=256, out32=x, resid=x, gate=gate, seq_len=seq)
    elif kind == "ff1":
        ops.gemm(a_d, w, epilogue=nv.EPI_GELU_TANH_F16, bias=b, block_n=256, out16=o16)
    elif kind == "bias":
        ops.gemm(a_d, w, epilogue=nv.EPI_BIAS_F16, bias=b, block_n=256, out16=o16)
    elif kind == "qkv":
        ops.gemm(a_d, w, epilogue=nv.EPI_QKV_ROPE, bias=b, block_n=256, out16=qk, rope=rope, rope_cols=D, inner=D, vt=vt, seq_len=seq)
torch.cuda.synchronize()
