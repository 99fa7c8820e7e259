"""QKV epilogue dissection (graph-timed)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "lemas-tts_b200"), str(ROOT / "tools")]
import torch
from lemas_tts import _native as nv, ops
import importlib
seq, B2, D, H = 2187, 2, 1024, 16
M = seq * B2
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
a = (torch.randn(M, D, device=dev, generator=g) * 0.5).half()
w = (torch.randn(3 * D, D, device=dev, generator=g) * 0.03).half()
b = torch.randn(3 * D, device=dev, generator=g)
qk = torch.empty(M, 2 * D, device=dev, dtype=torch.float16)
o = torch.empty(M, 3 * D, device=dev, dtype=torch.float16)
npad = (seq + 63) // 64 * 64
vt = torch.zeros(B2, H, 64, npad, device=dev, dtype=torch.float16)
ang = torch.outer(torch.arange(seq).float(), 1.0 / (10000.0 ** (torch.arange(0, 64, 2).float() / 64)))
rope = torch.stack((ang.cos(), ang.sin()), -1).to(dev).contiguous()

def timeit(fn, name, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream(); graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for _ in range(iters): fn()
    torch.cuda.synchronize(); graph.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); graph.replay(); e1.record(); torch.cuda.synchronize()
    print(f"{name:50s} {e0.elapsed_time(e1)/iters*1e3:8.2f} us", flush=True)

timeit(lambda: ops.gemm(a, w, epilogue=nv.EPI_BIAS_F16, bias=b, block_n=256, out16=o), "N3072 bias_f16 (pair)")
timeit(lambda: ops.gemm(a, w, epilogue=nv.EPI_QKV_ROPE, bias=b, block_n=256, out16=qk, rope=rope, rope_cols=D, inner=D, vt=vt, seq_len=seq), "qkv_rope full (pair)")
timeit(lambda: ops.gemm(a, w, epilogue=nv.EPI_QKV_ROPE, bias=b, block_n=256, out16=qk, rope=rope, rope_cols=0, inner=D, vt=vt, seq_len=seq), "qkv_rope rope_cols=0 (pair)")
# only q,k columns: n = 2*inner is not allowed by the validator, so time V-only via a weight slice trick is skipped
timeit(lambda: ops.gemm(a, w, epilogue=nv.EPI_QKV_ROPE, bias=b, block_n=128, out16=qk, rope=rope, rope_cols=D, inner=D, vt=vt, seq_len=seq), "qkv_rope full (single-CTA bn128)")
