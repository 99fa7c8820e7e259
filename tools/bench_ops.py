"""Op-level timing on the GPU box (CUDA events, back-to-back launches, L2-warm like inside the sampler).
usage: python tools/bench_ops.py [M=4374] [seq=2187]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "lemas-tts_b200")]
import torch

from lemas_tts import _native as nv
from lemas_tts import ops

seq = int(sys.argv[2]) if len(sys.argv) > 2 else 2187
B2 = int(sys.argv[3]) if len(sys.argv) > 3 else 2
M = B2 * seq
D, F, H = 1024, 2048, 16
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)


def r16(*s):
    return (torch.randn(*s, device=dev, generator=g) * 0.5).half()


def r32(*s):
    return torch.randn(*s, device=dev, generator=g)


def timeit(fn, flops=None, bytes_=None, name="", iters=20):
    """`iters` back-to-back launches captured in a CUDA graph (no host launch overhead in the number)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for _ in range(iters):
                fn()
    torch.cuda.synchronize()
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    extra = ""
    if flops:
        extra += f"  {flops / us / 1e6:8.1f} TFLOP/s"
    if bytes_:
        extra += f"  {bytes_ / us / 1e3:8.1f} GB/s"
    print(f"{name:44s} {us:9.2f} us{extra}", flush=True)
    return us


a_d, a_f = r16(M, D), r16(M, F)
w_qkv, w_out, w_ff1, w_ff2 = r16(3 * D, D), r16(D, D), r16(F, D), r16(D, F)
b_d, b_f, b_qkv = r32(D), r32(F), r32(3 * D)
x = r32(M, D)
gate = r32(D)
o16_d = torch.empty(M, D, device=dev, dtype=torch.float16)
o16_f = torch.empty(M, F, device=dev, dtype=torch.float16)
o32_d = torch.empty(M, D, device=dev)
qk = torch.empty(M, 2 * D, device=dev, dtype=torch.float16)
npad = (seq + 63) // 64 * 64
vt = torch.zeros(B2, H, 64, npad, device=dev, dtype=torch.float16)
inv_freq = 1.0 / (10000.0 ** (torch.arange(0, 64, 2).float() / 64))
ang = torch.outer(torch.arange(seq).float(), inv_freq)
rope = torch.stack((ang.cos(), ang.sin()), -1).to(dev).contiguous()

print(f"M={M} seq={seq} B'={B2}")
for bn in (128, 256):
    timeit(lambda: ops.gemm(a_d, w_out, epilogue=nv.EPI_BIAS_F16, bias=b_d, block_n=bn, out16=o16_d),
           2.0 * M * D * D, name=f"gemm K1024 N1024 bias_f16 bn{bn}")
    timeit(lambda: ops.gemm(a_d, w_out, epilogue=nv.EPI_BIAS_F32, bias=b_d, block_n=bn, out32=o32_d),
           2.0 * M * D * D, name=f"gemm K1024 N1024 bias_f32 bn{bn}")
    timeit(lambda: ops.gemm(a_d, w_out, epilogue=nv.EPI_GATE_RESID_F32, bias=b_d, block_n=bn, out32=x, resid=x, gate=gate,
                            seq_len=seq), 2.0 * M * D * D, name=f"gemm K1024 N1024 gate_resid bn{bn}  (to_out)")
    timeit(lambda: ops.gemm(a_f, w_ff2, epilogue=nv.EPI_GATE_RESID_F32, bias=b_d, block_n=bn, out32=x, resid=x, gate=gate,
                            seq_len=seq), 2.0 * M * D * F, name=f"gemm K2048 N1024 gate_resid bn{bn}  (ff2)")
    timeit(lambda: ops.gemm(a_d, w_ff1, epilogue=nv.EPI_GELU_TANH_F16, bias=b_f, block_n=bn, out16=o16_f),
           2.0 * M * D * F, name=f"gemm K1024 N2048 gelu_tanh bn{bn}  (ff1)")
    timeit(lambda: ops.gemm(a_d, w_ff1, epilogue=nv.EPI_BIAS_F16, bias=b_f, block_n=bn, out16=o16_f),
           2.0 * M * D * F, name=f"gemm K1024 N2048 bias_f16 bn{bn}")
    timeit(lambda: ops.gemm(a_d, w_qkv, epilogue=nv.EPI_QKV_ROPE, bias=b_qkv, block_n=bn, out16=qk, rope=rope,
                            rope_cols=D, inner=D, vt=vt, seq_len=seq), 2.0 * M * D * 3 * D,
           name=f"gemm K1024 N3072 qkv_rope bn{bn}  (qkv)")
att_out = torch.empty(M, D, device=dev, dtype=torch.float16)
qk = torch.randn(M, 2 * D, device=dev, generator=g).half()  # unit-variance q, k: scores/8 ~ N(0, 1) like LN'd activations
vt = torch.randn(B2, H, 64, npad, device=dev, generator=g).half()
timeit(lambda: nv.check(nv.load().lemas_attention_f16(nv.ptr(qk), 2 * D, nv.ptr(vt), npad, None, nv.ptr(att_out), B2, seq,
                                                       H, nv.stream())), 4.0 * seq * seq * D * B2, name="attention")
sc, sh = r32(D), r32(D)
ln_out = torch.empty(M, D, device=dev, dtype=torch.float16)
timeit(lambda: nv.check(nv.load().lemas_ln_modulate(nv.ptr(x), nv.ptr(sc), nv.ptr(sh), 0, nv.ptr(ln_out), M, D, seq,
                                                     nv.stream())), bytes_=6.0 * M * D, name="ln_modulate")
# cuBLAS reference points (library, for context only)
wt = w_ff1.t().contiguous()
timeit(lambda: torch.matmul(a_d, wt), 2.0 * M * D * F, name="torch.matmul fp16 K1024 N2048 (cuBLAS)")
wt2 = w_out.t().contiguous()
timeit(lambda: torch.matmul(a_d, wt2), 2.0 * M * D * D, name="torch.matmul fp16 K1024 N1024 (cuBLAS)")
q = r16(B2, H, seq, 64); k = r16(B2, H, seq, 64); v = r16(B2, H, seq, 64)
timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v), 4.0 * seq * seq * D * B2,
       name="torch SDPA fp16 (library)")
