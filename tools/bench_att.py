"""Attention kernel variants on the BASELINE shapes: parity against fp32 softmax + CUDA-graph timing.

    python tools/bench_att.py <variant> [shape ...]      shape = C2 | C4 | C5 | C3r (ragged), default C2 C4
variant 0 = v3 (attention.cu), 1.. = v5 (attention5.cu, see lemas_attention_f16), sdpa = torch SDPA (library, context).
One process per variant (a trapped kernel poisons the context); run under `timeout`.
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "lemas-tts_b200"))
from lemas_tts import _native as nv  # noqa: E402

SHAPES = {"C2": (2, 2187, None), "C4": (64, 768, None), "C5": (2, 2814, None),
          "C3r": (8, 3889, [1165, 3889, 951, 2069, 1673, 1591, 2532, 1829])}
H, D = 16, 1024


def timeit(fn, iters=20):
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
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters * 1e3)
    return best


def main():
    variant = sys.argv[1]
    shapes = sys.argv[2:] or ["C2", "C4"]
    lib = nv.load()
    if variant != "sdpa":
        lib.lemas_debug_attention_variant(int(variant))
    for name in shapes:
        if name not in SHAPES and "x" in name:      # ad-hoc uniform shape "B2xSEQ", e.g. 2x4736
            b2, sq = name.split("x")
            SHAPES[name] = (int(b2), int(sq), None)
        B2, seq, lens = SHAPES[name]
        g = torch.Generator(device="cuda").manual_seed(0)
        npad = (seq + 63) // 64 * 64
        M = B2 * seq
        qk = torch.randn(M, 2 * D, device="cuda", generator=g).half()
        vt = torch.randn(B2, H, 64, npad, device="cuda", generator=g).half()
        out = torch.zeros(M, D, device="cuda", dtype=torch.float16)
        kv = None if lens is None else torch.tensor(lens, device="cuda", dtype=torch.int32)
        if lens is None:
            flops = 4.0 * seq * seq * D * B2
        else:
            flops = sum(4.0 * l * l * D for l in lens)
        if variant == "sdpa":
            q = qk[:, :D].view(B2, seq, H, 64).transpose(1, 2).contiguous()
            k = qk[:, D:].view(B2, seq, H, 64).transpose(1, 2).contiguous()
            v = vt[..., :seq].transpose(-1, -2).contiguous()
            us = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v))
            print(f"{name:4s} torch SDPA (dense, no key mask)   {us:9.2f} us  {flops / us / 1e6:8.1f} TFLOP/s", flush=True)
            continue

        def run():
            nv.check(lib.lemas_attention_f16(nv.ptr(qk), 2 * D, nv.ptr(vt), npad, nv.ptr(kv), nv.ptr(out), B2, seq, H,
                                             nv.stream()))
        run()
        torch.cuda.synchronize()
        # parity on a few (batch, head) slices against fp32 softmax
        err = 0.0
        for b, h in ((0, 0), (B2 - 1, H - 1), (B2 // 2, 5)):
            n = seq if lens is None else lens[b]
            q = qk[b * seq:b * seq + n, h * 64:(h + 1) * 64].float()
            k = qk[b * seq:b * seq + n, D + h * 64:D + (h + 1) * 64].float()
            v = vt[b, h, :, :n].float().t()
            ref = torch.softmax(q @ k.t() / 8.0, -1) @ v
            got = out[b * seq:b * seq + n, h * 64:(h + 1) * 64].float()
            err = max(err, (got - ref).abs().max().item())
        us = timeit(run)
        # determinism: a second launch must give the same bits
        first = out.clone()
        run()
        torch.cuda.synchronize()
        same = torch.equal(first, out)
        print(f"{name:4s} variant {variant}  {us:9.2f} us  {flops / us / 1e6:8.1f} TFLOP/s   max|err| {err:.2e}  "
              f"deterministic {same}", flush=True)


if __name__ == "__main__":
    main()
