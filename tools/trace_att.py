"""clock64 trace of the softmax warps of one attention CTA (blockIdx (1,0,0)).

Builds a tracing variant of the library (attention.cu with -DLEMAS_ATT_TRACE) next to the production one:
    python tools/trace_att.py --build        # here (nvcc, no GPU) -> lemas-tts_b200/lib/liblemas_b200_trace.so
    python tools/trace_att.py [seq]          # on the GPU box: prints per-phase clocks per warp and the block period
Stamps per KV block: 0 loop top, 1 S_j visible, 2 S_j in registers, 3 max / lazy rescale done, 4 exponentials done,
6 P_j stored to TMEM + arrive."""
import ctypes
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "lemas-tts_b200"
LIB = PKG / "lib" / "liblemas_b200_trace.so"


def build():
    srcs = ["common.cu", "attention.cu"]
    cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "-DLEMAS_ATT_TRACE", "-shared", "-o", str(LIB), *[str(PKG / "csrc" / s) for s in srcs],
           "-lcuda"]
    subprocess.run(cmd, check=True)
    print("built", LIB)


def main():
    import torch
    seq = int(sys.argv[1]) if len(sys.argv) > 1 else 2187
    lib = ctypes.CDLL(str(LIB))
    B2, H, D = 2, 16, 1024
    g = torch.Generator(device="cuda").manual_seed(0)
    npad = (seq + 63) // 64 * 64
    qk = torch.randn(B2 * seq, 2 * D, device="cuda", generator=g).half()
    vt = torch.randn(B2, H, 64, npad, device="cuda", generator=g).half()
    out = torch.empty(B2 * seq, D, device="cuda", dtype=torch.float16)
    n_cta = (seq + 127) // 128 * H * B2
    trace = torch.zeros(8 * 32 * 8 + 8 * n_cta, device="cuda", dtype=torch.int64)
    lib.lemas_debug_attention_trace.argtypes = [ctypes.c_void_p]
    lib.lemas_attention_f16.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    lib.lemas_debug_attention_trace(ctypes.c_void_p(trace.data_ptr()))
    for _ in range(200):  # clocks up
        rc = lib.lemas_attention_f16(qk.data_ptr(), 2 * D, vt.data_ptr(), npad, None, out.data_ptr(), B2, seq, H, None)
        assert rc == 0
    torch.cuda.synchronize()
    trace.zero_()
    target = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    trace[7] = target
    torch.cuda.synchronize()
    rc = lib.lemas_attention_f16(qk.data_ptr(), 2 * D, vt.data_ptr(), npad, None, out.data_ptr(), B2, seq, H, None)
    torch.cuda.synchronize()
    rec = trace[8 * 32 * 8:].view(n_cta, 8).cpu()
    t = trace[:8 * 32 * 8].view(8, 32, 8).cpu()
    nb = min((seq + 127) // 128, 32)
    names = ["wait S", "ld S", "max", "exp", "store P"]
    print(f"seq {seq}, traced CTA {target}: {nb} KV blocks; clocks per phase, mean over blocks 3..{nb - 2}")
    print("warp  half sub | " + " | ".join(f"{n:>10s}" for n in names) + " |   period")
    for w in range(8):
        tw = t[w, :nb]
        st = tw[:, [0, 1, 2, 3, 4, 6]]          # stamp 5 is unused since P goes to TMEM
        d = (st[:, 1:] - st[:, :-1]).double()
        per = (tw[1:, 0] - tw[:-1, 0]).double()
        sl = slice(3, nb - 1)
        print(f"{w + 2:4d}  {w // 4:4d} {(w + 2) & 3:3d} | " + " | ".join(f"{d[sl, i].mean().item():10.0f}" for i in range(5))
              + f" | {per[3:nb - 2].mean().item():8.0f}")
    t0 = t[:, 0, 0].min()
    print("first stamp -> last stamp (clk):", (t[:, nb - 1, 6].max() - t0).item())
    cta_report(rec)
    import numpy as np
    out_dir = ROOT / "gpurun_out"
    out_dir.mkdir(exist_ok=True)
    np.save(out_dir / "att_cta_trace.npy", rec.numpy())
    np.save(out_dir / "att_warp_trace.npy", t.numpy())
    print("per-block loop-top times of warp 2 relative to start:", [(t[0, j, 0] - t0).item() for j in range(nb)])


def cta_report(rec):
    import torch
    r = rec[rec[:, 1] > 0]
    t0 = r[:, 1].min()
    ent, loop, merge, ex = [(r[:, i] - t0).double() / 1e3 for i in (1, 2, 3, 4)]
    q = torch.tensor([0.0, 0.1, 0.5, 0.9, 1.0], dtype=torch.double)
    fmt = lambda x: " ".join(f"{v:7.2f}" for v in torch.quantile(x, q).tolist())
    print(f"CTAs {len(r)}: kernel span {ex.max().item():.2f} us (first entry -> last exit, globaltimer)")
    print("  quantiles 0/10/50/90/100 %   [us]")
    print("  entry           :", fmt(ent))
    print("  entry -> loop   :", fmt(loop - ent))
    print("  loop            :", fmt(merge - loop))
    print("  merge -> exit   :", fmt(ex - merge))
    print("  exit            :", fmt(ex))
    order = torch.argsort(ent)
    n1 = min(296, len(r))
    w1, w2 = order[:n1], order[n1:]
    print(f"  first {n1} CTAs: entry <= {ent[w1].max().item():.2f}, exit {ex[w1].min().item():.2f}..{ex[w1].max().item():.2f}")
    if len(w2):
        print(f"  remaining {len(w2)}: entry {ent[w2].min().item():.2f}..{ent[w2].max().item():.2f}, "
              f"exit {ex[w2].min().item():.2f}..{ex[w2].max().item():.2f}")
    per_sm = torch.bincount(r[:, 0].long())
    print(f"  CTAs per SM: min {per_sm[per_sm > 0].min().item()} max {per_sm.max().item()} over {(per_sm > 0).sum().item()} SMs")


if __name__ == "__main__":
    if "--build" in sys.argv:
        build()
    else:
        main()
