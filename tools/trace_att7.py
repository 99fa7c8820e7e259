"""clock64 trace of attention v5 (csrc/attention5.cu): the 16 softmax warps and the two MMA issuers of one CTA
during its first work item.

    python tools/trace_att5.py --build          # here (nvcc, no GPU) -> lemas-tts_b200/lib/liblemas_b200_trace.so
    python tools/trace_att5.py <variant> [seq] [cta]   # on the GPU box
Softmax stamps per KV block: 0 loop top, 1 S_j visible, 2 S_j in registers, 3 max / lazy rescale done, 4 exponentials
done, 6 P_j stored + arrive.  MMA stamps per (block, half): p_full observed, P V + next S issued."""
import ctypes
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "lemas-tts_b200"
LIB = PKG / "lib" / "liblemas_b200_trace.so"


def build():
    srcs = ["common.cu", "attention.cu", "attention7.cu"]
    cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "-DLEMAS_ATT_TRACE", "-shared", "-o", str(LIB), *[str(PKG / "csrc" / s) for s in srcs],
           "-lcuda"]
    subprocess.run(cmd, check=True)
    print("built", LIB)


def main():
    import torch
    variant = int(sys.argv[1])
    seq = int(sys.argv[2]) if len(sys.argv) > 2 else 2187
    target = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    lib = ctypes.CDLL(str(LIB))
    B2, H, D = 2, 16, 1024
    g = torch.Generator(device="cuda").manual_seed(0)
    npad = (seq + 63) // 64 * 64
    qk = torch.randn(B2 * seq, 2 * D, device="cuda", generator=g).half()
    vt = torch.randn(B2, H, 64, npad, device="cuda", generator=g).half()
    out = torch.empty(B2 * seq, D, device="cuda", dtype=torch.float16)
    trace = torch.zeros(4096 + 512, device="cuda", dtype=torch.int64)
    lib.lemas_debug_attention_trace.argtypes = [ctypes.c_void_p]
    lib.lemas_attention_f16.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    lib.lemas_debug_attention_variant(variant)
    lib.lemas_debug_attention_trace(ctypes.c_void_p(trace.data_ptr()))
    trace[7] = -1
    for _ in range(100):  # clocks up
        assert lib.lemas_attention_f16(qk.data_ptr(), 2 * D, vt.data_ptr(), npad, None, out.data_ptr(), B2, seq, H, None) == 0
    torch.cuda.synchronize()
    trace.zero_()
    trace[7] = target
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    assert lib.lemas_attention_f16(qk.data_ptr(), 2 * D, vt.data_ptr(), npad, None, out.data_ptr(), B2, seq, H, None) == 0
    e1.record()
    torch.cuda.synchronize()
    t = trace[:4096].view(16, 32, 8).cpu()
    m = trace[4096:4096 + 256].view(2, 32, 2, 2).cpu()
    nb = min((seq + 127) // 128, 32)
    names = ["wait S", "ld S", "max", "exp", "store P"]
    print(f"variant {variant} seq {seq} traced CTA {target} (first item): {nb} KV blocks, kernel {e0.elapsed_time(e1) * 1e3:.1f} us "
          f"(traced); clocks per phase, mean over blocks 3..{nb - 2}")
    print("sw tile half sub | " + " | ".join(f"{n:>8s}" for n in names) + " |  period | first S at")
    t0 = t[:, 0, 0][t[:, 0, 0] > 0].min()
    for w in range(16):
        tw = t[w, :nb]
        if tw[0, 0] == 0:
            continue
        st = tw[:, [0, 1, 2, 3, 4, 6]]
        d = (st[:, 1:] - st[:, :-1]).double()
        per = (tw[1:, 0] - tw[:-1, 0]).double()
        sl = slice(3, nb - 1)
        print(f"{w:2d} {w >> 3:4d} {(w >> 2) & (3 if variant >= 7 else 1):4d} {w & 3:3d} | " + " | ".join(f"{d[sl, i].mean().item():8.0f}" for i in range(5))
              + f" | {per[3:nb - 2].mean().item():7.0f} | {(tw[0, 1] - t0).item():6d}")
    print("item span (first stamp -> last P stored):", (t[:, nb - 1, 6].max() - t0).item(), "clk")
    if variant >= 7:
        ss = trace[4096:4096 + 128].view(32, 4).cpu().double()[:nb]
        ps = trace[4096 + 128:4096 + 256].view(32, 4).cpu().double()[:nb]
        sl = slice(3, nb - 1)
        print("S issuer: wait k_full %.0f | wait pv_done x4 %.0f | issue + commits %.0f | period %.0f" % (
            (ss[sl, 1] - ss[sl, 0]).mean(), (ss[sl, 2] - ss[sl, 1]).mean(), (ss[sl, 3] - ss[sl, 2]).mean(),
            (ss[1:, 0] - ss[:-1, 0])[3:nb - 2].mean()))
        print("P V issuer 0: wait v_full %.0f | wait p_full %.0f | issue + commits %.0f | period %.0f" % (
            (ps[sl, 1] - ps[sl, 0]).mean(), (ps[sl, 2] - ps[sl, 1]).mean(), (ps[sl, 3] - ps[sl, 2]).mean(),
            (ps[1:, 0] - ps[:-1, 0])[3:nb - 2].mean()))
        print("S issue start of block g relative to P V issuer 0 done with block g-2:",
              [int(x) for x in (ss[2:nb, 2] - ps[:nb - 2, 3])[:12].tolist()])
        return
    if variant >= 4:  # v6: P V issuers; stamps per (block, half): P observed, P V issued
        mm = trace[4096:4096 + 128].view(32, 2, 2).cpu().double()[:nb]
        for x in range(2):
            print("P V issuer %d (v6), mean clocks over blocks 3..: issue %.0f | period %.0f" % (
                x, (mm[3:nb - 1, x, 1] - mm[3:nb - 1, x, 0]).mean(), (mm[1:, x, 0] - mm[:-1, x, 0])[3:nb - 2].mean()))
        return
    print("MMA issuers: mean clocks from p_full observed to P V + S issued, and period between services")
    for tile in range(2):
        for x in range(2):
            seen, done = m[tile, :nb, x, 0].double(), m[tile, :nb, x, 1].double()
            if seen[0] == 0:
                continue
            print(f"  tile {tile} half {x}: issue {(done - seen)[3:nb - 1].mean().item():6.0f} clk, period "
                  f"{(seen[1:] - seen[:-1])[3:nb - 2].mean().item():7.0f} clk, first at {(seen[0] - t0.double()).item():7.0f}")
    # phase of each pipeline's S-ready times relative to tile 0 half A, block by block (stagger actually achieved)
    ref = t[0, :nb, 1].double()
    for w in (4, 8, 12):
        off = (t[w, :nb, 1].double() - ref)[3:nb - 1]
        print(f"  S-ready offset of sw {w} (tile {w >> 3} half {(w >> 2) & 1}) vs sw 0: mean {off.mean().item():7.0f}  "
              f"min {off.min().item():7.0f}  max {off.max().item():7.0f}")


if __name__ == "__main__":
    if "--build" in sys.argv:
        build()
    else:
        main()
