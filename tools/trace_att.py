"""Per-block clock64 timeline of the softmax warps of one attention CTA (debug aid)."""
import ctypes as C, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "lemas-tts_b200")]
import torch
from lemas_tts import _native as nv
seq, B2, H, D = 2187, 2, 16, 1024
M = seq * B2
g = torch.Generator(device="cuda").manual_seed(0)
npad = (seq + 63) // 64 * 64
qk = torch.randn(M, 2 * D, device="cuda", generator=g).half()
vt = torch.randn(B2, H, 64, npad, device="cuda", generator=g).half()
out = torch.empty(M, D, device="cuda", dtype=torch.float16)
lib = nv.load()
run = lambda: nv.check(lib.lemas_attention_f16(nv.ptr(qk), 2 * D, nv.ptr(vt), npad, None, nv.ptr(out), B2, seq, H, nv.stream()))
for _ in range(3): run()
torch.cuda.synchronize()
tr = torch.zeros(8, 32, 8, dtype=torch.int64, device="cuda")
lib.lemas_debug_attention_trace.argtypes = [C.c_void_p]
lib.lemas_debug_attention_trace(tr.data_ptr())
run(); torch.cuda.synchronize()
lib.lemas_debug_attention_trace(None)
t = tr.cpu()
t0 = t[:, 0, 0].min()
names = ["wait_s", "ld_S", "max/resc", "exp", "wait_o", "sts+arr"]
for w in (0, 4):
    print(f"warp {w+2} (half {w//4}):  start  " + "  ".join(f"{n:>8s}" for n in names) + "    total")
    for j in range(18):
        s = t[w, j]
        d = [int(s[i + 1] - s[i]) for i in range(6)]
        print(f"  blk {j:2d} {int(s[0]-t0):9d}  " + "  ".join(f"{x:8d}" for x in d) + f"   {int(s[6]-s[0]):6d}")
print("per-block mean over warps (cycles):", [round(float((t[:, 1:17, i+1]-t[:, 1:17, i]).float().mean()),1) for i in range(6)],
      "block period:", round(float((t[:, 16, 0]-t[:, 1, 0]).float().mean())/15, 1))
