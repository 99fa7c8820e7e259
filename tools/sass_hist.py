"""Histogram of the Blackwell-specific SASS opcodes per kernel of liblemas_b200.so (evidence that the hot kernels are
tcgen05 / TMEM / TMA code):  python tools/sass_hist.py > profiles/r02_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "lemas-tts_b200" / "lib" / "liblemas_b200.so"
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCATOMSWS", "SYNCS", "MUFU.EX2",
        "MUFU.TANH", "FFMA2", "FADD2", "FMNMX3", "HMMA", "ELECT", "UCGABAR", "ACQBULK", "USETMAXREG"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur)
            per[cur] = collections.Counter()
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            per[cur]["_total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    per[cur][k] += 1
    tot = collections.Counter()
    print(f"# cuobjdump -sass {LIB.name}: instruction counts per kernel (only kernels that use tensor cores / TMEM / TMA)")
    for name, c in per.items():
        hits = {k: v for k, v in c.items() if k != "_total" and k in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTCBAR")}
        if not hits:
            continue
        tot.update({k: v for k, v in c.items() if k != "_total"})
        print(f"{name}\n    {c['_total']} instructions: " + "  ".join(f"{k} {c[k]}" for k in KEYS if c[k]))
    print("TOTAL over the kernels above: " + "  ".join(f"{k} {tot[k]}" for k in KEYS if tot[k]))


if __name__ == "__main__":
    main()
