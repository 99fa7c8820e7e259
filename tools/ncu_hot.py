"""Top stall-sampled SASS lines of one kernel in an .ncu-rep: python tools/ncu_hot.py file.ncu-rep [kernel-regex] [topN]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else None
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
cmd = ["ncu", "-i", rep, "--page", "source", "--csv"]
if kre:
    cmd += ["--kernel-name", f"regex:{kre}"]
out = subprocess.run(cmd, capture_output=True, text=True).stdout.splitlines()
# several kernels may follow each other; take the first
blocks, cur = [], []
for line in out:
    if line.startswith('"Kernel Name"'):
        if cur:
            blocks.append(cur)
        cur = [line]
    else:
        cur.append(line)
if cur:
    blocks.append(cur)
b = blocks[0]
print(b[0][:160])
rows = list(csv.reader(b[1:]))
hdr = rows[0]
ia, isrc, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for pos, r in enumerate(rows[1:]):
    try:
        s = int(r[isamp])
    except ValueError:
        continue
    data.append((s, pos, r))
total = sum(d[0] for d in data)
print("total samples", total)
for s, pos, r in sorted(data, key=lambda d: -d[0])[:top]:
    st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:3]
    print(f"{100*s/total:5.1f}%  #{pos:5d} {r[isrc].strip()[:70]:70s} {' '.join(f'{n}:{v}' for v, n in st if v)}")
