"""Per-kernel summary of an ncu launch list (`--metrics gpu__time_duration.sum --csv`):
    python tools/launch_list_summary.py profiles/<file>.csv
ncu times are cold-cache and serialised: compare the SHARES with bench.py's `kernels`, not the absolute numbers."""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10 and r[0].isdigit()]
tot = collections.defaultdict(lambda: [0.0, 0])
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("lemas::", "")
    val = float(r[-1].replace(",", ""))
    unit = r[-2]
    us = val / 1e3 if unit in ("ns", "nsecond") else (val if unit.startswith("us") else val * 1e3)
    tot[name][0] += us
    tot[name][1] += 1
total = sum(v[0] for v in tot.values())
print(f"{len(rows)} launches, {total / 1e3:.2f} ms of kernel time (ncu, cold caches, serialised)")
for name, (us, n) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"{us / total * 100:6.2f} %  {us / n:9.2f} us x {n:4d}   {name}")
