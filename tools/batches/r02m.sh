#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 900 python bench.py --steps 5 --warmup 3 --no-c4 > $O/r02m_bench_C2.json 2> $O/r02m_bench_C2.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02m_bench_C2.json").read())
print({k: d[k] for k in ("value", "ms_per_step", "x_realtime")}, d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["launch_ms"])
for k, v in d["kernels"].items(): print(" ", k, v)
print(d["profiled_step_ms"], d["clocks"], d.get("cpu_baseline", {}).get("value"))
PY
tail -2 $O/r02m_bench_C2.err
