#!/bin/bash
# Round-2 GPU batch O: folded LayerNorm — parity of the whole suite, C2 bench with and without
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x -s > $O/r02o_tests.txt 2>&1; grep -E "full_C|passed|failed|Error|assert" $O/r02o_tests.txt | tail -12
for f in 1 0; do
  LEMAS_FUSED_LN=$f timeout 600 python bench.py --steps 5 --warmup 3 --no-c4 --no-cpu-baseline > $O/r02o_bench_C2_fused$f.json 2> $O/r02o_bench_C2_fused$f.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r02o_bench_C2_fused$f.json").read())
print("fused=$f", {k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["clocks"]["sm_mhz"])
for k, v in d["kernels"].items(): print("   ", k, v)
PY
done
