#!/bin/bash
# Round-2 GPU batch L: bench with in-graph kernel timing, ncu launch list with graph-node profiling, sanitizers
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 900 python bench.py --steps 5 --warmup 3 > $O/r02l_bench_C2.json 2> $O/r02l_bench_C2.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02l_bench_C2.json").read())
print({k: d[k] for k in ("value", "ms_per_step", "x_realtime")}, d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["launch_ms"])
for k, v in d["kernels"].items(): print(" ", k, v)
print(d["profiled_step_ms"], d.get("c4_sharded", {}).get("seconds"), d["clocks"])
PY
tail -2 $O/r02l_bench_C2.err
timeout 900 ncu --graph-profiling node --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/r02l_launches_bench_C2.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-c4 > $O/r02l_ncu_bench.log 2>&1; tail -2 $O/r02l_ncu_bench.log; wc -l $O/r02l_launches_bench_C2.csv
for tool in racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python -m pytest tests/test_attention_gpu.py tests/test_gemm_gpu.py -q -x -k "matches_fp32 or gate_resid or qkv_rope" > $O/r02l_sanitizer_$tool.log 2>&1; tail -4 $O/r02l_sanitizer_$tool.log
done
