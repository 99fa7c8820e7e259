#!/bin/bash
# Round-2 GPU batch K: whole GPU suite (row skipping, scripts, sharding, full-NFE goldens), C3 with / without row skipping
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -s > $O/r02k_tests.txt 2>&1; grep -E "mel-MSE|passed|failed|Error|assert" $O/r02k_tests.txt | tail -30
timeout 600 python bench.py --workload C3 --steps 2 --warmup 1 --no-cpu-baseline > $O/r02k_bench_C3.json 2> $O/r02k_bench_C3.err; head -c 700 $O/r02k_bench_C3.json; echo; tail -2 $O/r02k_bench_C3.err
timeout 600 python bench.py --workload C3 --steps 2 --warmup 1 --no-cpu-baseline --all-rows > $O/r02k_bench_C3_allrows.json 2> $O/r02k_bench_C3_allrows.err; head -c 700 $O/r02k_bench_C3_allrows.json; echo
