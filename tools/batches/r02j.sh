#!/bin/bash
# Round-2 GPU batch J: sharded synthesis, reference scripts unchanged, full-NFE goldens, bench C2 (+ c4_sharded), reference arm
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_sharded_gpu.py tests/test_reference_scripts_gpu.py tests/test_fullnfe_gpu.py -x -q -s > $O/r02j_newtests.log 2>&1; grep -E "mel-MSE|passed|failed|Error|error|assert" $O/r02j_newtests.log | tail -20
timeout 900 python bench.py --steps 5 --warmup 3 > $O/r02j_bench_C2.json 2> $O/r02j_bench_C2.err; tail -c 2500 $O/r02j_bench_C2.json; tail -3 $O/r02j_bench_C2.err
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > $O/r02j_bench_C2_reference.json 2> $O/r02j_bench_C2_reference.err; cat $O/r02j_bench_C2_reference.json; tail -3 $O/r02j_bench_C2_reference.err
