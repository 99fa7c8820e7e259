#!/bin/bash
# Round-2 GPU batch A: attention v5 parity + variant timings, full-NFE goldens, regression of the GPU suite, C2 bench.
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
nvidia-smi -L > $O/r02a_gpu.txt 2>&1
echo "== attention tests (v5 default)" | tee $O/r02a_att.log
if timeout 300 python -m pytest tests/test_attention_gpu.py -x -q >> $O/r02a_att.log 2>&1; then echo "v5 parity OK"; else echo "v5 FAILED -> rest on v3"; export LEMAS_ATT_VARIANT=0; fi
tail -5 $O/r02a_att.log
for v in 0 1 2 3 4 sdpa; do timeout 180 python tools/bench_att.py $v C2 C4 C5 C3r 2>&1 | grep -v "^$" | tee -a $O/r02a_bench_att.log; done
echo "== full NFE parity" 
timeout 900 python -m pytest tests/test_fullnfe_gpu.py -x -q -s > $O/r02a_fullnfe.log 2>&1; grep -E "mel-MSE|passed|failed|Error|error" $O/r02a_fullnfe.log | tail -12
echo "== whole gpu suite"
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02a_tests.txt 2>&1; tail -4 $O/r02a_tests.txt
echo "== bench C2"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r02a_bench_C2.json 2> $O/r02a_bench_C2.err; cat $O/r02a_bench_C2.json | head -c 1500
