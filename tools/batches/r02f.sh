#!/bin/bash
# Round-2 GPU batch F: attention v7 (four key parts, double-buffered scores) parity, timings, trace.
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
L=$O/r02f_bench_att.log; : > $L
for v in 8 7 9; do
  echo "== tests variant $v" >> $L
  LEMAS_ATT_VARIANT=$v timeout 300 python -m pytest tests/test_attention_gpu.py -x -q 2>&1 | tail -3 >> $L
done
for v in 0 7 8 9; do timeout 180 python tools/bench_att.py $v C2 C4 C5 C3r 2>&1 | grep -E "variant|SDPA|rror" >> $L; done
for d in 0 150 300; do echo "dephase $d" >> $L; LEMAS_A5_DEPHASE_HALF=$d timeout 120 python tools/bench_att.py 8 C2 C4 2>&1 | grep variant >> $L; done
cat $L
timeout 120 python tools/trace_att5.py 8 2187 5 > $O/r02f_trace_v7.txt 2>&1; cat $O/r02f_trace_v7.txt
