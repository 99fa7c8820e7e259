#!/bin/bash
# Round-2 GPU batch D: attention v6 (double-buffered scores) parity, timings, trace.
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
L=$O/r02d_bench_att.log; : > $L
for v in 5 4 6; do
  echo "== tests variant $v" >> $L
  LEMAS_ATT_VARIANT=$v timeout 300 python -m pytest tests/test_attention_gpu.py -x -q 2>&1 | tail -3 >> $L
done
for v in 0 4 5 6 sdpa; do timeout 180 python tools/bench_att.py $v C2 C4 C5 C3r 2>&1 | grep -E "variant|SDPA|rror" >> $L; done
for d in 0 200 800; do echo "dephase $d" >> $L; LEMAS_A5_DEPHASE_HALF=$d timeout 120 python tools/bench_att.py 5 C2 C4 2>&1 | grep variant >> $L; done
cat $L
timeout 120 python tools/trace_att5.py 5 2187 5 > $O/r02d_trace_v6.txt 2>&1; cat $O/r02d_trace_v6.txt
