#!/bin/bash
# Round-2 GPU batch B: attention v5 scheduling experiments (control-warp position, pipeline stagger) + traces.
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
L=$O/r02b_bench_att.log; : > $L
for cfg in "600 300" "0 0" "1000 500" "1300 650" "500 1000"; do
  set -- $cfg
  for v in 1 2 3; do
    echo "dephase half $1 tile $2" >> $L
    LEMAS_A5_DEPHASE_HALF=$1 LEMAS_A5_DEPHASE_TILE=$2 timeout 120 python tools/bench_att.py $v C2 2>&1 | grep variant >> $L
  done
done
timeout 60 python tools/bench_att.py 0 C2 | grep variant >> $L
cat $L
for v in 1 2; do timeout 120 python tools/trace_att5.py $v 2187 5 > $O/r02b_trace_v$v.txt 2>&1; cat $O/r02b_trace_v$v.txt; done
LEMAS_A5_DEPHASE_HALF=1300 LEMAS_A5_DEPHASE_TILE=650 timeout 120 python tools/trace_att5.py 2 2187 5 > $O/r02b_trace_v2_stagger.txt 2>&1; cat $O/r02b_trace_v2_stagger.txt
