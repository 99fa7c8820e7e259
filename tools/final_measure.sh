#!/bin/bash
# Round-end evidence on the GPU box: full GPU test suite, bench lines for every config, ncu launch list of the bench
# command and full captures of the top kernels.  usage: bash tools/final_measure.sh <tag>
tag=${1:-final}
mkdir -p gpurun_out
bash tools/gpu_check.sh 2>&1 | grep -E "exit|passed|failed|Error" | tee gpurun_out/${tag}_tests.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench_C2.json 2> gpurun_out/${tag}_bench_C2.err; echo "bench C2 rc $?"
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${tag}_bench_C2_reference.json 2>/dev/null; echo "reference rc $?"
for w in C4 C5 C3; do python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err; echo "bench $w rc $?"; done
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 400 --csv --log-file gpurun_out/${tag}_launches_bench_C2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1; echo "ncu list rc $?"
ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 2 -c 1 -o gpurun_out/${tag}_attention -f python tools/prof_att.py > /dev/null 2>&1; echo "ncu att rc $?"
for k in out ff1 qkv; do ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 2 -c 1 -o gpurun_out/${tag}_gemm2_$k -f python tools/prof_gemm.py $k > /dev/null 2>&1; done; echo "ncu gemm done"
python tools/bench_ops.py 2>&1 | grep -E "bn256|attention|ln_mod|cuBLAS|SDPA" > gpurun_out/${tag}_bench_ops.log
ls gpurun_out | grep ${tag}
