#!/bin/bash
# Round-end evidence on the GPU box: full GPU test suite, bench lines for every config, ncu launch list of the bench
# command, full captures of the top kernels, sanitizer pass over the new kernels.  usage: bash tools/final_measure.sh <tag>
tag=${1:-final}
K="timeout -s KILL"
mkdir -p gpurun_out
bash tools/gpu_check.sh 2>&1 | grep -E "exit|passed|failed|Error" | tee gpurun_out/${tag}_tests.txt
$K 400 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench_C2.json 2> gpurun_out/${tag}_bench_C2.err; echo "bench C2 rc $?"
$K 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${tag}_bench_C2_reference.json 2>/dev/null; echo "reference rc $?"
for w in C4 C5 C3; do $K 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err; echo "bench $w rc $?"; done
$K 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 400 --csv --log-file gpurun_out/${tag}_launches_bench_C2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1; echo "ncu list rc $?"
$K 200 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 2 -c 1 -o gpurun_out/${tag}_attention -f python tools/prof_att.py > /dev/null 2>&1; echo "ncu att rc $?"
for k in out ff1 qkv; do $K 200 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 2 -c 1 -o gpurun_out/${tag}_gemm2_$k -f python tools/prof_gemm.py $k > /dev/null 2>&1; done; echo "ncu gemm done"
$K 200 python tools/bench_ops.py 2>&1 | grep -E "bn256|attention|ln_mod|cuBLAS|SDPA" > gpurun_out/${tag}_bench_ops.log
$K 100 python tools/bench_prosody.py 32 > gpurun_out/${tag}_bench_prosody.log 2>&1
$K 100 python tools/trace_att.py 2187 6 > gpurun_out/${tag}_attention_trace_cta6.txt 2>&1
$K 100 python tools/trace_att.py 2187 229 > gpurun_out/${tag}_attention_trace_cta229.txt 2>&1
$K 500 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_attention_gpu.py tests/test_frontend_gpu.py tests/test_prosody_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | tail -8 > gpurun_out/${tag}_compute_sanitizer_memcheck.log
ls gpurun_out | grep ${tag}
