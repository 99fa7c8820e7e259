#!/bin/bash
# Runs each GPU test file in its own process (a trapped kernel poisons the CUDA context) and keeps logs.
# usage (on the GPU box): bash tools/gpu_check.sh [file ...]
mkdir -p gpurun_out
files=("$@")
[ ${#files[@]} -eq 0 ] && files=(tests/test_ops_gpu.py tests/test_frontend_gpu.py tests/test_prosody_gpu.py tests/test_gemm_gpu.py tests/test_attention_gpu.py tests/test_vocos_gpu.py tests/test_sampler_gpu.py tests/test_dropin_gpu.py tests/test_fullsize_gpu.py)
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
rc=0
for f in "${files[@]}"; do
  name=$(basename "$f" .py)
  timeout 900 python -m pytest "$f" -m gpu -q -s --timeout 600 -p no:cacheprovider > "gpurun_out/$name.log" 2>&1
  r=$?
  echo "== $f exit $r"; tail -n 30 "gpurun_out/$name.log"
  [ $r -ne 0 ] && rc=$r
done
exit $rc
