#!/bin/bash
# Build a variant of the library with extra nvcc flags for attention.cu only (experiments):
#   bash tools/build_variant.sh <name> "<flags>"  ->  lemas-tts_b200/lib/liblemas_b200_<name>.so  (use with LEMAS_B200_LIB=...)
set -e
cd "$(dirname "$0")/../lemas-tts_b200"
name=$1; shift
src=${SRC:-attention}
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $@ -c csrc/$src.cu -o build/${src}_$name.o
objs=$(ls build/*.o | grep -v "_[a-z0-9]*\.o$" | grep -v "build/$src.o")
nvcc -shared -o lib/liblemas_b200_$name.so $objs build/${src}_$name.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC
echo built lib/liblemas_b200_$name.so
