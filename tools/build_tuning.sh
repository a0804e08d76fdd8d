#!/bin/bash
# Tuning build of the library (not the product): every K1 geometry behind RXM_K1_VARIANT, extra flags in $1.
# -> re2-modification_b200/build_alt/librxm_<tag>.so ; use with RXM_LIB=... tools/k1_time.py
set -e
cd "$(dirname "$0")/../re2-modification_b200"
TAG=${2:-tuning}
OUT=build_alt/$TAG
mkdir -p $OUT
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DRXM_TUNING $1"
for f in rxm_api rxm_k1 rxm_k1b rxm_k2 rxm_k3 rxm_k4 rxm_tok; do
  nvcc $FLAGS -c csrc/$f.cu -o $OUT/$f.o &
done
g++ -std=c++17 -O2 -fPIC -c csrc/rxm_tables.cpp -o $OUT/rxm_tables.o &
g++ -std=c++17 -O2 -fPIC -c csrc/rxm_plan.cpp -o $OUT/rxm_plan.o &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build_alt/librxm_$TAG.so $OUT/*.o -cudart static
echo built build_alt/librxm_$TAG.so
