#!/bin/bash
# Tuning builds of the library (not the product): every K1 geometry behind RXM_K1_VARIANT; $1 = extra flags for
# rxm_k1.cu, $3 = extra flags for rxm_k4.cu (optional), $2 = tag -> re2-modification_b200/build_alt/librxm_<tag>.so ; use with RXM_LIB=... tools/k1_time.py
set -e
cd "$(dirname "$0")/../re2-modification_b200"
TAG=${2:-tuning}
OUT=build_alt/common
mkdir -p $OUT build_alt/$TAG
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DRXM_TUNING"
for f in rxm_api rxm_k1b rxm_k2 rxm_k3 rxm_k4 rxm_tok; do
  [ $OUT/$f.o -nt csrc/$f.cu ] || nvcc $FLAGS -c csrc/$f.cu -o $OUT/$f.o &
done
[ $OUT/rxm_tables.o -nt csrc/rxm_tables.cpp ] || g++ -std=c++17 -O2 -fPIC -c csrc/rxm_tables.cpp -o $OUT/rxm_tables.o &
[ $OUT/rxm_plan.o -nt csrc/rxm_plan.cpp ] || g++ -std=c++17 -O2 -fPIC -c csrc/rxm_plan.cpp -o $OUT/rxm_plan.o &
nvcc $FLAGS $1 -c csrc/rxm_k1.cu -o build_alt/$TAG/rxm_k1.o &
if [ -n "$3" ]; then nvcc $FLAGS $3 -c csrc/rxm_k4.cu -o build_alt/$TAG/rxm_k4.o & fi
wait
OBJS=$(ls $OUT/*.o | grep -v "rxm_k1.o$")
if [ -n "$3" ]; then OBJS=$(echo "$OBJS" | grep -v "rxm_k4.o$"); OBJS="$OBJS build_alt/$TAG/rxm_k4.o"; fi
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build_alt/librxm_$TAG.so $OBJS build_alt/$TAG/rxm_k1.o -cudart static
echo built build_alt/librxm_$TAG.so
