#!/bin/bash
# one gpurun call: K1 tuning builds (tools/build_tuning.sh) x geometries on the config-2 workload
cd "$(dirname "$0")/.."
L=re2-modification_b200/build_alt
for t in ${TAGS:-tuning}; do
  for v in ${VARIANTS:-0 7}; do
    RXM_LIB=$L/librxm_$t.so RXM_K1_VARIANT=$v python tools/k1_time.py 1000000 20 2>&1 | tail -1
  done
done
