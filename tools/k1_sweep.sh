#!/bin/bash
# one gpurun call: every K1 geometry of the tuning build on the config-2 workload (tools/build_tuning.sh first)
cd "$(dirname "$0")/.."
L=re2-modification_b200/build_alt
for fl in 0 8; do
  for v in 0 1 2 3 4 5 6 8 9; do
    RXM_LIB=$L/librxm_tuning.so RXM_K1_VARIANT=$v RXM_K1_FLAGS=$fl python tools/k1_time.py 1000000 20 2>&1 | tail -1
  done
done
for v in 0 1 8; do
  RXM_LIB=$L/librxm_probe.so RXM_K1_VARIANT=$v python tools/k1_time.py 1000000 20 2>&1 | tail -1
done
