#!/bin/bash
# one gpurun call: K4 tuning builds (tools/build_tuning.sh "" tag "<k4 flags>") on configs 3 and 5
cd "$(dirname "$0")/.."
L=re2-modification_b200/build_alt
for t in ${TAGS}; do
  echo "== $t"
  RXM_LIB=$L/librxm_$t.so python tools/mfa_time.py config3 k4 2>&1 | grep "ms/step" | head -1
  RXM_LIB=$L/librxm_$t.so python tools/mfa_time.py config5 k4 2>&1 | grep "ms/step"
done
