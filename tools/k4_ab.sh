#!/bin/bash
# one gpurun call: K4 tuning builds (tools/build_tuning.sh "" tag "<k4 flags>") A/B on configs 3 and 5, each twice
cd "$(dirname "$0")/.."
L=re2-modification_b200/build_alt
for rep in 1 2; do
for t in ${TAGS}; do
  echo "== $t rep $rep"
  RXM_LIB=$L/librxm_$t.so python tools/mfa_time.py config3 k4 2>&1 | grep "ms/step" | head -1
  [ -n "$SKIP5" ] || { RXM_LIB=$L/librxm_$t.so python tools/mfa_time.py config5 k4 2>&1 | grep "ms/step" | tr "\n" " "; echo; }
done
done
