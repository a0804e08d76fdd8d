cd /root/repo
L=re2-modification_b200/build_alt
for rep in 1 2; do
for t in ${TAGS:-pol0 pol64}; do
  echo "== $t rep $rep"
  RXM_LIB=$L/librxm_$t.so python tools/mfa_time.py config3 k4 2>&1 | grep "ms/step" | head -1
  RXM_LIB=$L/librxm_$t.so python tools/mfa_time.py config5 k4 2>&1 | grep "ms/step" | tr "\n" " "; echo
done
done
