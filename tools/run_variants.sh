#!/bin/bash
# usage: tools/run_variants.sh [profile_run args]  -- runs every variant under libacm_b200/_lib/var
for d in libacm_b200/_lib/var/*/; do
  n=$(basename $d)
  echo "== $n"
  ACM_B200_LIB=$d/libacm_b200.so timeout 300 python tools/profile_run.py "$@" 2>&1 | grep -E "^ms|Mcycles|Error|error" | tail -4
done
