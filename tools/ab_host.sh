#!/bin/bash
# A/B of host-side library builds kept under ab/<name>/ on one GPU box: one-shot phases of c4 and c5, alternating
for rep in 1 2 3; do
  for name in "$@"; do
    echo -n "$name: "; RTC_LIB_DIR=$PWD/ab/$name RTC_TIMING=1 python tools/one_shot_phases.py c4 c5 2>&1 | grep -e "^c[45]:" | tr '\n' ' '; echo
  done
done
