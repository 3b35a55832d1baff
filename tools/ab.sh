#!/bin/bash
# A/B timing of library builds kept under ab/<name>/ on one GPU box:  tools/ab.sh "c3 c1" typed cur
wl="$1"; shift
for rep in 1 2; do
  for name in "$@"; do
    for w in $wl; do
      echo -n "$name $w: "; RTC_LIB_DIR=$PWD/ab/$name python tools/profile_frame.py --workload $w --frames 4 | tail -1
    done
  done
done
