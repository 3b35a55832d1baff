#!/bin/bash
# A/B timing of library builds kept under ab/<name>/ on one GPU box:  tools/ab.sh "c3 c1" typed cur
# Prints frame 2 (natural tile order, timed), frame 4 (longest-first trial) and frame 7 (after the verdict).
wl="$1"; shift
for rep in 1 2; do
  for name in "$@"; do
    for w in $wl; do
      echo -n "$name $w: "; RTC_LIB_DIR=$PWD/ab/$name python tools/profile_frame.py --workload $w --frames 8 | awk '/^frame (2|4|7):/ {printf "%s %s ms | ", $2, $3} END {print ""}'
    done
  done
done
