#!/bin/bash
# usage: tools/ab_stream.sh "workloads" names... : each library build with the streaming kernel off and on
wl="$1"; shift
for name in "$@"; do
  for w in $wl; do
    for st in 0 1; do
      echo -n "$name $w stream=$st: "; RTC_STREAM=$st RTC_LIB_DIR=$PWD/ab/$name python tools/profile_frame.py --workload $w --frames 8 | awk '/^frame (2|4|7):/ {printf "%s %s ms | ", $2, $3} END {print ""}'
    done
  done
done
