#!/bin/bash
# usage: ab_opt.sh "workloads" names... ; runs each with filter on, and c3 with RTC_SHADOW_FILTER=0
wl="$1"; shift
for rep in 1 2; do
  for name in "$@"; do
    for w in $wl; do
      echo -n "$name $w: "; RTC_LIB_DIR=$PWD/ab/$name python tools/profile_frame.py --workload $w --frames 4 | tail -1
    done
    echo -n "$name c3 nofilter: "; RTC_SHADOW_FILTER=0 RTC_LIB_DIR=$PWD/ab/$name python tools/profile_frame.py --workload c3 --frames 4 | tail -1
  done
done
