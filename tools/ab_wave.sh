#!/bin/bash
# usage: tools/ab_wave.sh "workloads": the wavefront renderer off and on (current tree build)
for w in $1; do
  for wv in 0 1; do
    echo -n "$w wavefront=$wv: "; RTC_WAVEFRONT=$wv python tools/profile_frame.py --workload $w --frames 6 | awk '/^frame (1|3|5):/ {printf "%s %s ms %s rays | ", $2, $3, $6} END {print ""}'
  done
done
