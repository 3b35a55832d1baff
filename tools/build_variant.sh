#!/bin/bash
# tools/build_variant.sh <name> <sed-script-file-or-empty> [extra nvcc defines]: build a variant of the current tree into ab/<name>/
set -e
name=$1; script=$2; defs=$3
rm -rf /tmp/variant_$name; mkdir -p /tmp/variant_$name
cp -r /root/repo/ray_tracer_challenge_b200 /root/repo/include /tmp/variant_$name/
rm -rf /tmp/variant_$name/ray_tracer_challenge_b200/build /tmp/variant_$name/ray_tracer_challenge_b200/*.so
if [ -n "$script" ]; then (cd /tmp/variant_$name && python3 "$script"); fi
(cd /tmp/variant_$name && RTC_NVCC_DEFINES="$defs" python -m ray_tracer_challenge_b200.build > /tmp/variant_$name/build.log 2>&1) || { tail -20 /tmp/variant_$name/build.log; exit 1; }
mkdir -p /root/repo/ab/$name && cp /tmp/variant_$name/ray_tracer_challenge_b200/*.so /root/repo/ab/$name/
echo built ab/$name
