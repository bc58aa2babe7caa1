#!/bin/bash
# compute-sanitizer (memcheck, synccheck) over tiny forwards of every path: fused (T = 27), split (T = 81, 243), exact mode
mkdir -p gpurun_out
for tool in memcheck synccheck; do
for cfg in "27 fast" "81 fast" "243 fast" "27 exact"; do
set -- $cfg
timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_small.py $1 $2 > gpurun_out/san_${tool}_T$1_$2.log 2>&1; echo "$tool T=$1 $2 exit $?: $(tail -n 1 gpurun_out/san_${tool}_T$1_$2.log)"
done; done
