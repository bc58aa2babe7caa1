#!/bin/bash
# compute-sanitizer over the split-path kernels (T = 81: one M-tile, T = 150 / 243: two), tiny batches
mkdir -p gpurun_out
for tool in memcheck synccheck; do
for T in 81 243; do
timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_small.py $T fast > gpurun_out/san_${tool}_T$T.log 2>&1; echo "$tool T=$T exit $?"; tail -n 3 gpurun_out/san_${tool}_T$T.log
done; done
