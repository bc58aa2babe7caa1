#!/bin/bash
# per-kernel launch times (ncu, serialised) at the given frame counts, 256 clips.  usage: gpu_launches.sh <tag> T...
TAG=$1; shift; mkdir -p gpurun_out
for T in "$@"; do
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_${TAG}_T$T.csv python bench.py --frames $T --batch ${BATCH:-256} --steps 1 --warmup 1 --no-cpu --no-extras --no-sweep > gpurun_out/ncu_T$T.log 2>&1; echo T$T rc $?
done
