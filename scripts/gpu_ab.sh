#!/bin/bash
# A/B: per-phase cycles of the default build and of every variant library under variants/.  usage: gpu_ab.sh <tag>
TAG=${1:-x}; mkdir -p gpurun_out
python scripts/phase_profile.py 1024 27 > gpurun_out/phases_${TAG}_base.log 2>&1; echo "base exit $?"; tail -n 6 gpurun_out/phases_${TAG}_base.log
for so in variants/*.so; do
  n=$(basename $so .so)
  KASF_LIB=$PWD/$so python scripts/phase_profile.py 1024 27 > gpurun_out/phases_${TAG}_$n.log 2>&1; echo "$n exit $?"; tail -n 6 gpurun_out/phases_${TAG}_$n.log
done
