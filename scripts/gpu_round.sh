#!/bin/bash
# tests + bench + per-phase cycles on one B200.  usage: gpu_round.sh <tag>
TAG=${1:-x}; mkdir -p gpurun_out
bash scripts/gpu_tests.sh
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
tail -n 3 gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json
python scripts/phase_profile.py 1024 27 > gpurun_out/phases_$TAG.log 2>&1; echo "phases exit $?"; cat gpurun_out/phases_$TAG.log | tail -n 8
cp gpurun_out/phases.json gpurun_out/phases_$TAG.json 2>/dev/null
