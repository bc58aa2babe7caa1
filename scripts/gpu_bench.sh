#!/bin/bash
# smoke + bench (+ optional ncu passes) on one B200.  usage: gpu_bench.sh [tag] [ncu]
TAG=${1:-r01}; mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -n 3 gpurun_out/smoke_$TAG.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
tail -n 5 gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json
if [ "$2" == "ncu" ]; then
  # launch list of one whole step: skip weight packing (53) + 3 warm-up steps (187 launches each)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 614 -c 190 --csv \
      --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu1_$TAG.log 2>&1
  echo "ncu launch list exit $?"
  # full capture of the six FormerModule kernels of one layer (after the warm-up forwards)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:former_module -s 468 -c 6 \
      -f -o gpurun_out/prof_$TAG python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu2_$TAG.log 2>&1
  echo "ncu full exit $?"
fi
