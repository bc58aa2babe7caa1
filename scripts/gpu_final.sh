#!/bin/bash
# Round-end evidence on one B200: full GPU test suite, smoke, the default bench line, ncu launch list of the bench
# command, ncu --set full of one layer's kernels (+ head).  usage: gpu_final.sh <tag>
TAG=${1:-final}; mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then bash scripts/gpu_tests.sh 2>&1 | grep -E "exit|passed|failed|error"; fi
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -n 3 gpurun_out/smoke_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"; tail -n 2 gpurun_out/bench_$TAG.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_$TAG.json 2> gpurun_out/bench_reference_$TAG.err; echo "reference arm exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras --no-sweep > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "launch list exit $?"
timeout 600 ncu --set full --clock-control none -k regex:"former_module_kernel|fusion_kernel" -s 14 -c 7 -o gpurun_out/full_$TAG -f python bench.py --steps 1 --warmup 1 --no-cpu --no-extras --no-sweep > gpurun_out/ncu_full_$TAG.log 2>&1; echo "full capture exit $?"
ncu -i gpurun_out/full_$TAG.ncu-rep --page raw --csv > gpurun_out/full_$TAG.csv 2>/dev/null; rm -f gpurun_out/full_$TAG.ncu-rep   # (the merge-back limit is 64 MiB)
timeout 600 ncu --set full --clock-control none -k regex:"head_tc_kernel|features_kernel|limb_tiles" -c 4 -o gpurun_out/full_small_$TAG -f python bench.py --steps 1 --warmup 1 --no-cpu --no-extras --no-sweep > gpurun_out/ncu_full_small_$TAG.log 2>&1; echo "small capture exit $?"
ncu -i gpurun_out/full_small_$TAG.ncu-rep --page raw --csv > gpurun_out/full_small_$TAG.csv 2>/dev/null; rm -f gpurun_out/full_small_$TAG.ncu-rep
timeout 300 python scripts/phase_profile.py 1024 27 > gpurun_out/phases_$TAG.log 2>&1; echo "phases exit $?"
