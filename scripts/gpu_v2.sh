#!/bin/bash
# Two-tiles-in-flight kernel: quick sanity first (a protocol error hangs: bounded by timeouts), then the module and
# forward parity tests, then the bench.  usage: gpu_v2.sh <tag> [nobench]
TAG=${1:-x}; mkdir -p gpurun_out
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 60 --timeout-method=thread -x"
timeout 150 $PT tests/test_gpu_stages.py -k "former_module and 27-3" > gpurun_out/v2_sanity_$TAG.log 2>&1; rc=$?
echo "sanity exit $rc"; tail -n 15 gpurun_out/v2_sanity_$TAG.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 600 $PT tests/test_gpu_stages.py -k "former_module" > gpurun_out/v2_modules_$TAG.log 2>&1; echo "modules exit $?"; tail -n 8 gpurun_out/v2_modules_$TAG.log
timeout 600 $PT tests/test_gpu_forward.py > gpurun_out/v2_forward_$TAG.log 2>&1; echo "forward exit $?"; tail -n 8 gpurun_out/v2_forward_$TAG.log
if [ "$2" != "nobench" ]; then
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
  tail -n 3 gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json
  KASF_ONE_TILE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_${TAG}_onetile.json 2>/dev/null; echo "one-tile bench exit $?"; cat gpurun_out/bench_${TAG}_onetile.json
fi
