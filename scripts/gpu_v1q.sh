#!/bin/bash
# quick loop for the default (one-tile) kernel: module parity + forward parity + phase profile + bench.  usage: gpu_v1q.sh <tag>
TAG=${1:-x}; mkdir -p gpurun_out
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 60 --timeout-method=thread -x"
timeout 150 $PT tests/test_gpu_stages.py -k "former_module and 27-3" > gpurun_out/v1_sanity_$TAG.log 2>&1; rc=$?
echo "sanity exit $rc"; tail -n 4 gpurun_out/v1_sanity_$TAG.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 400 $PT tests/test_gpu_stages.py -k "former_module" > gpurun_out/v1_modules_$TAG.log 2>&1; echo "modules exit $?"; tail -n 3 gpurun_out/v1_modules_$TAG.log
timeout 400 $PT tests/test_gpu_forward.py > gpurun_out/v1_forward_$TAG.log 2>&1; echo "forward exit $?"; tail -n 3 gpurun_out/v1_forward_$TAG.log
timeout 200 python scripts/phase_profile.py 1024 27 > gpurun_out/phases_$TAG.log 2>&1; echo "phases exit $?"; tail -n 6 gpurun_out/phases_$TAG.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-extras > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
tail -n 3 gpurun_out/bench_$TAG.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("clips/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["whole_forward_frac"],4), d["roofline"]["per_kind_ms_per_forward"])
print(d.get("sweep"))
PY
