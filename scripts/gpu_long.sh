#!/bin/bash
# split-path loop: module parity (all T), forward parity, bench at T=81 and T=243 (256 clips).  usage: gpu_long.sh <tag>
TAG=${1:-x}; mkdir -p gpurun_out
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 120 --timeout-method=thread -x"
timeout 600 $PT tests/test_gpu_stages.py -k "former_module" > gpurun_out/long_modules_$TAG.log 2>&1; echo "modules exit $?"; tail -n 5 gpurun_out/long_modules_$TAG.log
timeout 600 $PT tests/test_gpu_forward.py > gpurun_out/long_forward_$TAG.log 2>&1; echo "forward exit $?"; tail -n 3 gpurun_out/long_forward_$TAG.log
for T in 81 243; do
timeout 300 python bench.py --frames $T --batch 256 --steps 5 --warmup 3 --no-cpu --no-extras --no-sweep > gpurun_out/bench_T${T}_$TAG.json 2> gpurun_out/bench_T${T}_$TAG.err; echo "bench T=$T exit $?"
tail -n 2 gpurun_out/bench_T${T}_$TAG.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_T${T}_$TAG.json"))
print("T=$T clips/s", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "frac", round(d["roofline"]["whole_forward_frac"],4), d["roofline"]["per_kind_ms_per_forward"])
PY
done
