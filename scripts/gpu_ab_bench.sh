#!/bin/bash
# A/B: bench value + att_s phase cycles of the default build and of every variant library under variants/.  usage: gpu_ab_bench.sh <tag>
TAG=${1:-x}; mkdir -p gpurun_out
run() { n=$1; lib=$2
  KASF_LIB=$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-extras --no-sweep > gpurun_out/ab_${TAG}_$n.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_${TAG}_$n.json"))
print("$n", "clips/s", round(d["value"]), d["roofline"]["per_kind_ms_per_forward"])
PY
  KASF_LIB=$lib timeout 200 python scripts/phase_profile.py 1024 27 2>&1 | grep "attention spatial"
}
run base $PWD/kasportsformer_b200/libkasf.so
for so in variants/*.so; do run $(basename $so .so) $PWD/$so; done
run base2 $PWD/kasportsformer_b200/libkasf.so
