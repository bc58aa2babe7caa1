#!/bin/bash
# BASELINE.json configs[4]: large-batch sweep (256 .. 65,536 clips, T = 27) on one GPU.  usage: gpu_batch_sweep.sh <tag>
TAG=${1:-x}; mkdir -p gpurun_out; : > gpurun_out/${TAG}_batch_sweep.jsonl
for B in 256 1024 4096 16384 65536; do
  timeout 600 python bench.py --batch $B --steps 3 --warmup 3 --no-cpu --no-extras --no-sweep >> gpurun_out/${TAG}_batch_sweep.jsonl 2> gpurun_out/${TAG}_batch_sweep.err
  echo "B=$B exit $?"
done
python - <<PY
import json
for l in open("gpurun_out/${TAG}_batch_sweep.jsonl"):
    d = json.loads(l)
    print(d["config"]["batch_per_gpu"], round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["frac"], 4), round(d["roofline"]["whole_forward_frac"], 4))
PY
