#!/bin/bash
# Run the GPU parity suite in isolated steps (a hung kernel only loses its own step).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 120 --timeout-method=thread"
run() { name=$1; shift; timeout 600 $PT "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?"; tail -n 25 gpurun_out/$name.log; }
run t1_gemm tests/test_gpu_stages.py -k "gemm_primitive"
run t2_simple tests/test_gpu_stages.py -k "kinematic or fusion or tables"
run t2_metrics tests/test_gpu_metrics.py
run t3_modules tests/test_gpu_stages.py -k "former_module"
run t4_forward tests/test_gpu_forward.py
run t5_io tests/test_gpu_io.py
run t6_precision tests/test_gpu_precision.py
run t7_dropin tests/test_gpu_dropin.py
