#!/bin/bash
# Runs the GPU test tiers in separate processes (a trapped kernel poisons its CUDA
# context; isolation keeps the other tiers' results) with per-tier timeouts.
# Usage (on the GPU box, via gpurun):  bash tools/gpu_check.sh [pytest -k expr]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, timeout, args...
  local name=$1 t=$2; shift 2
  timeout $t python -m pytest -q -m gpu --no-header -p no:cacheprovider "$@" > gpurun_out/$name.log 2>&1
  echo "== $name exit $?"; tail -n 25 gpurun_out/$name.log
}
run ln_cast 300 tests/test_gpu_kernels.py -k "layernorm or cast or embed or feature or softmax or generator or label or simple_loss"
run linear 600 tests/test_gpu_kernels.py -k "linear"
run attn 600 tests/test_gpu_kernels.py -k "attn"
for f in "$@"; do run extra_$(basename $f .py) 900 $f; done
