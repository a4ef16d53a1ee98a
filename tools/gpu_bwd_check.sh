#!/bin/bash
# Backward-kernel bring-up on the B200 box: each group in its own process (a trapped kernel poisons the CUDA
# context of its process only), logs merged back through gpurun_out/.
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python -m pytest tests/test_gpu_backward.py -q --timeout 300 -k "$@" > gpurun_out/bwd_$name.log 2>&1; echo "== $name: exit $?"; tail -n 25 gpurun_out/bwd_$name.log; }
run gemm "gemm or linear_dgrad"
run rows "layernorm_bwd or cast_colsum or embed_bwd or log_softmax"
run attn "attn_core_bwd"
run model "training_step"
