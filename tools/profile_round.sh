#!/bin/bash
# Run on the GPU box (gpurun): ncu evidence for the round.  Outputs land in gpurun_out/.
#   1. launch list of ONE bench step: per-launch duration + DRAM bytes (cold cache, serialised)
#   2. --set full captures of the dominant kernels (big/small GEMM, attention, LayerNorm)
set -x
R=${1:-r01}
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/${R}_launches.csv python tools/profile_step.py > gpurun_out/${R}_p1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_f16_tc -s 2 -c 1 -o gpurun_out/${R}_gemm_big -f \
    python tools/gemm_bench.py --one 0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_f16_tc -s 2 -c 1 -o gpurun_out/${R}_gemm_small -f \
    python tools/gemm_bench.py --one 7 > /dev/null 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_core -c 4 \
    -o gpurun_out/${R}_attn -f python tools/profile_step.py > /dev/null 2>&1
ncu --profile-from-start off --set full --clock-control none -k regex:layernorm_rows -s 10 -c 2 \
    -o gpurun_out/${R}_ln -f python tools/profile_step.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/${R}_launches.csv
