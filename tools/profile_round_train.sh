#!/bin/bash
# Run on the GPU box (gpurun): ncu evidence for the TRAINING step.  Outputs land in gpurun_out/.
#   1. launch list of ONE eager training step: per-launch duration + DRAM bytes (cold cache, serialised)
#   2. --set full captures of the dominant backward kernels
set -x
R=${1:-r01_train}
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/${R}_launches.csv python tools/profile_train_step.py > gpurun_out/${R}_p1.log 2>&1
# attention backward: the target self-attention site (Lq = Lk = 256) and a memory site
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_core_bwd -c 2 \
    -o gpurun_out/${R}_attn_bwd -f python tools/profile_train_step.py > /dev/null 2>&1
# weight gradient (both operands MN-major, split-K) and data gradient (MN-major B) GEMMs: template args <.., 1, 1, 0> / <.., 0, 1, 0>
ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:gemm_f16_tc_kernel<.*1, 1, 0>" -s 4 -c 2 \
    -o gpurun_out/${R}_gemm_wgrad -f python tools/profile_train_step.py > /dev/null 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:gemm_f16_tc_kernel<.*0, 1, 0>" -s 4 -c 2 \
    -o gpurun_out/${R}_gemm_dgrad -f python tools/profile_train_step.py > /dev/null 2>&1
ncu --profile-from-start off --set full --clock-control none -k regex:layernorm_bwd -s 10 -c 1 \
    -o gpurun_out/${R}_ln_bwd -f python tools/profile_train_step.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/${R}_launches.csv
