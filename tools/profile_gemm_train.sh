#!/bin/bash
# ncu --set full captures of the backward GEMM forms of one training step (kernel names matched with their template
# arguments: <BN, STAGES, CL, A_MN, B_MN, DROP>).
R=${1:-r01_train}
ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:gemm_f16_tc_kernel<128, 6, 1, 1, 1, 0>" -s 6 -c 1 -o gpurun_out/${R}_gemm_wgrad -f python tools/profile_train_step.py > /dev/null 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:gemm_f16_tc_kernel<256, 4, 1, 1, 1, 0>" -s 2 -c 1 -o gpurun_out/${R}_gemm_wgrad_wide -f python tools/profile_train_step.py > /dev/null 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:gemm_f16_tc_kernel<256, 4, 1, 0, 1, 0>" -s 6 -c 1 -o gpurun_out/${R}_gemm_dgrad -f python tools/profile_train_step.py > /dev/null 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/${R}_launches.csv python tools/profile_train_step.py > gpurun_out/${R}_p1.log 2>&1
ls -la gpurun_out/${R}_gemm*.ncu-rep
