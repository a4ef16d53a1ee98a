#!/bin/bash
# Round-2 ncu evidence (run on the GPU box via gpurun).  Everything lands in gpurun_out/.
R=${1:-r02}
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
N="ncu --clock-control none"
# 1. launch lists (cold cache, serialised): cfg2 forward, one cached decoding step, cfg5 forward
timeout 300 $N --profile-from-start off --metrics $M --csv --log-file gpurun_out/${R}_launches.csv python tools/profile_step.py > gpurun_out/${R}_p1.log 2>&1
timeout 300 $N --profile-from-start off --metrics $M --csv --log-file gpurun_out/${R}_decode_launches.csv python tools/profile_decode_step.py > gpurun_out/${R}_p2.log 2>&1
timeout 300 $N --profile-from-start off --metrics $M --csv --log-file gpurun_out/${R}_cfg5_launches.csv python tools/profile_step.py --preset cfg5 > gpurun_out/${R}_p3.log 2>&1
# 2. --set full captures
timeout 300 $N --profile-from-start off --set full --import-source on -k regex:attn_site_fused -c 2 -o gpurun_out/${R}_site_fused -f python tools/profile_step.py > /dev/null 2>&1
timeout 300 $N --profile-from-start off --set full --import-source on -k regex:attn_core_tc -c 3 -o gpurun_out/${R}_attn -f python tools/profile_step.py > /dev/null 2>&1
timeout 300 $N --profile-from-start off --set full -k regex:gemm_f16_tc -s 3 -c 4 -o gpurun_out/${R}_gemm -f python tools/profile_step.py > /dev/null 2>&1
timeout 300 $N --profile-from-start off --set full --import-source on -k regex:decode_attn -c 3 -o gpurun_out/${R}_decode_attn -f python tools/profile_decode_step.py > /dev/null 2>&1
timeout 300 $N --profile-from-start off --set full --import-source on -k regex:rows_l -c 5 -o gpurun_out/${R}_rows_linear -f python tools/profile_decode_step.py > /dev/null 2>&1
timeout 300 $N --profile-from-start off --set full --import-source on -k regex:attn_core_tc -s 20 -c 4 -o gpurun_out/${R}_cfg5_attn -f python tools/profile_step.py --preset cfg5 > /dev/null 2>&1
timeout 300 $N --profile-from-start off --set full -k regex:gemm_f16_tc -s 6 -c 5 -o gpurun_out/${R}_cfg5_gemm -f python tools/profile_step.py --preset cfg5 > /dev/null 2>&1
ls -la gpurun_out/${R}_*.ncu-rep gpurun_out/${R}_*launches.csv
