#!/bin/bash
# Round-2 final ncu pass (after the two-engine fused site kernel and the few-row decode kernels landed): launch lists of
# one cfg2 forward and one cached decoding step, --set full of the fused site kernel and the decode kernels.
R=${1:-r02b}
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
N="ncu --clock-control none"
timeout 300 $N --profile-from-start off --metrics $M --csv --log-file gpurun_out/${R}_launches.csv python tools/profile_step.py > gpurun_out/${R}_p1.log 2>&1
timeout 300 $N --profile-from-start off --metrics $M --csv --log-file gpurun_out/${R}_decode_launches.csv python tools/profile_decode_step.py > gpurun_out/${R}_p2.log 2>&1
timeout 300 $N --profile-from-start off --set full --import-source on -k regex:attn_site_fused2 -c 3 -o gpurun_out/${R}_site_fused2 -f python tools/profile_step.py > /dev/null 2>&1
timeout 300 $N --profile-from-start off --set full --import-source on -k regex:decode_attn -c 3 -o gpurun_out/${R}_decode_attn -f python tools/profile_decode_step.py > /dev/null 2>&1
timeout 300 $N --profile-from-start off --set full --import-source on -k regex:rows_l -c 5 -o gpurun_out/${R}_rows_linear -f python tools/profile_decode_step.py > /dev/null 2>&1
ls -la gpurun_out/${R}_*
