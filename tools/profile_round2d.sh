#!/bin/bash
# Round-2, third session: the cluster decoding step (csrc/decode_cluster.cu).  Launch list of one cached greedy step and
# --set full of the kernel; in-kernel timeline.
R=${1:-r02d}
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
N="ncu --clock-control none"
timeout 300 $N --profile-from-start off --metrics $M --csv --log-file gpurun_out/${R}dec_launches.csv python tools/profile_decode_step.py > gpurun_out/${R}_p2.log 2>&1
timeout 300 $N --profile-from-start off --set full --import-source on -k regex:decode_cluster -c 1 -o gpurun_out/${R}dec_decode_cluster -f python tools/profile_decode_step.py > /dev/null 2>&1
timeout 200 python tools/decode_cluster_debug.py --batch 64 --steps 12 --stamps --time > gpurun_out/${R}_decode_cluster_timeline.txt 2>&1
ls -la gpurun_out/${R}*
