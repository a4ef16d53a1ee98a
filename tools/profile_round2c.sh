#!/bin/bash
# Round-2, second session: ncu pass after P-through-TMEM (attention core + fused site), the TMA reduce-add epilogues and
# the one-kernel feed-forward sublayer.  Launch list of one cfg2 forward; --set full of the fused site kernel, the
# attention core, the linear kernel (incl. an in-place residual launch) and the fused feed-forward kernel.
R=${1:-r02c}
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
N="ncu --clock-control none"
timeout 300 $N --profile-from-start off --metrics $M --csv --log-file gpurun_out/${R}_launches.csv python tools/profile_step.py > gpurun_out/${R}_p1.log 2>&1
timeout 300 $N --profile-from-start off --set full --import-source on -k regex:attn_site_fused2 -c 3 -o gpurun_out/${R}_site_fused2 -f python tools/profile_step.py > /dev/null 2>&1
timeout 300 $N --profile-from-start off --set full --import-source on -k regex:attn_core_tc -c 4 -o gpurun_out/${R}_attn -f python tools/profile_step.py > /dev/null 2>&1
timeout 300 $N --profile-from-start off --set full -k regex:gemm_f16_tc -s 3 -c 8 -o gpurun_out/${R}_gemm -f python tools/profile_step.py > /dev/null 2>&1
MTN_B200_FFN_FUSED=1 timeout 300 $N --profile-from-start off --set full --import-source on -k regex:ffn_fused -c 1 -o gpurun_out/${R}_ffn_fused -f python tools/profile_step.py > /dev/null 2>&1
timeout 300 $N --set full --import-source on -k regex:attn_core_tc -s 2 -c 1 -o gpurun_out/${R}_attn_ns -f python tools/attn_bench.py --one 0 > /dev/null 2>&1
ls -la gpurun_out/${R}_*
