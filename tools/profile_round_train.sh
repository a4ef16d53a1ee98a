#!/bin/bash
# Run on the GPU box (gpurun): ncu evidence for the TRAINING step.  Outputs land in gpurun_out/.
#   1. launch list of ONE eager training step: per-launch duration + DRAM bytes (cold cache, serialised)
#   2. --set full captures of the dominant backward kernels, picked by launch index from (1): the slowest
#      weight-gradient GEMM (both operands MN-major, split-K), data-gradient GEMM (MN-major B), attention backward and
#      LayerNorm backward launches of the step
R=${1:-r01_train}
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/${R}_launches.csv python tools/profile_train_step.py > gpurun_out/${R}_p1.log 2>&1
cap() { # tag, kernel-name substring
  id=$(python - "$2" "$R" <<'PY'
import csv, sys
pat, R = sys.argv[1], sys.argv[2]
rows = [l for l in open("gpurun_out/%s_launches.csv" % R) if not l.startswith("==")]
best = None
for x in csv.DictReader(rows):
    if x["Metric Name"] == "gpu__time_duration.sum" and pat in x["Kernel Name"]:
        v = float(x["Metric Value"].replace(",", ""))
        if best is None or v > best[1]:
            best = (int(x["ID"]), v)
print(best[0] if best else -1)
PY
)
  echo "capture $1: launch index $id ($2)"
  [ "$id" -ge 0 ] && ncu --profile-from-start off --set full --clock-control none --import-source on --launch-skip $id --launch-count 1 \
      -o gpurun_out/${R}_$1 -f python tools/profile_train_step.py > /dev/null 2>&1
}
cap gemm_wgrad "gemm_f16_tc_kernel<128, 6, 1, 1, 1, 0>"
cap gemm_wgrad_wide "gemm_f16_tc_kernel<256, 4, 1, 1, 1, 0>"
cap gemm_dgrad "gemm_f16_tc_kernel<256, 4, 1, 0, 1, 0>"
cap attn_bwd "attn_core_bwd_tc_kernel"
cap ln_bwd "layernorm_bwd_kernel<4, 0>"
ls -la gpurun_out/${R}_*.ncu-rep
