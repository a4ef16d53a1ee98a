#!/bin/bash
# First GPU call of round 2: triage of the open batch-64 decode item.
# Usage (via gpurun, ~3 GPU-minutes):  bash tools/triage_open_item.sh
# Everything lands in gpurun_out/triage_*.{log,txt}; each step has its own timeout.
mkdir -p gpurun_out
step() { # name, timeout, command...
  local name=$1 t=$2; shift 2
  echo "== $name"
  timeout "$t" "$@" > "gpurun_out/triage_$name.log" 2>&1
  echo "   exit $?"; tail -n 6 "gpurun_out/triage_$name.log"
}
# 1. the reproducers, as real failures (not xfail): did the coherent-load pass fix it?
step reproducers 240 python -m pytest -q -m gpu --no-header -p no:cacheprovider --runxfail -rA \
  "tests/test_gpu_model.py::test_cfg4_decoder_instances_agree_at_batch_64" \
  "tests/test_gpu_model.py::test_cfg4_batch_64_step_vs_oracle" \
  "tests/test_gpu_kernels.py::test_attn_core_many_items"
# 2. graph instances alone / concurrently / vs the eager decoder, with the localisation lines
step concurrency 120 python tools/concurrency_check.py gpurun_out/triage_concurrency.txt
cat gpurun_out/triage_concurrency.txt
# 3. same check without programmatic dependent launch: if the instances agree here, it is a PDL ordering problem
MTN_B200_PDL=0 step concurrency_nopdl 120 python tools/concurrency_check.py gpurun_out/triage_concurrency_nopdl.txt --decode-only
cat gpurun_out/triage_concurrency_nopdl.txt
# 4. sanitizer passes over one batch-64 decode step in eager mode (slow: bounded)
for tool in memcheck initcheck racecheck; do
  step "sanitizer_$tool" 300 /usr/local/cuda/bin/compute-sanitizer --tool $tool --print-limit 20 \
    python tools/stale_memory_check.py "gpurun_out/triage_stale_$tool.txt" 12
done
