#!/bin/bash
# compute-sanitizer over the small fixtures (SURVEY.md section 5 / VERDICT round 1 #8): memcheck and racecheck of the
# tcgen05 coarse kernel (14 mbarriers, setmaxnreg, TMEM aliasing), the exact kernel and the selects.
#   gpurun --timeout 1500 -- 'bash tools/gpu_sanitize.sh'
# Logs: gpurun_out/sanitize_<tool>.log (summaries are copied to profiles/ by hand).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SEL='top_k_matches_reference and (cfg3 or cfg1 or edge_ragged) or coarse_pass_uneven or test_topk_kernel_ties'
for tool in memcheck racecheck; do
  timeout ${SAN_TIMEOUT:-600} compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 86 \
    python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" gpurun_out/sanitize_$tool.log | tail -5
done
