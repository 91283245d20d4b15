#!/bin/bash
# compute-sanitizer over the round-2 kernels: the streaming dot-product top-k (dot_filter_kernel), the tensor-core index build
# (linear_x3_kernel) and the in-search exclusion path.  Sizes are the tests' own (>= 64k items): expect minutes under the tools.
#   gpurun --timeout 1500 -- 'bash tools/gpu_sanitize_stream.sh'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SEL='131072 or 20001 or exact'   # MIPS K = 128 at 131k items, the 16x16x64 index build, in-search exclusion (exact mode)
for tool in memcheck racecheck; do
  timeout ${SAN_TIMEOUT:-500} compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 86 \
    python -m pytest tests/test_gpu_streaming.py -m gpu -x -q -k "$SEL" > gpurun_out/sanitize_stream_$tool.log 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" gpurun_out/sanitize_stream_$tool.log | tail -5
done
