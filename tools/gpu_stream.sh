#!/bin/bash
# Streaming dot-product top-k: parity tests, then timings (everything under timeout: a hung tcgen05 pipeline must not
# hold the box).
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_streaming.py -m gpu -x -q ${PYTEST_ARGS:-} > gpurun_out/pytest_stream.log 2>&1; echo "pytest stream exit $?" >> gpurun_out/pytest_stream.log
tail -n 30 gpurun_out/pytest_stream.log
if grep -q "stream exit 0" gpurun_out/pytest_stream.log; then
  timeout 300 python tools/bench_streaming.py > gpurun_out/bench_streaming.json 2> gpurun_out/bench_streaming.err; echo "bench exit $?"
  cat gpurun_out/bench_streaming.json
  tail -n 5 gpurun_out/bench_streaming.err
fi
