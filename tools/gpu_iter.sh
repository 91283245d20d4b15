#!/bin/bash
# Iteration loop on the GPU box: coarse-pass tests first, then the full gpu suite, smoke and a bench line.
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "coarse" > gpurun_out/pytest_coarse.log 2>&1; echo "pytest coarse exit $?" >> gpurun_out/pytest_coarse.log
tail -n 25 gpurun_out/pytest_coarse.log
if grep -q "exit 0" gpurun_out/pytest_coarse.log; then
  timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
  tail -n 25 gpurun_out/pytest_gpu.log
  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
  tail -n 4 gpurun_out/smoke.log
  timeout 300 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
  tail -n 3 gpurun_out/bench.log
fi
