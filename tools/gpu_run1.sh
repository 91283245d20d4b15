#!/bin/bash
# First GPU session: hardware probes, parity tests, smoke, a short bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import torch; print(torch.cuda.get_device_name(0))" >> gpurun_out/gpu.txt 2>&1
for v in 0 1 2; do timeout 120 python tools/run_probe.py mma $v >> gpurun_out/probe.log 2>&1; echo "mma $v exit $?" >> gpurun_out/probe.log; done
timeout 120 python tools/run_probe.py tma >> gpurun_out/probe.log 2>&1; echo "tma exit $?" >> gpurun_out/probe.log
timeout 120 python tools/run_probe.py mufu >> gpurun_out/probe.log 2>&1; echo "mufu exit $?" >> gpurun_out/probe.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 --batch 64 > gpurun_out/bench_b64.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_b64.log
tail -5 gpurun_out/probe.log gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/bench_b64.log
