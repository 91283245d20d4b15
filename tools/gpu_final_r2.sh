#!/bin/bash
# Round-2 evidence run (1 GPU): tests, smoke, bench (both arms), ncu launch list of the bench command, one full capture each of
# the coarse kernel, the streaming dot-filter kernel and the tensor-core index-build kernel, secondary timings.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/gpu.txt
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -n 3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -n 2 gpurun_out/smoke.log
timeout 700 python bench.py > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log; tail -n 2 gpurun_out/bench.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "reference exit $?" >> gpurun_out/bench_reference.log; tail -n 2 gpurun_out/bench_reference.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary --no-gpu-eager > gpurun_out/launches_bench.log 2>&1; echo "launch list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mol_coarse_kernel -s 3 -c 1 -f -o gpurun_out/prof_coarse_b512 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary --no-gpu-eager > gpurun_out/prof_bench.log 2>&1; echo "coarse capture exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dot_filter -s 2 -c 1 -f -o gpurun_out/prof_dotfilter python tools/prof_streaming.py mips > /dev/null 2>&1; echo "dot filter capture exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:linear_x3 -s 0 -c 1 -f -o gpurun_out/prof_linear_x3 python tools/prof_index.py > /dev/null 2>&1; echo "linear_x3 capture exit $?"
timeout 300 python tools/bench_streaming.py > gpurun_out/bench_streaming.json 2>/dev/null; echo "streaming exit $?"
timeout 200 python tools/bench_index.py > gpurun_out/bench_index.json 2>/dev/null; echo "index exit $?"
timeout 200 python tools/time_small.py > gpurun_out/time_small.json 2>/dev/null; echo "small exit $?"
ls -la gpurun_out | head -40
