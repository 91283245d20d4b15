#!/bin/bash
# ncu launch list + one full capture of the coarse kernel + a plain bench run (1 GPU)
mkdir -p gpurun_out
cp /root/repo/MEASURED_PEAKS.json gpurun_out/ 2>/dev/null
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mol_coarse_kernel -s 1 -c 1 -f -o gpurun_out/prof_coarse python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch 128 > gpurun_out/prof_bench.log 2>&1
echo "full capture exit $?"
tail -3 gpurun_out/bench.log
ls -la gpurun_out/
