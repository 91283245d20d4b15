#!/bin/bash
# one full ncu capture of the coarse kernel (1 GPU, small batch so the replay stays short)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mol_coarse_kernel -s 1 -c 1 -f -o gpurun_out/prof_coarse python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch ${PROF_BATCH:-128} > gpurun_out/prof_bench.log 2>&1
echo "full capture exit $?"
