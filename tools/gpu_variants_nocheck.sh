#!/bin/bash
# timing-only comparison of library variants (results of ablation builds are numerically meaningless)
mkdir -p gpurun_out
for v in ${VARIANTS:-default}; do
  if [ "$v" = "default" ]; then unset MOL_B200_LIB; else export MOL_B200_LIB=$PWD/rails_b200/lib/libmol_b200_$v.so; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --mode tensor ${BENCH_ARGS:-} > gpurun_out/bench_$v.log 2>&1; echo "$v bench exit $?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$v.log").readline())
print("$v", "q/s %.0f"%d["value"], "step %.2f ms"%d["ms_per_step"], "kernel %.2f ms"%d["roofline"]["kernel_ms_per_step"])
PY
done
