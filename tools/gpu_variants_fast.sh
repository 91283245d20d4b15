#!/bin/bash
# like gpu_variants.sh with short timeouts (experimental kernels that may hang): VARIANTS="default name ..."
mkdir -p gpurun_out
for v in ${VARIANTS:-default}; do
  if [ "$v" = "default" ]; then unset MOL_B200_LIB; else export MOL_B200_LIB=$PWD/rails_b200/lib/libmol_b200_$v.so; fi
  timeout ${T_TEST:-60} python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "coarse or oracle_seeded" > gpurun_out/pytest_$v.log 2>&1; echo "$v pytest exit $?"
  timeout ${T_BENCH:-60} python bench.py --steps 5 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/bench_$v.log 2>&1; echo "$v bench exit $?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$v.log").readline())
    print("$v", "q/s %.0f"%d["value"], "step %.2f ms"%d["ms_per_step"], "kernel %.2f ms"%d["roofline"]["kernel_ms_per_step"], "e2e %.0f"%d["e2e"]["value"])
except Exception as e:
    print("$v", "no bench line:", e)
PY
done
