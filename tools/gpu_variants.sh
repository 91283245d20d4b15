#!/bin/bash
# bench several library variants back to back: VARIANTS="name1 name2" (lib/libmol_b200_<name>.so); "default" = the main build
mkdir -p gpurun_out
for v in ${VARIANTS:-default}; do
  if [ "$v" = "default" ]; then unset MOL_B200_LIB; else export MOL_B200_LIB=$PWD/rails_b200/lib/libmol_b200_$v.so; fi
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "coarse or oracle_seeded" > gpurun_out/pytest_$v.log 2>&1; echo "$v pytest exit $?"
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/bench_$v.log 2>&1; echo "$v bench exit $?"
  timeout 120 python tools/coarse_err.py 2>/dev/null | tail -1
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$v.log").readline())
print("$v", "q/s %.0f"%d["value"], "step %.2f ms"%d["ms_per_step"], "kernel %.2f ms"%d["roofline"]["kernel_ms_per_step"], "e2e %.0f"%d["e2e"]["value"])
PY
done
