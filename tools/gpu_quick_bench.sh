#!/bin/bash
# quick kernel timing of the main library: bench without the baseline legs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline --no-secondary --no-gpu-eager ${BENCH_ARGS:-} > gpurun_out/bench_quick.log 2> gpurun_out/bench_quick.err; echo "bench exit $?"; tail -3 gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_quick.log").readline())
print("q/s %.0f"%d["value"], "step %.2f ms"%d["ms_per_step"], "kernel %.2f ms"%d["roofline"]["kernel_ms_per_step"], "frac %.3f"%d["roofline"]["frac"], "e2e %.0f"%d["e2e"]["value"], d["search_stats"])
PY
