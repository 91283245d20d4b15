#!/bin/bash
# 8-GPU call (also works with fewer): north-star corpus in both layouts, BASELINE config 4 (10M items sharded), a config-5
# shaped run (16x16x64, items reduced to fit the time budget).  Lines go to gpurun_out/bench_n${N}_*.log.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-8}
run() {  # name, bench args
  name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" > gpurun_out/bench_n${N}_$name.log 2> gpurun_out/bench_n${N}_$name.err
  echo "$name exit $?"; tail -2 gpurun_out/bench_n${N}_$name.err | cut -c1-300
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/bench_n${N}_$name.log") if l.startswith("{")][0]
    print("$name", "q/s %.0f"%d["value"], "step %.2f ms"%d["ms_per_step"], "share %.3f"%d["roofline"]["kernel_share_of_step"], "frac %.3f"%d["roofline"]["frac"], "e2e %.0f"%d["e2e"]["value"], d["multi_gpu_check"], d["config"]["parallelism"], d["fallback_queries"])
except Exception as e: print("$name parse failed", e)
PY
}
run north_replicate --steps 10 --warmup 3
run north_shard --steps 10 --warmup 3 --parallelism shard
run cfg4 --config cfg4 --steps 5 --warmup 3
run cfg5_16m --config cfg5 --items 16000000 --steps 2 --warmup 1
