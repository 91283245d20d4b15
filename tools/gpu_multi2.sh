#!/bin/bash
# 2-GPU call: NCCL parity test + bench in both layouts
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_gpus" > gpurun_out/pytest_multi.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_multi.log
for lay in replicate shard; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --parallelism $lay > gpurun_out/bench_n2_$lay.log 2> gpurun_out/bench_n2_$lay.err; echo "bench $lay exit $?"; tail -2 gpurun_out/bench_n2_$lay.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n2_$lay.log").readline())
print("$lay", "q/s %.0f"%d["value"], "step %.2f ms"%d["ms_per_step"], "share %.3f"%d["roofline"]["kernel_share_of_step"], "e2e %.0f"%d["e2e"]["value"], d["multi_gpu_check"], d["config"]["parallelism"])
PY
done
