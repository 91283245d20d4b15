#!/bin/bash
# N-GPU checks: sharded test, then the bench launched the way the driver does it.
mkdir -p gpurun_out
G=${GPUS:-2}
nvidia-smi -L > gpurun_out/gpus.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_gpus" > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi exit $?"; tail -3 gpurun_out/pytest_multi.log
for n in 1 $G; do
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.log 2>&1
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n$n.log 2>&1
  fi
  echo "bench n=$n exit $?"; tail -n 2 gpurun_out/bench_n$n.log | cut -c1-400
done
