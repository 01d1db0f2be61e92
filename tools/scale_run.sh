#!/bin/bash
# bench.py at N = 8, 4, 2, 1 GPUs of one box, back to back (the driver's scaling run); lines go to gpurun_out/scale_N.log
set -u
mkdir -p gpurun_out
for n in "$@"; do
  if [ "$n" = 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$n.log 2>&1
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/scale_$n.log 2>&1
  fi
  echo "n=$n exit $?"
  grep -o '"value": [0-9.]*' gpurun_out/scale_$n.log | head -2 | tr '\n' ' '; echo
done
