#!/bin/bash
# One line of the scaling sweep: bench.py on N GPUs of one box (run under `gpurun --gpus N`).
#   tools/scale_sweep.sh N [extra bench.py args...]      e.g.  tools/scale_sweep.sh 8 --model janus-pro-7b --batch 32
N=$1; shift
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --steps 2 --warmup 3 --no-cpu-baseline --no-extra "$@"
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 2 --warmup 3 --no-cpu-baseline --no-extra "$@"
fi
