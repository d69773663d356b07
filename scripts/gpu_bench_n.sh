#!/bin/bash
# multi-GPU bench as the driver launches it
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_n${N}.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n${N}.json 2> gpurun_out/bench_n${N}.err
echo "exit $?"; tail -5 gpurun_out/bench_n${N}.err; cat gpurun_out/bench_n${N}.json
