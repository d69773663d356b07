#!/bin/bash
# round-2 visit O (8 GPUs): BASELINE configs 3 / 5 / 4 prompt-sharded over 8 ranks + config 2 strong scaling (J1)
mkdir -p gpurun_out
run() { # name, extra args
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --no-cpu-baseline $2 > gpurun_out/r2o_$1.json 2> gpurun_out/r2o_$1.err
  echo "$1 exit $?"; tail -c 1800 gpurun_out/r2o_$1.json; tail -2 gpurun_out/r2o_$1.err
}
run cfg3_n8 "--config 3 --steps 3 --warmup 3"
run cfg5_n8 "--config 5 --steps 4 --warmup 3"
run cfg2_strong_n8 "--config 2 --scaling strong --steps 4 --warmup 3"
run cfg4_n8 "--config 4 --steps 2 --warmup 3"
