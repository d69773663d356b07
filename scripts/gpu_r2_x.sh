#!/bin/bash
# round-2 visit X (2 GPUs): where the config-5 D step spends its time over NCCL
mkdir -p gpurun_out
ADVGRPO_TRACE_DSTEP=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config 5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2x_cfg5_n2.json 2> gpurun_out/r2x_cfg5_n2.err; echo "cfg5 exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/r2x_cfg5_n2.json').read().strip().splitlines()[-1]); print(d['value'], d['n_gpus'], d['ms_per_step'], d['phases_ms'])"
