#!/bin/bash
# round-2 visit: 8-GPU weak-scaling line of the default config at head (the driver's SCALE form)
mkdir -p gpurun_out
N=${N:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_head_cfg2_n$N.json 2> gpurun_out/r2_head_cfg2_n$N.err; echo "exit $?"; tail -2 gpurun_out/r2_head_cfg2_n$N.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_head_cfg2_n$N.json').read().strip().splitlines()[-1]); print(d['value'], d['n_gpus'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
