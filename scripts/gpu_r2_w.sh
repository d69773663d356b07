#!/bin/bash
# round-2 visit W (2 GPUs): discriminator-step configs over NCCL with the native D optimizer / gradient sync
mkdir -p gpurun_out
for c in 5 3; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config $c --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2w_cfg${c}_n2.json 2> gpurun_out/r2w_cfg${c}_n2.err; echo "cfg$c exit $?"; tail -c 600 gpurun_out/r2w_cfg${c}_n2.json | head -c 600; echo; python -c "
import json; d=json.loads(open('gpurun_out/r2w_cfg${c}_n2.json').read().strip().splitlines()[-1]); print(d['value'], d['n_gpus'], d['ms_per_step'], d['phases_ms'])"
done
timeout 300 python -m pytest tests/test_rewards_gpu.py -q -m gpu -k pickscore_scorer > gpurun_out/r2w_test.log 2>&1; echo "test exit $?"; tail -2 gpurun_out/r2w_test.log
