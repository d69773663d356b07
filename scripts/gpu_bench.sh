#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python bench.py --gpus 1 --steps ${STEPS:-3} --warmup 3 ${EXTRA} > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "exit $?"; tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
