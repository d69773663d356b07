#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tensorcore_gpu.py -q -m gpu --timeout 600 -x -k "tn_skinny" > gpurun_out/r2q_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/r2q_tests.log
timeout 300 python - > gpurun_out/r2q_tn_perf.log 2>&1 <<'PY'
import torch, sys
sys.path.insert(0, ".")
from adv_grpo_b200 import ops
def t(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/it*1e3
for (Kt,Ms,Nb) in [(16384,128,1536),(16384,64,1536),(16384,128,4608),(16384,64,6144),(3280,128,1536),(3280,64,6144)]:
    a=torch.randn(Kt,Ms,device="cuda").bfloat16(); b=torch.randn(Kt,Nb,device="cuda").bfloat16()
    us=t(lambda: ops.gemm_tn_skinny(a,b)); us2=t(lambda: a.t()@b)
    print(f"tn skinny Kt={Kt} Ms={Ms} Nb={Nb}: ours {us:6.1f} us ({(Kt*Nb*2+Kt*Ms*2)/us/1e3:6.0f} GB/s of B+A)   torch a.t()@b {us2:6.1f} us")
PY
cat gpurun_out/r2q_tn_perf.log
