"""Timeline of the quad-layout attention forward (dispatch variant 19): clock64 stamps of softmax warp 0 and of the
MMA warp for the first 4 CTAs, printed as per-tile deltas (cycles).
softmax stamps: 0 loop top, 1 S ready, 2 S in registers (+s_free), 3 max done, 4 exp done, 5 PV(g-1) done (+rescale), 6 P published
QK warp stamps: 0 loop top for QK(g), 1 s_free(g-1) + Q + K ready, 2 QK(g) issued;  PV warp: 4 want p_full(g), 5 got it, 6 V ready, 7 PV(g) issued"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adv_grpo_b200 import _lib, ops
B, S, H, D = 16, 1229, 24, 64
qkv = torch.randn(B, S, 3, H, D, device="cuda").bfloat16()
trace = torch.zeros(4, 2, 128, 8, dtype=torch.int64, device="cuda")
_lib.load().advgrpo_debug_set_attn_trace(trace.data_ptr())
for _ in range(3):
    ops.attention_fwd(qkv, want_lse=False, variant=19)
torch.cuda.synchronize()
t = trace.cpu()
for cta in range(2):
    sm, mm = t[cta, 0], t[cta, 1]
    t0 = sm[20, 0].item()
    print(f"--- CTA {cta}: tiles 20..49 (10 tiles per work item); all times relative to softmax loop top of tile 20")
    print("tile | softmax: top  S_ready S_inreg max_done exp_done pv_ok  P_pub | QK: top ready issued (unused) | PV: want_pfull got V_ok PV_iss")
    for g in range(20, 50):
        a = [(x.item() - t0) for x in sm[g, :7]]
        b = [(x.item() - t0) for x in mm[g]]
        print(f"{g:4d} | " + " ".join(f"{x:7d}" for x in a) + " | " + " ".join(f"{x:7d}" for x in b))
    per = (sm[100, 0] - sm[20, 0]).item() / 80
    print(f"average cycles per tile (this CTA): {per:.0f}")
