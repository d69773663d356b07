"""CUDA-event timing of the tcgen05 TF32 convolution against cuDNN (TF32) at the SD3 VAE decoder shapes (B = 8)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adv_grpo_b200 import ops
torch.backends.cudnn.allow_tf32 = True
dev = "cuda"
shapes = [(8, 64, 64, 512, 512, 3), (8, 128, 128, 512, 512, 3), (8, 256, 256, 512, 512, 3), (8, 256, 256, 512, 256, 3),
          (8, 256, 256, 256, 256, 3), (8, 512, 512, 256, 256, 3), (8, 512, 512, 256, 128, 3), (8, 512, 512, 128, 128, 3),
          (8, 256, 256, 512, 256, 1)]
def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for B, H, W, Cin, Cout, k in shapes:
    x = torch.randn(B, Cin, H, W, device=dev).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Cout, Cin, k, k, device=dev) / (Cin * k * k) ** 0.5).contiguous(memory_format=torch.channels_last)
    wp = ops.pack_conv_weight_tf32(w)
    fl = 2.0 * B * H * W * Cin * Cout * k * k
    t_ours = timeit(lambda: ops.conv2d_nhwc_tf32(x, wp, None, k))
    t_lib = timeit(lambda: torch.nn.functional.conv2d(x, w, None, padding=k // 2))
    print(f"conv {H}x{W} {Cin}->{Cout} k{k}: ours {t_ours:7.3f} ms {fl / t_ours / 1e9:7.1f} TFLOP/s | cuDNN tf32 {t_lib:7.3f} ms {fl / t_lib / 1e9:7.1f} TFLOP/s", flush=True)
    del x, w, wp
