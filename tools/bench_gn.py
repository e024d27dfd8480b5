"""GroupNorm-apply microbenchmark at the big decoder / encoder tensor shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hunyuanvideo_efficiency_b200 import _native as N
dev = torch.device("cuda:0")
for (C, T, H, W) in ((128, 33, 256, 256), (256, 33, 128, 128), (512, 17, 64, 64), (512, 17, 32, 32)):
    x = N.Vol(1, T, H, W, C, torch.float16, dev); x.t.normal_()
    g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    x.gn_sums = torch.zeros((1, 32, 2), dtype=torch.float64, device=dev); x.gn_sums[:, :, 1] = T * H * W * (C // 32); x.gn_groups = 32
    for _ in range(3):
        N.groupnorm(x, g, b, 32, 1e-6, True, pad=(2, 1, 1))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        N.groupnorm(x, g, b, 32, 1e-6, True, pad=(2, 1, 1))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"gn_apply C={C} {T}x{H}x{W}: {ms:.3f} ms  {2 * x.t.numel() * 2 / ms / 1e9:.2f} TB/s algorithmic (1R+1W)", flush=True)
