"""Run one conv_tc / gn case a few times (target of `ncu --set full`)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hunyuanvideo_efficiency_b200 import _native as N
dev = torch.device("cuda:0")
what = sys.argv[1] if len(sys.argv) > 1 else "conv256"
if what.startswith("conv"):
    Cin, Cout, T, H, W = {"conv128": (128, 128, 17, 256, 256), "conv256": (256, 256, 17, 128, 128), "conv512": (512, 512, 17, 64, 64)}[what]
    x = N.Vol(1, T, H, W, Cin, torch.float16, dev, (2, 1, 1)); x.t.normal_()
    w = (torch.randn(27, Cout, Cin, device=dev) / (27 * Cin) ** 0.5).half()
    b = torch.randn(Cout, device=dev)
    y = N.Vol(1, T, H, W, Cout, torch.float16, dev)
    for _ in range(4):
        N.conv3d_tc(x, w, b, 3, (1, 1, 1), Cout, out=y, gn_groups=32)
else:
    x = N.Vol(1, 17, 256, 256, 128, torch.float16, dev); x.t.normal_()
    g, b = torch.ones(128, device=dev), torch.zeros(128, device=dev)
    for _ in range(4):
        N.groupnorm(x, g, b, 32, 1e-6, True, pad=(2, 1, 1))
torch.cuda.synchronize()
print("done")
