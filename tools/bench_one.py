"""Run one conv / gn case a few times (target of `ncu --set full`).  usage: bench_one.py <case>"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hunyuanvideo_efficiency_b200 import _native as N
dev = torch.device("cuda:0")
what = sys.argv[1] if len(sys.argv) > 1 else "conv256"
CASES = {  # Cin, Cout, T, H, W, residual, variant
    "conv128": (128, 128, 17, 256, 256, True, 0), "conv128_1cta": (128, 128, 17, 256, 256, True, 2),
    "conv256": (256, 256, 17, 128, 128, True, 0), "conv512": (512, 512, 17, 64, 64, True, 0),
    "convin": (8, 128, 17, 256, 256, False, 0), "convout": (128, 8, 17, 256, 256, False, 0),
}
if what in CASES:
    Cin, Cout, T, H, W, res, variant = CASES[what]
    x = N.Vol(1, T, H, W, Cin, torch.float16, dev, (2, 1, 1)); x.t.normal_()
    w = (torch.randn(27, Cout, Cin, device=dev) / (27 * Cin) ** 0.5).half()
    b = torch.randn(Cout, device=dev)
    y = N.Vol(1, T, H, W, Cout, torch.float16, dev)
    r = None
    if res:
        r = N.Vol(1, T, H, W, Cout, torch.float16, dev); r.t.normal_()
    for _ in range(4):
        N.conv3d_tc(x, w, b, 3, (1, 1, 1), Cout, residual=r, out=y, gn_groups=32 if Cout >= 64 else 0, variant=variant)
else:
    x = N.Vol(1, 17, 256, 256, 128, torch.float16, dev); x.t.normal_()
    g, b = torch.ones(128, device=dev), torch.zeros(128, device=dev)
    for _ in range(4):
        N.groupnorm(x, g, b, 32, 1e-6, True, pad=(2, 1, 1))
torch.cuda.synchronize()
print("done")
