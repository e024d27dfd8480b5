"""Which kernel's fused GroupNorm sums are closer to an fp64 evaluation?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hunyuanvideo_efficiency_b200 import _native as N
dev = torch.device("cuda:0")
torch.manual_seed(0)
B, Cin, Cout, T, H, W = 2, 128, 128, 3, 40, 24
x = N.Vol(B, T, H, W, Cin, torch.float16, dev, (2, 1, 1)); x.t.normal_()
w = (torch.randn(27, Cout, Cin, device=dev) / (27 * Cin) ** 0.5).half()
b = torch.randn(Cout, device=dev)
r = N.Vol(B, T, H, W, Cout, torch.float16, dev); r.t.normal_()
# fp64 reference of conv + bias + residual from the same fp16 operands
xp = x.t.double()  # [B][T+2][H+2][W+2][C]
ref = torch.zeros(B, T, H, W, Cout, dtype=torch.float64, device=dev)
for kt in range(3):
    for kh in range(3):
        for kw in range(3):
            ref += xp[:, kt:kt + T, kh:kh + H, kw:kw + W, :] @ w[(kt * 3 + kh) * 3 + kw].double().T
ref += b.double() + r.t.double()
g = ref.reshape(B, -1, 32, 4)
s_ref = torch.stack([g.sum((1, 3)), (g * g).sum((1, 3))], -1)
for v in (5, 2, 4, 5, 2):
    y = N.conv3d_tc(x, w, b, 3, (1, 1, 1), Cout, residual=r, variant=v, gn_groups=32)
    d = (y.gn_sums - s_ref).abs()
    print(f"variant {v}: max |sum - ref| = {d[..., 0].max().item():.3e}  max |sumsq - ref| = {d[..., 1].max().item():.3e}  out rel err = {((y.t.double() - ref).norm() / ref.norm()).item():.2e}")
