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
    "convin": (16, 128, 17, 256, 256, False, 0), "convout": (128, 8, 17, 256, 256, False, 0),
}
if what in CASES:
    Cin, Cout, T, H, W, res, variant = CASES[what]
    x = N.Vol(1, T, H, W, Cin, torch.float16, dev, (2, 1, 1)); x.t.normal_()
    w = (torch.randn(27, Cout, Cin, device=dev) / (27 * Cin) ** 0.5).half()
    if variant == 0 and Cout >= 64:   # as the model path does: 18 folded first-frame tap slices appended (variant bit 8)
        w32 = w.float()
        w = torch.cat([w32, w32[0:9] + w32[9:18] + w32[18:27], w32[0:9] + w32[9:18]]).half().contiguous()
        variant = N.VARIANT_TFOLD
    b = torch.randn(Cout, device=dev)
    y = N.Vol(1, T, H, W, Cout, torch.float16, dev)
    r = None
    if res:
        r = N.Vol(1, T, H, W, Cout, torch.float16, dev); r.t.normal_()
    for _ in range(4):
        N.conv3d_tc(x, w, b, 3, (1, 1, 1), Cout, residual=r, out=y, gn_groups=32 if Cout >= 64 else 0, variant=variant)
elif what.startswith("wino"):   # Winograd-T conv (+ the plane-writing GroupNorm apply that feeds it): wino128 / wino256 / wino512 / winogn128
    from hunyuanvideo_efficiency_b200.vae.blocks import CausalConv3d, _GroupNorm
    Cn, T, H, W = {"wino128": (128, 17, 256, 256), "winogn128": (128, 17, 256, 256), "wino256": (256, 17, 128, 128), "wino512": (512, 17, 32, 32)}[what]
    conv = CausalConv3d(Cn, Cn, 3).to(dev); conv.emit_gn_groups = 32
    norm = _GroupNorm(32, Cn).to(dev)
    x = N.Vol(1, T, H, W, Cn, torch.float16, dev); x.t.normal_()
    N.groupnorm(x, *norm._params(), 32, 1e-6, True)   # statistics once (x.gn_sums stays None: use the stats kernel below only once)
    sums = torch.zeros((1, 32, 2), dtype=torch.float64, device=dev); sums[:, :, 1] = float(T * H * W * (Cn // 32))
    x.gn_sums, x.gn_groups = sums, 32
    r = N.Vol(1, T, H, W, Cn, torch.float16, dev); r.t.normal_()
    for _ in range(4):
        pl = norm.forward_vol(x, True, wino=True)
        conv.forward_vol(pl, residual=r)
elif what == "phase":   # decoder up_block2 upsampler: 256 -> 256, low-res 9 x 128 x 128 -> 17 x 256 x 256
    from hunyuanvideo_efficiency_b200.vae.blocks import UpsampleCausal3D
    m = UpsampleCausal3D(256, use_conv=True, out_channels=256, upsample_factor=(2, 2, 2)).to(dev)
    x = N.Vol(1, 9, 128, 128, 256, torch.float16, dev); x.t.normal_()
    for _ in range(3):
        m.forward_vol(x)
elif what == "conv512s":  # mid-block conv at the canonical tile: 512 -> 512, 17 x 32 x 32
    x = N.Vol(1, 17, 32, 32, 512, torch.float16, dev, (2, 1, 1)); x.t.normal_()
    w = (torch.randn(27, 512, 512, device=dev) / (27 * 512) ** 0.5).half()
    b = torch.randn(512, device=dev)
    r = N.Vol(1, 17, 32, 32, 512, torch.float16, dev); r.t.normal_()
    for _ in range(6):
        N.conv3d_tc(x, w, b, 3, (1, 1, 1), 512, residual=r, gn_groups=32)
elif what == "strided128":   # encoder down_block 0 downsampler: 128 -> 128, stride (1, 2, 2), 17 x 256 x 256 -> 17 x 128 x 128
    x = N.Vol(1, 17, 256, 256, 128, torch.float16, dev, (2, 1, 1)); x.t.normal_()
    w = (torch.randn(27, 128, 128, device=dev) / (27 * 128) ** 0.5).half()
    b = torch.randn(128, device=dev)
    for _ in range(5):
        N.conv3d_tc(x, w, b, 3, (1, 2, 2), 128, gn_groups=32)
elif what == "gemm512":   # attention projection: [17408 x 512] x [512 x 512]^T as a k = 1 launch of the pair kernel
    x = N.Vol(1, 1, 1, 17408, 512, torch.float16, dev); x.t.normal_()
    w = (torch.randn(512, 512, device=dev) / 512 ** 0.5).half()
    b = torch.randn(512, device=dev)
    for _ in range(6):
        N.conv3d_tc(x, w, b, 1, (1, 1, 1), 512)
elif what == "attn":     # fused mid-block attention at the canonical tile: L = 17 x 32 x 32, D = 512
    L, D = 17408, 512
    q = torch.randn(L, D, device=dev).half(); k = torch.randn(L, D, device=dev).half()
    vt = torch.randn(D, L, device=dev).half(); bv = torch.randn(D, device=dev)
    for _ in range(4):
        N.attn_block_causal(q, k, vt, bv, 1024, D ** -0.5)
else:
    x = N.Vol(1, 17, 256, 256, 128, torch.float16, dev); x.t.normal_()
    g, b = torch.ones(128, device=dev), torch.zeros(128, device=dev)
    for _ in range(4):
        N.groupnorm(x, g, b, 32, 1e-6, True, pad=(2, 1, 1))
torch.cuda.synchronize()
print("done")
