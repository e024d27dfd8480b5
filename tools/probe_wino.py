"""Where does the Winograd-T conv lose time / energy?  Times the kernel in a sustained loop (power-capped state) with the weight
loads (HYVAE_TC_PROBE bit 0) and the plane loads (bit 3) switched off once the rings are primed (results are garbage then)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hunyuanvideo_efficiency_b200 import _native as N  # noqa: E402
from hunyuanvideo_efficiency_b200.vae.blocks import CausalConv3d, _GroupNorm  # noqa: E402

dev = torch.device("cuda:0")
for (Cn, T, H, W) in [(128, 17, 256, 256), (256, 17, 128, 128), (512, 17, 64, 64)]:
    conv = CausalConv3d(Cn, Cn, 3).to(dev)
    conv.emit_gn_groups = 32
    norm = _GroupNorm(32, Cn).to(dev)
    xv = N.Vol(1, T, H, W, Cn, torch.float16, dev)
    xv.t.normal_()
    rv = N.Vol(1, T, H, W, Cn, torch.float16, dev)
    rv.t.normal_()
    pl = norm.forward_vol(xv, True, wino=True)
    fl = 2.0 * T * H * W * Cn * Cn * 27
    for probe in (0, 1, 8, 9):
        os.environ["HYVAE_TC_PROBE"] = str(probe)
        for _ in range(50):
            conv.forward_vol(pl, residual=rv)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 400
        for _ in range(n):
            conv.forward_vol(pl, residual=rv)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / n * 1e3
        print(f"{Cn}->{Cn} T{T} {H}x{W} probe={probe}: {us:.0f} us = {fl / us / 1e6:.0f} TFLOP/s algorithmic", flush=True)
    os.environ["HYVAE_TC_PROBE"] = "0"
