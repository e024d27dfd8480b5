"""What bounds the sub-pixel phase convs?  Times UpsampleCausal3D (256 -> 256, low-res 33 x 128 x 128, 8 phase launches) in a
sustained loop with HYVAE_TC_PROBE: bit 0 = no TMA loads once the rings are primed, bit 2 = no epilogue (results are garbage then)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hunyuanvideo_efficiency_b200 import _native as N  # noqa: E402
from hunyuanvideo_efficiency_b200.vae.blocks import UpsampleCausal3D  # noqa: E402

dev = torch.device("cuda:0")
m = UpsampleCausal3D(256, use_conv=True, out_channels=256, upsample_factor=(2, 2, 2)).to(dev)
x = N.Vol(1, 17, 128, 128, 256, torch.float16, dev, (1, 1, 1))
x.t.normal_()
fl = 2.0 * 33 * 256 * 256 * 256 * 256 * 27
for probe in (0, 1, 4, 5):
    os.environ["HYVAE_TC_PROBE"] = str(probe)
    for _ in range(5):
        m.forward_vol(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 40
    for _ in range(n):
        m.forward_vol(x)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    print(f"phases 256->256 lo 17x128x128 probe={probe}: {us:.0f} us per layer = {fl / us / 1e6:.0f} TFLOP/s algorithmic", flush=True)
os.environ["HYVAE_TC_PROBE"] = "0"
