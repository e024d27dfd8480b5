"""Upsample phase convs: time with and without the epilogue (HYVAE_TC_PROBE=4) and without TMA loads (1)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hunyuanvideo_efficiency_b200 import _native as N
from hunyuanvideo_efficiency_b200.vae.blocks import UpsampleCausal3D
dev = torch.device("cuda:0")
for (C, T, H, W) in ((256, 33, 128, 128), (512, 17, 64, 64), (512, 17, 32, 32)):
    m = UpsampleCausal3D(C, use_conv=True, out_channels=C, upsample_factor=(2, 2, 2)).to(dev)
    x = N.Vol(1, T, H, W, C, torch.float16, dev); x.t.normal_()
    exe = 2.0 * C * C * 8 * (2 * T - 1) * 2 * H * 2 * W
    for probe in (0, 4):
        os.environ["HYVAE_TC_PROBE"] = str(probe)
        for _ in range(2):
            m.forward_vol(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            m.forward_vol(x)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"C={C} low {T}x{H}x{W} probe={probe}: {ms:.3f} ms for pad + 8 phases + finalize, executed {exe / ms / 1e9:.0f} TFLOP/s", flush=True)
    os.environ["HYVAE_TC_PROBE"] = "0"
