"""Micro-benchmark of the hot kernels at the canonical-tile layer shapes. GPU only."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hunyuanvideo_efficiency_b200 import _native as N

dev = torch.device("cuda:0")
peaks = {"bf16_tflops": 1652.1, "hbm_gbs": 6536.4}
try:
    peaks.update(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))))
except Exception:
    pass


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def conv_case(Cin, Cout, T, H, W, k=3, stride=(1, 1, 1), variant=0):
    pad = (k - 1, k // 2, k // 2)
    x = N.Vol(1, T, H, W, Cin, torch.bfloat16, dev, pad)
    x.t.normal_()
    w = (torch.randn(k ** 3, Cout, Cin, device=dev) / (k ** 3 * Cin) ** 0.5).bfloat16()
    b = torch.randn(Cout, device=dev)
    To, Ho, Wo = N.conv_out_dims(T, H, W, stride)
    y = N.Vol(1, To, Ho, Wo, Cout, torch.bfloat16, dev)
    ms = timeit(lambda: N.conv3d_tc(x, w, b, k, stride, Cout, out=y, variant=variant))
    fl = 2.0 * Cout * Cin * k ** 3 * To * Ho * Wo
    tf = fl / ms / 1e9
    print(f"conv_tc Cin={Cin:4d} Cout={Cout:4d} T={T:3d} H={H:4d} W={W:4d} k={k} s={stride} v={variant}: {ms:8.3f} ms  {tf:7.1f} TFLOP/s  "
          f"{100 * tf / peaks['bf16_tflops']:5.1f}% of measured bf16 peak", flush=True)


def gn_case(C, T, H, W):
    x = N.Vol(1, T, H, W, C, torch.bfloat16, dev)
    x.t.normal_()
    g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    ms = timeit(lambda: N.groupnorm(x, g, b, 32, 1e-6, True, pad=(2, 1, 1)))
    by = 2.0 * x.t.numel() * 2
    print(f"gn+silu C={C:4d} T={T:3d} H={H:4d} W={W:4d}: {ms:8.3f} ms  {by / ms / 1e6:7.1f} GB/s algorithmic (1R+1W)  "
          f"{100 * by / ms / 1e6 / peaks['hbm_gbs']:5.1f}% of measured HBM peak", flush=True)


if __name__ == "__main__":
    print("peaks:", peaks.get("bf16_tflops"), "TFLOP/s", peaks.get("hbm_gbs"), "GB/s")
    conv_case(128, 128, 9, 256, 256)
    conv_case(256, 256, 9, 128, 128)
    conv_case(512, 512, 9, 64, 64)
    conv_case(512, 512, 17, 32, 32)
    conv_case(256, 128, 9, 256, 256)
    conv_case(128, 128, 9, 256, 256, stride=(1, 2, 2))
    conv_case(512, 512, 1, 1, 17408, k=1)
    conv_case(8, 128, 9, 256, 256)      # encoder conv_in (3 -> 128, channels padded to 8)
    conv_case(128, 8, 9, 256, 256)      # decoder conv_out (128 -> 3, Cout padded to 8)
    conv_case(128, 128, 65, 256, 256)   # the real canonical-tile size (1.1 GB in, 1.1 GB out)
    conv_case(256, 128, 65, 256, 256)
    gn_case(128, 17, 256, 256)
    gn_case(512, 17, 32, 32)
