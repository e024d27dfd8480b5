"""A/B of the first-frame temporal fold on single conv launches (CUDA events, same box, same process)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hunyuanvideo_efficiency_b200 import _native as N

dev = torch.device("cuda:0")


def timeit(fn, iters=10, warm=3):
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


for Cin, Cout, T, H, W in [(128, 128, 17, 256, 256), (128, 128, 65, 256, 256), (256, 256, 33, 128, 128), (512, 512, 17, 64, 64),
                           (512, 512, 17, 32, 32), (512, 512, 9, 32, 32)]:
    x = N.Vol(1, T, H, W, Cin, torch.float16, dev, (2, 1, 1)); x.t.normal_()
    w32 = torch.randn(27, Cout, Cin, device=dev) / (27 * Cin) ** 0.5
    w27 = w32.half().contiguous()
    w45 = torch.cat([w32, w32[0:9] + w32[9:18] + w32[18:27], w32[0:9] + w32[9:18]]).half().contiguous()
    b = torch.randn(Cout, device=dev)
    r = N.Vol(1, T, H, W, Cout, torch.float16, dev); r.t.normal_()
    y = N.Vol(1, T, H, W, Cout, torch.float16, dev)
    res = []
    for rep in range(2):
        t0 = timeit(lambda: N.conv3d_tc(x, w27, b, 3, (1, 1, 1), Cout, residual=r, out=y, gn_groups=32))
        t1 = timeit(lambda: N.conv3d_tc(x, w45, b, 3, (1, 1, 1), Cout, residual=r, out=y, gn_groups=32, variant=N.VARIANT_TFOLD))
        res.append((t0, t1))
    t0, t1 = min(r_[0] for r_ in res), min(r_[1] for r_ in res)
    fl = 2.0 * Cout * Cin * 27 * T * H * W
    print(f"{Cin:4d}->{Cout:4d} {T:3d}x{H}x{W}: plain {t0 * 1e3:8.1f} us ({fl / t0 / 1e9:6.0f} TF/s)  fold {t1 * 1e3:8.1f} us ({fl / t1 / 1e9:6.0f} TF/s alg)  "
          f"{100 * (t1 / t0 - 1):+5.1f} %  (MACs {-100.0 / T:+.1f} %)", flush=True)
