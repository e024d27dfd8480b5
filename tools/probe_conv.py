"""Where does the tcgen05 conv lose time?  Runs each kernel variant with the HYVAE_TC_PROBE knobs
(bit 0: no TMA after the ring is primed, bit 1: all loads hit tile 0 (L2-hot), bit 2: no epilogue). GPU only; results of probed runs are garbage by design."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hunyuanvideo_efficiency_b200 import _native as N

dev = torch.device("cuda:0")


def timeit(fn, iters=6, warm=2):
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


VARIANTS = {0: "auto", 1: "1cta MT1", 2: "1cta MTauto", 3: "2cta", 4: "2cta KHT"}


def case(Cin, Cout, T, H, W, variants, gn=0):
    x = N.Vol(1, T, H, W, Cin, torch.float16, dev, (2, 1, 1)); x.t.normal_()
    w = (torch.randn(27, Cout, Cin, device=dev) / (27 * Cin) ** 0.5).half()
    b = torch.randn(Cout, device=dev)
    y = N.Vol(1, T, H, W, Cout, torch.float16, dev)
    fl = 2.0 * Cout * Cin * 27 * T * H * W
    for v in variants:
        row = []
        for probe in (0, 2, 1, 4, 5):
            os.environ["HYVAE_TC_PROBE"] = str(probe)
            ms = timeit(lambda: N.conv3d_tc(x, w, b, 3, (1, 1, 1), Cout, out=y, variant=v, gn_groups=gn))
            row.append(f"p{probe}: {fl / ms / 1e9:7.1f}")
        os.environ["HYVAE_TC_PROBE"] = "0"
        print(f"{Cin:4d}->{Cout:4d} {T}x{H}x{W} gn={gn} {VARIANTS[v]:12s} TFLOP/s  " + "  ".join(row), flush=True)


if __name__ == "__main__":
    case(128, 128, 17, 256, 256, (1, 2, 3, 4))
    case(128, 128, 17, 256, 256, (2, 4), gn=32)
    case(256, 128, 17, 256, 256, (2, 3, 4))
    case(256, 256, 17, 128, 128, (1, 3, 4))
    case(512, 512, 17, 64, 64, (1, 3, 4))
    case(128, 128, 65, 256, 256, (2, 4))
