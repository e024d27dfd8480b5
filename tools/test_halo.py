"""Correctness + speed of the halo conv kernel variants against the 1-CTA kernel.  Each case runs in its own process
(a trap in one variant must not hide the others).  usage: test_halo.py [case variant probe]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_case(name, variant, probe):
    import torch
    from hunyuanvideo_efficiency_b200 import _native as N
    dev = torch.device("cuda:0")
    Cin, Cout, T, H, W, res, gn = {"small": (128, 128, 3, 40, 24, True, 32), "rag": (64, 128, 2, 21, 19, False, 32),
                                   "c256": (256, 128, 3, 32, 48, True, 32), "thin": (128, 8, 3, 40, 40, False, 0),
                                   "c8": (16, 128, 3, 40, 40, False, 32), "c64": (64, 64, 3, 32, 32, True, 32),
                                   "big": (128, 128, 17, 256, 256, True, 32), "bignr": (128, 128, 17, 256, 256, False, 32),
                                   "big256": (256, 128, 17, 256, 256, False, 32), "bigthin": (128, 8, 17, 256, 256, False, 0), "bigc8": (16, 128, 17, 256, 256, False, 32), "bigc8ng": (16, 128, 17, 256, 256, False, 0), "thinrag": (128, 8, 2, 21, 35, False, 0)}[name]
    torch.manual_seed(0)
    B = 2 if name in ("small", "rag", "thinrag") else 1
    x = N.Vol(B, T, H, W, Cin, torch.float16, dev, (2, 1, 1)); x.t.normal_()
    w = (torch.randn(27, Cout, Cin, device=dev) / (27 * Cin) ** 0.5).half()
    b = torch.randn(Cout, device=dev)
    r = None
    if res:
        r = N.Vol(B, T, H, W, Cout, torch.float16, dev); r.t.normal_()
    y0 = N.conv3d_tc(x, w, b, 3, (1, 1, 1), Cout, residual=r, variant=2, gn_groups=gn)
    os.environ["HYVAE_TC_PROBE"] = str(probe)
    y1 = N.conv3d_tc(x, w, b, 3, (1, 1, 1), Cout, residual=r, variant=variant, gn_groups=gn)
    torch.cuda.synchronize()
    d = (y1.t.float() - y0.t.float()).abs().max().item()
    msg = f"{name} v{variant} p{probe}: max|diff|={d:.3e} ref max={y0.t.float().abs().max().item():.2f}"
    if gn:
        gd = ((y1.gn_sums - y0.gn_sums).abs() / (y0.gn_sums.abs() + 1e-3)).max().item()
        msg += f" gn rel diff={gd:.2e}"
    if name.startswith("big"):
        def t(v):
            for _ in range(2):
                N.conv3d_tc(x, w, b, 3, (1, 1, 1), Cout, residual=r, variant=v, gn_groups=gn)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(6):
                N.conv3d_tc(x, w, b, 3, (1, 1, 1), Cout, residual=r, variant=v, gn_groups=gn)
            e1.record(); torch.cuda.synchronize()
            return 2.0 * Cout * Cin * 27 * T * H * W * 6 / e0.elapsed_time(e1) / 1e9
        os.environ["HYVAE_TC_PROBE"] = "0"
        t0 = t(2)
        os.environ["HYVAE_TC_PROBE"] = str(probe)
        msg += f"  TFLOP/s: v2={t0:.0f} v{variant}={t(variant):.0f}"
    print(msg, flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_case(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]))
    else:
        cases = [("thin", 0, 0), ("thin", 5, 0), ("thinrag", 0, 0), ("bigthin", 0, 0), ("bigthin", 5, 0)]
        for c in cases:
            p = subprocess.run([sys.executable, __file__, c[0], str(c[1]), str(c[2])], capture_output=True, text=True, timeout=120)
            print((p.stdout.strip() or "(no output)") + ("" if p.returncode == 0 else f"  [rc={p.returncode}] {p.stderr.strip()[-300:]}"), flush=True)
