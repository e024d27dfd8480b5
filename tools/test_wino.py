"""Winograd-T conv (hyvae_groupnorm_apply_wino + hyvae_conv3d_causal_wino) against the oracle, case by case, with timing.
Developer tool for GPU sessions (the pytest versions live in tests/test_gpu_parity.py)."""
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hunyuanvideo_efficiency_b200 import _native as N  # noqa: E402
from hunyuanvideo_efficiency_b200.vae.blocks import CausalConv3d, _GroupNorm  # noqa: E402
from oracle import vae_oracle as O  # noqa: E402

dev = torch.device("cuda:0")


def planes_ref(f):
    """f: [B][C][T][H][W] fp32 -> list of planes per oracle/winograd.py with the shifted pairing of conv_wino.cu."""
    T = f.shape[2]
    xp = torch.cat([f[:, :, :1], f[:, :, :1], f], 2)
    out = [f[:, :, 0]]
    for p in range((T - 1) // 2):
        d = [xp[:, :, 2 * p + 1 + i] for i in range(4)]
        out += [d[0] - d[2], d[1] + d[2], d[2] - d[1], d[1] - d[3]]
    if T % 2 == 0:
        d = [xp[:, :, T - 1 + i] for i in range(3)]
        out += [d[0] - d[2], d[1] + d[2], d[2] - d[1]]
    return torch.stack(out, 2)


def case(B, Cin, Cout, T, H, W, res, gn, dtype=torch.float16, seed=0, check_planes=True):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, Cin, T, H, W, generator=g).to(dtype)
    conv = CausalConv3d(Cin, Cout, 3).to(dev)
    norm = _GroupNorm(32, Cin).to(dev)
    with torch.no_grad():
        norm.weight.copy_(1 + 0.1 * torch.randn(Cin, generator=g))
        norm.bias.copy_(0.1 * torch.randn(Cin, generator=g))
    conv.emit_gn_groups = 32 if gn else 0
    w = conv.conv.weight.detach().cpu().to(dtype).float()
    b = conv.conv.bias.detach().cpu().float()
    conv.conv.weight.data = conv.conv.weight.data.to(dtype).float()
    f = F.silu(F.group_norm(x.float(), 32, norm.weight.detach().cpu(), norm.bias.detach().cpu(), 1e-6))
    ref = O.causal_conv3d(f, w, b)
    r = torch.randn(B, Cout, T, H, W, generator=g).to(dtype) if res else None
    if res:
        ref = ref + r.float()
    xv = N.Vol.from_ncthw(x.to(dev))
    pl = norm.forward_vol(xv, True, wino=True)
    torch.cuda.synchronize()
    if check_planes:
        pr = planes_ref(f)                                                   # [B][C][NP][H][W]
        got = pl.t[:, :, 1:-1, 1:-1, :].permute(0, 4, 1, 2, 3).float().cpu()
        e = (got - pr).abs().max().item()
        halo_ok = torch.equal(pl.t[:, :, 0], pl.t[:, :, 1]) and torch.equal(pl.t[:, :, :, -1], pl.t[:, :, :, -2])
        print(f"   planes max abs err {e:.2e} halo replicate {halo_ok}")
    rv = N.Vol.from_ncthw(r.to(dev)) if res else None
    t0 = time.time()
    y = conv.forward_vol(pl, residual=rv)
    torch.cuda.synchronize()
    out = y.to_ncthw().float().cpu()
    err = O.rel_err(ref, out)
    msg = f"B{B} {Cin}->{Cout} T{T} {H}x{W} res={res} gn={gn} {dtype}: rel err {err:.3e} ({time.time() - t0:.3f}s)"
    if gn and y.gn_sums is not None:
        cpg = Cout // 32
        o64 = out.double().reshape(B, 32, cpg, -1)
        s_ref = torch.stack([o64.sum((2, 3)), (o64 * o64).sum((2, 3))], -1)
        ge = ((y.gn_sums.cpu() - s_ref).abs() / (s_ref.abs() + 1.0)).max().item()
        msg += f" gn sums rel {ge:.2e}"
    print(msg, flush=True)
    return err


if __name__ == "__main__":
    cases = [(1, 64, 128, 1, 16, 16, False, False), (1, 64, 128, 3, 16, 16, False, False), (1, 64, 128, 5, 16, 16, False, False),
             (1, 128, 128, 9, 40, 24, True, True), (2, 64, 256, 4, 20, 18, False, True), (1, 256, 512, 3, 16, 24, True, True),
             (1, 64, 128, 2, 9, 21, True, False), (1, 512, 512, 17, 32, 32, True, True)]
    worst = 0.0
    for c in cases:
        worst = max(worst, case(*c))
    print("worst rel err (fp16):", worst)
    if len(sys.argv) > 1 and sys.argv[1] == "perf":
        for (Cin, Cout, T, H, W) in [(128, 128, 17, 256, 256), (256, 256, 17, 128, 128), (512, 512, 17, 32, 32), (128, 128, 65, 256, 256)]:
            conv = CausalConv3d(Cin, Cout, 3).to(dev)
            norm = _GroupNorm(32, Cin).to(dev)
            conv.emit_gn_groups = 32
            xv = N.Vol(1, T, H, W, Cin, torch.float16, dev)
            xv.t.normal_()
            rv = N.Vol(1, T, H, W, Cout, torch.float16, dev)
            rv.t.normal_()
            for wino in (True, False):
                os.environ["HYVAE_WINO"] = "1" if wino else "0"
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                for it in range(3):
                    if it == 1:
                        ev[0].record()
                    pl = norm.forward_vol(xv, True, conv.wants_halo(torch.float16), wino=conv.wants_wino(torch.float16))
                ev[1].record()
                for it in range(6):
                    if it == 1:
                        ev[2].record()
                    y = conv.forward_vol(pl, residual=rv)
                ev[3].record()
                torch.cuda.synchronize()
                fl = 2.0 * T * H * W * Cin * Cout * 27
                tg, tc = ev[0].elapsed_time(ev[1]) / 2, ev[2].elapsed_time(ev[3]) / 5
                print(f"{Cin}->{Cout} T{T} {H}x{W} wino={wino}: gn_apply {tg * 1e3:.0f} us, conv {tc * 1e3:.0f} us = {fl / tc / 1e9:.0f} TFLOP/s algorithmic", flush=True)
