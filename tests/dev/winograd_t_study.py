"""Precision study for the round-2 candidate of DESIGN.md section 5.1: Winograd F(2,3) along T for the stride-1 causal
3x3x3 convs, with the transformed operands rounded to fp16 as a tensor-core kernel would hold them.

For one conv layer with realistic operands (SiLU(GroupNorm(x)) activations, default-init weights) it compares, against the
fp32 convolution:
  direct : x, w rounded to fp16, fp32 accumulation, output rounded to fp16            (what the kernels do today)
  wino   : V = B^T d (fp32 -> fp16), U = G g (fp32 -> fp16), four 3x3 tap-GEMMs with fp32 accumulation,
           y = A^T M in fp32, output rounded to fp16                                  (1.5x fewer MACs)
CPU only (plain torch); prints relative L2 errors.  usage: python tests/dev/winograd_t_study.py
"""
import torch
import torch.nn.functional as F

torch.manual_seed(0)


def study(cin, cout, T, H, W):
    x = torch.randn(1, cin, T, H, W)
    x = F.silu(F.group_norm(x * 2 + 0.3, 32))                       # operand of a resnet conv
    conv = torch.nn.Conv3d(cin, cout, 3)
    w, b = conv.weight.detach(), conv.bias.detach()
    xp = F.pad(x, (1, 1, 1, 1, 2, 0), mode="replicate")            # causal replicate padding (unet_causal_3d_blocks.py:68,74)
    ref = F.conv3d(xp, w, b)
    h = lambda t: t.half().float()
    direct = h(F.conv3d(h(xp), h(w), b))
    # ---- Winograd F(2,3) along T: outputs (2p, 2p+1) from padded frames d0..d3 = xp[2p .. 2p+3]
    Tp = xp.shape[2]
    npair = (T + 1) // 2
    if Tp < 2 * npair + 2:                                         # odd T: one more (unused) frame so the last pair is complete
        xp = torch.cat([xp, xp[:, :, -1:]], 2)
    g0, g1, g2 = w[:, :, 0], w[:, :, 1], w[:, :, 2]                # [Cout][Cin][3][3] per frame tap
    U = [h(g0), h((g0 + g1 + g2) / 2), h((g0 - g1 + g2) / 2), h(g2)]
    out = torch.empty_like(ref)
    for p in range(npair):
        d = [xp[:, :, 2 * p + i] for i in range(4)]
        V = [h(d[0] - d[2]), h(d[1] + d[2]), h(d[2] - d[1]), h(d[1] - d[3])]
        M = [F.conv2d(V[i], U[i]) for i in range(4)]               # fp32 accumulation of fp16 products
        y0 = M[0] + M[1] + M[2] + b[None, :, None, None]
        y1 = M[1] - M[2] - M[3] + b[None, :, None, None]
        out[:, :, 2 * p] = y0
        if 2 * p + 1 < T:
            out[:, :, 2 * p + 1] = y1
    wino = h(out)
    rel = lambda a: (torch.linalg.vector_norm(a - ref) / torch.linalg.vector_norm(ref)).item()
    print(f"Cin={cin:4d} Cout={cout:4d} T={T:2d} {H}x{W}:  direct fp16 {rel(direct):.3e}   winograd-T fp16 {rel(wino):.3e}   ratio {rel(wino) / rel(direct):.2f}")


for shape in [(64, 64, 8, 16, 16), (128, 128, 5, 12, 12), (256, 128, 4, 8, 8)]:
    study(*shape)
