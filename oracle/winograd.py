"""CPU restatement (TEST INFRASTRUCTURE) of the Winograd F(2,3)-along-T form of the stride-1 causal 3x3x3 convolution
(DESIGN.md section 5.1): the algebra a future tensor-core kernel has to follow, checked against the plain causal conv
of oracle/vae_oracle.py (unet_causal_3d_blocks.py:68-75) in tests/test_oracle_golden.py.

Output frames (2p, 2p+1) are computed from the four padded frames d_i = xp[2p + i] (xp = x with TWO replicated copies of
frame 0 in front, the reference's causal padding):
    V0 = d0 - d2,  V1 = d1 + d2,  V2 = d2 - d1,  V3 = d1 - d3                       (input transform, per pair)
    U0 = g0,  U1 = (g0 + g1 + g2) / 2,  U2 = (g0 - g1 + g2) / 2,  U3 = g2              (weight transform, g_kt = W[:, :, kt])
    M_i = conv2d_3x3(V_i, U_i)   on the replicate-padded planes                         (4 instead of 6 tap groups)
    y[2p] = M0 + M1 + M2 + bias,   y[2p+1] = M1 - M2 - M3 + bias
"""
from __future__ import annotations

from typing import Callable, List, Optional

import torch
import torch.nn.functional as F
from torch import Tensor


def weight_transform(w: Tensor) -> List[Tensor]:
    """w: [Cout][Cin][3][3][3] -> four [Cout][Cin][3][3] tap sets."""
    g0, g1, g2 = w[:, :, 0], w[:, :, 1], w[:, :, 2]
    return [g0, (g0 + g1 + g2) / 2, (g0 - g1 + g2) / 2, g2]


def input_transform(xp: Tensor, p: int) -> List[Tensor]:
    """xp: [B][C][T+2(+1)][H][W] causally padded (and, for odd T, extended by one frame); planes of output pair p."""
    d = [xp[:, :, 2 * p + i] for i in range(4)]
    return [d[0] - d[2], d[1] + d[2], d[2] - d[1], d[1] - d[3]]


def causal_conv3d_winograd_t(x: Tensor, w: Tensor, b: Optional[Tensor], rnd: Callable[[Tensor], Tensor] = lambda t: t) -> Tensor:
    """Stride-1 causal 3x3x3 conv of x [B][Cin][T][H][W]; `rnd` rounds the transformed operands (e.g. to fp16 and back)."""
    B, _, T, H, W = x.shape
    xp = torch.cat([x[:, :, :1], x[:, :, :1], x], 2)              # causal replicate padding in T
    npair = (T + 1) // 2
    if xp.shape[2] < 2 * npair + 2:                               # odd T: the last pair's d3 is never used by a stored output
        xp = torch.cat([xp, xp[:, :, -1:]], 2)
    xp = F.pad(xp, (1, 1, 1, 1, 0, 0), mode="replicate")          # H / W replicate padding commutes with the transform
    U = [rnd(u) for u in weight_transform(w)]
    y = x.new_empty((B, w.shape[0], T, H, W))
    bias = 0 if b is None else b[None, :, None, None]
    for p in range(npair):
        V = [rnd(v) for v in input_transform(xp, p)]
        M = [F.conv2d(V[i], U[i]) for i in range(4)]
        y[:, :, 2 * p] = M[0] + M[1] + M[2] + bias
        if 2 * p + 1 < T:
            y[:, :, 2 * p + 1] = M[1] - M[2] - M[3] + bias
    return y
