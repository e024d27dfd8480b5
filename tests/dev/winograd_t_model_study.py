"""End-to-end precision of the Winograd-T candidate: the small-config VAE round trip through the CPU
oracle with every stride-1 3x3x3 conv evaluated (a) with fp16-rounded operands directly, (b) in the F(2,3)-along-T form
with fp16-rounded transformed operands — both against the fp32 evaluation.  CPU only.
usage: python tests/dev/winograd_t_model_study.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from oracle import vae_oracle as O
from oracle import weights as W
from oracle import winograd as WG

cfg = dict(W.SMALL_CONFIG, block_out_channels=[32, 64, 64, 64])
sd = W.make_state_dict(cfg)
x = W.make_video((1, 3, 9, 32, 32))
tl = O.Tiling.from_cfg(cfg)
h = lambda t: t.half().float()
plain = O.causal_conv3d


def run(conv):
    O.causal_conv3d = conv
    try:
        with torch.no_grad():
            dec, mean, _ = O.forward(sd, cfg, x, tl)
        return dec, mean
    finally:
        O.causal_conv3d = plain


def direct16(x_, w, b, stride=(1, 1, 1)):
    return h(plain(h(x_), h(w), b, stride))


def wino16(x_, w, b, stride=(1, 1, 1)):
    if w.shape[-1] == 3 and tuple(stride) == (1, 1, 1) and w.shape[1] >= 8:
        return h(WG.causal_conv3d_winograd_t(x_, w, b, rnd=h))
    return direct16(x_, w, b, stride)


ref_dec, ref_mean = run(plain)
for name, fn in (("direct fp16", direct16), ("winograd-T fp16", wino16)):
    dec, mean = run(fn)
    print(f"{name:16s}: latent rel err {O.rel_err(ref_mean, mean):.3e}   decode rel err {O.rel_err(ref_dec, dec):.3e}   PSNR {O.psnr(ref_dec, dec):.1f} dB")
