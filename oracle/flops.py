"""Algorithmic conv FLOPs of the VAE path (SURVEY.md §8d definition): sum over every nn.Conv3d the
reference executes, in its own tile decomposition, of 2*Cout*Cin*k^3*(B*To*Ho*Wo).  TEST / BENCH
infrastructure (used to scale the CPU baseline sample and to cross-check the library's own counters)."""
from __future__ import annotations

from .vae_oracle import Tiling, decoder_upfactors, encoder_strides


def _c(cout, cin, k, vox):
    return 2.0 * cout * cin * k ** 3 * vox


def _resnet(ci, co, vox):
    f = _c(co, ci, 3, vox) + _c(co, co, 3, vox)
    return f + (_c(co, ci, 1, vox) if ci != co else 0.0)


def encoder_tile_flops(cfg, B, T, H, W):
    boc, L, lc = cfg["block_out_channels"], cfg.get("layers_per_block", 2), cfg["latent_channels"]
    f = _c(boc[0], cfg["in_channels"], 3, B * T * H * W)
    c = boc[0]
    for i, s in enumerate(encoder_strides(cfg)):
        for j in range(L):
            f += _resnet(c if j == 0 else boc[i], boc[i], B * T * H * W)
        c = boc[i]
        if s is not None:
            T, H, W = (T - 1) // s[0] + 1, (H - 1) // s[1] + 1, (W - 1) // s[2] + 1
            f += _c(c, c, 3, B * T * H * W)
    f += 2 * _resnet(c, c, B * T * H * W) + _c(2 * lc, c, 3, B * T * H * W) + _c(2 * lc, 2 * lc, 1, B * T * H * W)
    return f, (T, H, W)


def decoder_tile_flops(cfg, B, T, H, W):
    boc, L, lc = cfg["block_out_channels"][::-1], cfg.get("layers_per_block", 2) + 1, cfg["latent_channels"]
    f = _c(lc, lc, 1, B * T * H * W) + _c(boc[0], lc, 3, B * T * H * W) + 2 * _resnet(boc[0], boc[0], B * T * H * W)
    c = boc[0]
    for i, u in enumerate(decoder_upfactors(cfg)):
        for j in range(L):
            f += _resnet(c if j == 0 else boc[i], boc[i], B * T * H * W)
        c = boc[i]
        if u is not None:
            T, H, W = (1 + 2 * (T - 1) if u[0] == 2 else T), H * u[1], W * u[2]
            f += _c(c, c, 3, B * T * H * W)
    f += _c(cfg["out_channels"], c, 3, B * T * H * W)
    return f, (T, H, W)


def _tiles(n, tile, stride):
    return [min(tile, n - i) for i in range(0, n, stride)]


def path_flops(cfg, shape, tl: Tiling, direction: str):
    """Conv FLOPs and sub-model call count of encode (shape = video) or decode (shape = latent)."""
    B, _, T, H, W = shape
    enc = direction == "encode"
    fn = encoder_tile_flops if enc else decoder_tile_flops
    min_t, min_s = (tl.sample_min_tsize, tl.sample_min_size) if enc else (tl.latent_min_tsize, tl.latent_min_size)
    if tl.temporal and T > min_t:
        ts = [min(min_t + 1, T - i) for i in range(0, T, int(min_t * (1 - tl.overlap)))]
    else:
        ts = [T]
    total, calls = 0.0, 0
    for t in ts:
        if tl.spatial and (H > min_s or W > min_s):
            s = int(min_s * (1 - tl.overlap))
            for h in _tiles(H, min_s, s):
                for w in _tiles(W, min_s, s):
                    total += fn(cfg, B, t, h, w)[0]
                    calls += 1
        else:
            total += fn(cfg, B, t, H, W)[0]
            calls += 1
    return total, calls
