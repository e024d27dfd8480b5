"""CPU oracle for the HunyuanVideo 3D causal VAE hot path.  TEST INFRASTRUCTURE ONLY.

A functional, fp32, plain-PyTorch restatement of the algorithm in /root/reference/hyvideo/vae
(every function cites the reference file:line it follows).  It exists to CHECK the CUDA path:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
it.  The product package (hunyuanvideo_efficiency_b200) never does.

Pinning: oracle/make_golden.py runs the UNMODIFIED reference (imported from /root/reference through
oracle/_refshim) on seeded inputs + the deterministic weights of oracle/weights.py and commits the
results under tests/golden/; tests/test_oracle_golden.py checks this file against those fixtures.
The reference's own tests hold no golden vectors for this path, and the mid-block
attention lives in the third-party `diffusers==0.31.0` (requirements.txt:2), restated below from its
published AttnProcessor2_0 semantics and anchored on the call site unet_causal_3d_blocks.py:580-592,661.

State is a flat dict `sd` with the reference's 248 state-dict key names; `cfg` is the model config
dict (same keys as AutoencoderKLCausal3D.__init__, autoencoder_kl_causal_3d.py:63-82).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
# "explicit": softmax(q k^T * scale + mask) v written out (fp32 checker).  "sdpa": the same through
# F.scaled_dot_product_attention with the dense additive mask, i.e. the call diffusers' AttnProcessor2_0 makes — used by
# bench.py's torch_gpu_baseline leg, where the library kernel the reference would run is what is being timed.
ATTN_IMPL = "explicit"
EPS = 1e-6  # resnet_eps everywhere on this path (vae.py:88,103; unet_causal_3d_blocks.py:266)


# ----------------------------------------------------------------------------- primitive ops
def causal_conv3d(x: Tensor, w: Tensor, b: Optional[Tensor], stride=(1, 1, 1)) -> Tensor:
    """unet_causal_3d_blocks.py:49-75: replicate-pad (k//2 each side of W and H, k-1 in FRONT of T,
    nothing behind), then an un-padded Conv3d."""
    k = w.shape[-1]
    x = F.pad(x, (k // 2, k // 2, k // 2, k // 2, k - 1, 0), mode="replicate")
    return F.conv3d(x, w, b, stride=tuple(stride))


def group_norm(x: Tensor, w: Tensor, b: Tensor, groups: int) -> Tensor:
    """torch.nn.GroupNorm(groups, C, eps=1e-6) as built at unet_causal_3d_blocks.py:302,323."""
    return F.group_norm(x, groups, w, b, EPS)


def t_avg_pool(x: Tensor, k: int, s: int) -> Tensor:
    """unet_causal_3d_blocks.py:665-668,767-772: k-1 replicated frames in front, avg_pool3d over T."""
    x = F.pad(x, (0, 0, 0, 0, k - 1, 0), mode="replicate")
    return F.avg_pool3d(x, kernel_size=(k, 1, 1), stride=(s, 1, 1))


def t_interp(x: Tensor, scale, mode: str) -> Tensor:
    """unet_causal_3d_blocks.py:893-897,906-910: F.interpolate along T only."""
    if x.shape[2] == 0:
        return x
    return F.interpolate(x, scale_factor=(scale, 1, 1), mode=mode)


def upsample_nearest_causal(x: Tensor, factor) -> Tensor:
    """unet_causal_3d_blocks.py:152-171: frame 0 is upsampled in (H, W) only, frames 1.. in (T, H, W)."""
    ft, fh, fw = factor
    first, rest = x[:, :, :1], x[:, :, 1:]
    first = first.repeat_interleave(fh, dim=3).repeat_interleave(fw, dim=4)
    if rest.shape[2] == 0:
        return first
    rest = rest.repeat_interleave(ft, dim=2).repeat_interleave(fh, dim=3).repeat_interleave(fw, dim=4)
    return torch.cat([first, rest], dim=2)


def frame_causal_mask(n_frame: int, n_hw: int, dtype=torch.float32, device=None) -> Tensor:
    """unet_causal_3d_blocks.py:38-46: query i may see every key of frames <= frame(i)."""
    f = torch.arange(n_frame * n_hw, device=device) // n_hw
    m = torch.zeros(n_frame * n_hw, n_frame * n_hw, dtype=dtype, device=device)
    m.masked_fill_(f[None, :] > f[:, None], float("-inf"))
    return m


# ----------------------------------------------------------------------------- blocks
def resnet_block(sd: Dict[str, Tensor], p: str, x: Tensor, groups: int) -> Tensor:
    """ResnetBlockCausal3D.forward, unet_causal_3d_blocks.py:350-417 (temb is None on this path)."""
    h = F.silu(group_norm(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"], groups))
    h = causal_conv3d(h, sd[p + "conv1.conv.weight"], sd[p + "conv1.conv.bias"])
    h = F.silu(group_norm(h, sd[p + "norm2.weight"], sd[p + "norm2.bias"], groups))
    h = causal_conv3d(h, sd[p + "conv2.conv.weight"], sd[p + "conv2.conv.bias"])
    if (p + "conv_shortcut.conv.weight") in sd:  # :338-348, present iff Cin != Cout
        x = causal_conv3d(x, sd[p + "conv_shortcut.conv.weight"], sd[p + "conv_shortcut.conv.bias"])
    return (x + h) / 1.0  # :415, output_scale_factor == 1


def sdpa_frame_causal(q: Tensor, k: Tensor, v: Tensor, n_frame: int, n_hw: int, scale: float) -> Tensor:
    """The F.scaled_dot_product_attention call inside diffusers' AttnProcessor2_0 (call site unet_causal_3d_blocks.py:661)
    with the additive mask of :38-46, one head: softmax(q k^T * scale + mask) v.  q, k, v: [L][D]."""
    s = torch.matmul(q, k.transpose(0, 1)) * scale + frame_causal_mask(n_frame, n_hw, q.dtype, q.device)
    return torch.matmul(torch.softmax(s, dim=-1), v)


def attention_block(sd: Dict[str, Tensor], p: str, x: Tensor, groups: int) -> Tensor:
    """unet_causal_3d_blocks.py:656-662 + diffusers Attention (one head, head_dim = C)."""
    B, C, T, H, W = x.shape
    seq = x.permute(0, 2, 3, 4, 1).reshape(B, T * H * W, C)  # 'b c f h w -> b (f h w) c'
    res = seq
    hn = group_norm(seq.transpose(1, 2), sd[p + "group_norm.weight"], sd[p + "group_norm.bias"], groups).transpose(1, 2)
    q = F.linear(hn, sd[p + "to_q.weight"], sd[p + "to_q.bias"])
    k = F.linear(hn, sd[p + "to_k.weight"], sd[p + "to_k.bias"])
    v = F.linear(hn, sd[p + "to_v.weight"], sd[p + "to_v.bias"])
    mask = frame_causal_mask(T, H * W, x.dtype, x.device)[None]
    if ATTN_IMPL == "sdpa":  # one head: [B, 1, L, C]; default scale = C ** -0.5
        o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None], attn_mask=mask[:, None])[:, 0]
    else:
        s = torch.matmul(q, k.transpose(1, 2)) * (C ** -0.5) + mask
        o = torch.matmul(torch.softmax(s, dim=-1), v)
    o = F.linear(o, sd[p + "to_out.0.weight"], sd[p + "to_out.0.bias"]) + res
    return o.reshape(B, T, H, W, C).permute(0, 4, 1, 2, 3).contiguous()


def _pool_conf(t_ops: Optional[dict], i: int) -> dict:
    """Per-resnet pool record as stored by apply_t_ops_config (unet_causal_3d_blocks.py:636-645,755-762)."""
    if not t_ops:
        return {}
    epb = t_ops.get("enable_t_pool_before_block", [])
    epa = t_ops.get("enable_t_pool_after_block", [])
    if not epb and not epa:
        return {}
    return {"before": epb[i], "after": epa[i], "k": t_ops.get("pool_t_kernel", 2), "s": t_ops.get("pool_t_stride", 2)}


def mid_block(sd, p: str, x: Tensor, groups: int, t_ops: Optional[dict] = None, attention: bool = True) -> Tensor:
    """UNetMidBlockCausal3D.forward, unet_causal_3d_blocks.py:647-678."""
    for i in range(2):
        if i > 0 and attention:
            x = attention_block(sd, p + "attentions.0.", x, groups)
        pc = _pool_conf(t_ops, i)
        if pc.get("before"):
            x = t_avg_pool(x, pc["k"], pc["s"])
        x = resnet_block(sd, f"{p}resnets.{i}.", x, groups)
        if pc.get("after"):
            x = t_avg_pool(x, pc["k"], pc["s"])
    return x


def encoder_strides(cfg) -> List[Optional[tuple]]:
    """vae.py:58-81: per down block the downsample stride, or None when the block has no downsampler."""
    n = len(cfg["block_out_channels"])
    ns = int(math.log2(cfg.get("spatial_compression_ratio", 8)))
    nt = int(math.log2(cfg.get("time_compression_ratio", 4)))
    out = []
    for i in range(n):
        sp = i < ns
        tm = (i >= n - 1 - nt) and i != n - 1
        out.append(((2 if tm else 1), (2 if sp else 1), (2 if sp else 1)) if (sp or tm) else None)
    return out


def encoder_forward(sd, cfg, x: Tensor, t_ops: Optional[dict] = None) -> Tensor:
    """EncoderCausal3D.forward, vae.py:118-136; down blocks unet_causal_3d_blocks.py:764-790."""
    g = cfg.get("norm_num_groups", 32)
    L = cfg.get("layers_per_block", 2)
    enc_ops = (t_ops or {}).get("encoder", {})
    blk_ops = {b["block_index"]: b for b in enc_ops.get("down_blocks", [])}
    x = causal_conv3d(x, sd["encoder.conv_in.conv.weight"], sd["encoder.conv_in.conv.bias"])
    for i, stride in enumerate(encoder_strides(cfg)):
        ops = blk_ops.get(i)
        for j in range(L):
            pc = _pool_conf(ops, j)
            if pc.get("before"):
                x = t_avg_pool(x, pc["k"], pc["s"])
            x = resnet_block(sd, f"encoder.down_blocks.{i}.resnets.{j}.", x, g)
            if pc.get("after"):
                x = t_avg_pool(x, pc["k"], pc["s"])
        if stride is not None:
            if ops and "downsample_stride" in ops:  # :736-742 run-time stride override
                stride = tuple(ops["downsample_stride"])
            q = f"encoder.down_blocks.{i}.downsamplers.0.conv.conv."
            x = causal_conv3d(x, sd[q + "weight"], sd[q + "bias"], stride)
    x = mid_block(sd, "encoder.mid_block.", x, g, enc_ops.get("mid_block") or None, cfg.get("mid_block_add_attention", True))
    x = F.silu(group_norm(x, sd["encoder.conv_norm_out.weight"], sd["encoder.conv_norm_out.bias"], g))
    return causal_conv3d(x, sd["encoder.conv_out.conv.weight"], sd["encoder.conv_out.conv.bias"])


def decoder_upfactors(cfg) -> List[Optional[tuple]]:
    """vae.py:176-201."""
    n = len(cfg["block_out_channels"])
    ns = int(math.log2(cfg.get("spatial_compression_ratio", 8)))
    nt = int(math.log2(cfg.get("time_compression_ratio", 4)))
    out = []
    for i in range(n):
        sp = i < ns
        tm = (i >= n - 1 - nt) and i != n - 1
        out.append(((2 if tm else 1), (2 if sp else 1), (2 if sp else 1)) if (sp or tm) else None)
    return out


def decoder_forward(sd, cfg, z: Tensor, t_ops: Optional[dict] = None) -> Tensor:
    """DecoderCausal3D.forward, vae.py:230-294; up blocks unet_causal_3d_blocks.py:873-917."""
    g = cfg.get("norm_num_groups", 32)
    L = cfg.get("layers_per_block", 2) + 1
    dec_ops = (t_ops or {}).get("decoder", {})
    blk_ops = {b["block_index"]: b for b in dec_ops.get("up_blocks", [])}
    x = causal_conv3d(z, sd["decoder.conv_in.conv.weight"], sd["decoder.conv_in.conv.bias"])
    x = mid_block(sd, "decoder.mid_block.", x, g, dec_ops.get("mid_block") or None, cfg.get("mid_block_add_attention", True))
    for i, fac in enumerate(decoder_upfactors(cfg)):
        ops = blk_ops.get(i) or {}
        eib = ops.get("enable_t_interp_before_block", [False] * L)
        eia = ops.get("enable_t_interp_after_block", [False] * L)
        sc, mode = ops.get("interp_t_scale_factor", 2), ops.get("interp_mode", "nearest")
        for j in range(L):
            if eib[j]:
                x = t_interp(x, sc, mode)
            x = resnet_block(sd, f"decoder.up_blocks.{i}.resnets.{j}.", x, g)
            if eia[j]:
                x = t_interp(x, sc, mode)
        if fac is not None:
            x = upsample_nearest_causal(x, fac)
            q = f"decoder.up_blocks.{i}.upsamplers.0.conv.conv."
            x = causal_conv3d(x, sd[q + "weight"], sd[q + "bias"])
    x = F.silu(group_norm(x, sd["decoder.conv_norm_out.weight"], sd["decoder.conv_norm_out.bias"], g))
    return causal_conv3d(x, sd["decoder.conv_out.conv.weight"], sd["decoder.conv_out.conv.bias"])


# ----------------------------------------------------------------------------- tiling
@dataclass
class Tiling:
    """Mutable tiling attributes of AutoencoderKLCausal3D (autoencoder_kl_causal_3d.py:118-132)."""
    spatial: bool = False
    temporal: bool = False
    sample_min_size: int = 256
    latent_min_size: int = 32
    sample_min_tsize: int = 64
    latent_min_tsize: int = 16
    overlap: float = 0.25

    @staticmethod
    def from_cfg(cfg, spatial=False, temporal=False) -> "Tiling":
        ss = cfg.get("sample_size", 32)
        ss = ss[0] if isinstance(ss, (list, tuple)) else ss
        st = cfg.get("sample_tsize", 64)
        return Tiling(spatial, temporal, ss, int(ss / (2 ** (len(cfg["block_out_channels"]) - 1))),
                      st, st // cfg.get("time_compression_ratio", 4), 0.25)


def blend_v(a: Tensor, b: Tensor, e: int) -> Tensor:
    """autoencoder_kl_causal_3d.py:344-348 (in place on b)."""
    e = min(a.shape[-2], b.shape[-2], e)
    for y in range(e):
        b[:, :, :, y, :] = a[:, :, :, -e + y, :] * (1 - y / e) + b[:, :, :, y, :] * (y / e)
    return b


def blend_h(a: Tensor, b: Tensor, e: int) -> Tensor:
    """autoencoder_kl_causal_3d.py:350-354."""
    e = min(a.shape[-1], b.shape[-1], e)
    for x in range(e):
        b[:, :, :, :, x] = a[:, :, :, :, -e + x] * (1 - x / e) + b[:, :, :, :, x] * (x / e)
    return b


def blend_t(a: Tensor, b: Tensor, e: int) -> Tensor:
    """autoencoder_kl_causal_3d.py:356-360."""
    e = min(a.shape[-3], b.shape[-3], e)
    for x in range(e):
        b[:, :, x, :, :] = a[:, :, -e + x, :, :] * (1 - x / e) + b[:, :, x, :, :] * (x / e)
    return b


def _enc_tile(sd, cfg, x, t_ops):
    h = encoder_forward(sd, cfg, x, t_ops)
    return F.conv3d(h, sd["quant_conv.weight"], sd["quant_conv.bias"])  # :290,391


def _dec_tile(sd, cfg, z, t_ops):
    z = F.conv3d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])  # :307,446
    return decoder_forward(sd, cfg, z, t_ops)


def spatial_assemble(rows, extent: int, limit: int) -> Tensor:
    """Second half of spatial_tiled_encode/decode (:397-412 / :451-465): IN-PLACE raster-order blend chain
    (v from the already blended tile above, then h from the already blended tile on the left), crop, cat."""
    out_rows = []
    for i, row in enumerate(rows):
        out = []
        for j, t in enumerate(row):
            if i > 0:
                t = blend_v(rows[i - 1][j], t, extent)
            if j > 0:
                t = blend_h(row[j - 1], t, extent)
            out.append(t[:, :, :, :limit, :limit])
        out_rows.append(torch.cat(out, dim=-1))
    return torch.cat(out_rows, dim=-2)


def _spatial_tiled(fn, x: Tensor, tile: int, stride: int, extent: int, limit: int) -> Tensor:
    """Shared body of spatial_tiled_encode/decode (:362-420 / :422-469): raster tile grid, then assemble."""
    rows = []
    for i in range(0, x.shape[-2], stride):
        rows.append([fn(x[:, :, :, i:i + tile, j:j + tile]) for j in range(0, x.shape[-1], stride)])
    return spatial_assemble(rows, extent, limit)


def spatial_tiled_encode(sd, cfg, x, tl: Tiling, t_ops=None) -> Tensor:
    """:362-420; returns moments."""
    stride = int(tl.sample_min_size * (1 - tl.overlap))
    extent = int(tl.latent_min_size * tl.overlap)
    return _spatial_tiled(lambda t: _enc_tile(sd, cfg, t, t_ops), x, tl.sample_min_size, stride, extent,
                          tl.latent_min_size - extent)


def spatial_tiled_decode(sd, cfg, z, tl: Tiling, t_ops=None) -> Tensor:
    """:422-469."""
    stride = int(tl.latent_min_size * (1 - tl.overlap))
    extent = int(tl.sample_min_size * tl.overlap)
    return _spatial_tiled(lambda t: _dec_tile(sd, cfg, t, t_ops), z, tl.latent_min_size, stride, extent,
                          tl.sample_min_size - extent)


def temporal_assemble(row, extent: int, limit: int) -> Tensor:
    """Second half of temporal_tiled_encode/decode (:493-501 / :529-537); `row` holds the tiles with the first
    frame of tiles i>0 already dropped (:491,527)."""
    out = []
    for i, t in enumerate(row):
        if i > 0:
            t = blend_t(row[i - 1], t, extent)
            out.append(t[:, :, :limit])
        else:
            out.append(t[:, :, :limit + 1])
    return torch.cat(out, dim=2)


def _temporal_tiled(fn_plain, fn_spatial, x: Tensor, tl: Tiling, tile_t: int, stride: int, extent: int,
                    limit: int, min_size: int) -> Tensor:
    """Shared body of temporal_tiled_encode/decode (:471-508 / :510-541)."""
    row = []
    for i in range(0, x.shape[2], stride):
        t = x[:, :, i:i + tile_t + 1]
        if tl.spatial and (t.shape[-1] > min_size or t.shape[-2] > min_size):
            t = fn_spatial(t)
        else:
            t = fn_plain(t)
        if i > 0:
            t = t[:, :, 1:]
        row.append(t)
    return temporal_assemble(row, extent, limit)


def encode_moments(sd, cfg, x: Tensor, tl: Tiling, t_ops=None) -> Tensor:
    """AutoencoderKLCausal3D.encode dispatch, :259-296; returns the 2*latent_channels moments."""
    assert x.ndim == 5
    if tl.temporal and x.shape[2] > tl.sample_min_tsize:
        return _temporal_tiled(lambda t: _enc_tile(sd, cfg, t, t_ops),
                               lambda t: spatial_tiled_encode(sd, cfg, t, tl, t_ops), x, tl,
                               tl.sample_min_tsize, int(tl.sample_min_tsize * (1 - tl.overlap)),
                               int(tl.latent_min_tsize * tl.overlap),
                               tl.latent_min_tsize - int(tl.latent_min_tsize * tl.overlap), tl.sample_min_size)
    if tl.spatial and (x.shape[-1] > tl.sample_min_size or x.shape[-2] > tl.sample_min_size):
        return spatial_tiled_encode(sd, cfg, x, tl, t_ops)
    return _enc_tile(sd, cfg, x, t_ops)


def decode(sd, cfg, z: Tensor, tl: Tiling, t_ops=None) -> Tensor:
    """AutoencoderKLCausal3D._decode dispatch, :298-313."""
    assert z.ndim == 5
    if tl.temporal and z.shape[2] > tl.latent_min_tsize:
        return _temporal_tiled(lambda t: _dec_tile(sd, cfg, t, t_ops),
                               lambda t: spatial_tiled_decode(sd, cfg, t, tl, t_ops), z, tl,
                               tl.latent_min_tsize, int(tl.latent_min_tsize * (1 - tl.overlap)),
                               int(tl.sample_min_tsize * tl.overlap),
                               tl.sample_min_tsize - int(tl.sample_min_tsize * tl.overlap), tl.latent_min_size)
    if tl.spatial and (z.shape[-1] > tl.latent_min_size or z.shape[-2] > tl.latent_min_size):
        return spatial_tiled_decode(sd, cfg, z, tl, t_ops)
    return _dec_tile(sd, cfg, z, t_ops)


def posterior_mean_logvar(moments: Tensor):
    """DiagonalGaussianDistribution.__init__, vae.py:297-316."""
    mean, logvar = torch.chunk(moments, 2, dim=1)
    return mean, torch.clamp(logvar, -30.0, 20.0)


def forward(sd, cfg, x: Tensor, tl: Tiling, t_ops=None, sample_posterior=False, generator=None):
    """AutoencoderKLCausal3D.forward, :543-578: encode -> mode()/sample() -> decode."""
    mean, logvar = posterior_mean_logvar(encode_moments(sd, cfg, x, tl, t_ops))
    z = mean
    if sample_posterior:
        z = mean + torch.exp(0.5 * logvar) * torch.randn(mean.shape, generator=generator, dtype=mean.dtype).to(mean.device)
    return decode(sd, cfg, z, tl, t_ops), mean, logvar


# ----------------------------------------------------------------------------- metrics
def psnr(ref: Tensor, out: Tensor, data_range: float = 2.0) -> float:
    """evaluation/compute_metrics.py:31-36 restated for tensors in [-1, 1] (range 2 instead of 255)."""
    mse = torch.mean((ref.double() - out.double()) ** 2).item()
    return float("inf") if mse == 0 else 10.0 * math.log10(data_range ** 2 / mse)


def rel_err(ref: Tensor, out: Tensor) -> float:
    """||out - ref||_2 / ||ref||_2, the 'relative' of BASELINE.json's tolerances."""
    return (torch.linalg.vector_norm(out.double() - ref.double()) / torch.linalg.vector_norm(ref.double())).item()
