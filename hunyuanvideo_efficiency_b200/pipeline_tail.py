"""The VAE tail of the text-to-video pipeline as one call.

`HunyuanVideoPipeline.__call__` (/root/reference/hyvideo/diffusion/pipelines/pipeline_hunyuan_video.py:1046-1092) ends
with: expand 4-D latents, `latents / scaling_factor (+ shift_factor)`, `vae.enable_tiling(); vae.decode(latents)`,
squeeze single-frame outputs, `(image / 2 + 0.5).clamp(0, 1)`, `.cpu().float()`.  Under `torchrun` the reference runs this
replicated on every rank; `decode_latents(..., runner=TileParallelVAE(...))` shards the tiles over the
ranks instead and returns the video on rank 0.
"""
from __future__ import annotations

from typing import Optional

import torch



def decode_latents(vae, latents: torch.Tensor, enable_tiling: bool = True, runner=None, to_cpu: bool = True) -> Optional[torch.Tensor]:
    """latents: (B, C, T, h, w) or (B, C, h, w) in the VAE dtype, as they leave the denoising loop.
    Returns the fp32 video in [0, 1] (on the CPU unless to_cpu=False); None on ranks != 0 when `runner` shards the work."""
    expand_temporal_dim = False
    if latents.ndim == 4:
        latents = latents.unsqueeze(2)
        expand_temporal_dim = True
    elif latents.ndim != 5:
        raise ValueError(f"Only support latents with shape (b, c, h, w) or (b, c, f, h, w), but got {latents.shape}.")
    cfg = vae.config
    shift = cfg.get("shift_factor", None) if hasattr(cfg, "get") else getattr(cfg, "shift_factor", None)
    # latents / scaling_factor (+ shift_factor) is folded into post_quant_conv's packed weights, and (x / 2 + 0.5).clamp(0, 1)
    # + the fp32 cast are the epilogue of the last tile-assembly kernel (AutoencoderKLCausal3D.decode_to_image): the tail
    # adds no pass over the latents or the video
    scale, shift = 1.0 / float(cfg.scaling_factor), float(shift) if shift else 0.0
    with torch.no_grad():
        if enable_tiling:
            vae.enable_tiling()
        image = (runner if runner is not None else vae).decode_to_image(latents, scale, shift)
    if image is None:
        return None
    if expand_temporal_dim or image.shape[2] == 1:
        image = image.squeeze(2)
    return image.cpu() if to_cpu else image
