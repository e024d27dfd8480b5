"""`hyvideo.vae`-compatible entry points: load_vae(), t-ops injection, AutoencoderKLCausal3D.

Mirrors /root/reference/hyvideo/vae/__init__.py:15-127 (same signature, return tuple, checkpoint
format and error behaviour) on top of the CUDA-backed model in model.py.
"""
import json
import os
from pathlib import Path

import torch

from .model import (AutoencoderKLCausal3D, AutoencoderKLOutput, DecoderCausal3D, DecoderOutput, DecoderOutput2,
                    DiagonalGaussianDistribution, EncoderCausal3D)

# hyvideo/constants.py:19-23,67-73
PRECISION_TO_TYPE = {"fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16}
MODEL_BASE = os.getenv("MODEL_BASE", "./ckpts")
VAE_PATH = {"884-16c-hy": f"{MODEL_BASE}/hunyuan-video-t2v-720p/vae"}


def _apply_t_ops_config_to_vae(vae: AutoencoderKLCausal3D, t_ops_config: dict):
    """hyvideo/vae/__init__.py:15-63: push the JSON's per-block records into the block objects."""
    enc_cfg = t_ops_config.get("encoder", {})
    for block_cfg in enc_cfg.get("down_blocks", []):
        idx = block_cfg["block_index"]
        if 0 <= idx < len(vae.encoder.down_blocks):
            vae.encoder.down_blocks[idx].apply_t_ops_config(block_cfg)
        else:
            print(f"[Warning] down_block index {idx} out of range of encoder.down_blocks.")
    vae.encoder.mid_block.apply_t_ops_config_midblock(enc_cfg.get("mid_block", {}))
    dec_cfg = t_ops_config.get("decoder", {})
    for block_cfg in dec_cfg.get("up_blocks", []):
        idx = block_cfg["block_index"]
        if 0 <= idx < len(vae.decoder.up_blocks):
            vae.decoder.up_blocks[idx].apply_t_ops_config(block_cfg)
        else:
            print(f"[Warning] up_block index {idx} out of range of decoder.up_blocks.")
    vae.decoder.mid_block.apply_t_ops_config_midblock(dec_cfg.get("mid_block", {}))


def load_t_ops_config(json_path: str) -> dict:
    with open(json_path, "r") as f:
        return json.load(f)


def load_vae(vae_type: str = "884-16c-hy", vae_precision: str = None, sample_size: tuple = None, vae_path: str = None,
             logger=None, device=None, t_ops_config_path: str = None, test: bool = False):
    """Load the 3D VAE (config.json + pytorch_model.pt under `vae_path`), exactly like the reference:
    returns (vae, vae_path, spatial_compression_ratio, time_compression_ratio)."""
    if vae_path is None:
        vae_path = VAE_PATH[vae_type]
    if logger is not None:
        logger.info(f"Loading 3D VAE model ({vae_type}) from: {vae_path}")
    config = AutoencoderKLCausal3D.load_config(vae_path)
    vae = AutoencoderKLCausal3D.from_config(config, sample_size=sample_size) if sample_size else AutoencoderKLCausal3D.from_config(config)

    vae_ckpt = Path(vae_path) / "pytorch_model.pt"
    assert vae_ckpt.exists(), f"VAE checkpoint not found: {vae_ckpt}"
    ckpt = torch.load(vae_ckpt, map_location=vae.device, weights_only=False)
    if "state_dict" in ckpt:
        ckpt = ckpt["state_dict"]
    if any(k.startswith("vae.") for k in ckpt.keys()):
        ckpt = {k.replace("vae.", ""): v for k, v in ckpt.items() if k.startswith("vae.")}
    vae.load_state_dict(ckpt)

    spatial_compression_ratio = vae.config.spatial_compression_ratio
    time_compression_ratio = vae.config.time_compression_ratio
    if vae_precision is not None:
        vae = vae.to(dtype=PRECISION_TO_TYPE[vae_precision])
    vae.requires_grad_(False)
    if logger is not None:
        logger.info(f"VAE to dtype: {vae.dtype}")
    if device is not None:
        vae = vae.to(device)
    vae.eval()
    if t_ops_config_path is not None and test:
        if logger is not None:
            logger.info("Applying T-pool/pad configs to the loaded VAE.")
        _apply_t_ops_config_to_vae(vae, load_t_ops_config(t_ops_config_path))
    return vae, vae_path, spatial_compression_ratio, time_compression_ratio
