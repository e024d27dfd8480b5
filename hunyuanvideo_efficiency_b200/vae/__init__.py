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


def _route_block_records(records, blocks, what: str):
    """Hand every JSON record to the block its `block_index` names (out-of-range indices only warn, like
    hyvideo/vae/__init__.py:27-31,47-51)."""
    for rec in records:
        i = rec["block_index"]
        if 0 <= i < len(blocks):
            blocks[i].apply_t_ops_config(rec)
        else:
            print(f"[Warning] {what} index {i} out of range of {what}s.")


def _apply_t_ops_config_to_vae(vae: AutoencoderKLCausal3D, t_ops_config: dict):
    """hyvideo/vae/__init__.py:15-63: the temporal pool / stride / interp experiment description is pushed into
    the encoder's down blocks, the decoder's up blocks and both mid blocks."""
    enc, dec = t_ops_config.get("encoder", {}), t_ops_config.get("decoder", {})
    _route_block_records(enc.get("down_blocks", []), vae.encoder.down_blocks, "encoder.down_block")
    vae.encoder.mid_block.apply_t_ops_config_midblock(enc.get("mid_block", {}))
    _route_block_records(dec.get("up_blocks", []), vae.decoder.up_blocks, "decoder.up_block")
    vae.decoder.mid_block.apply_t_ops_config_midblock(dec.get("mid_block", {}))


def load_t_ops_config(json_path: str) -> dict:
    with open(json_path, "r") as f:
        return json.load(f)


def _read_state_dict(vae_path: str, map_location):
    """`<vae_path>/pytorch_model.pt`, optionally wrapped in {"state_dict": ...} and / or prefixed "vae."
    (hyvideo/vae/__init__.py:94-101).  A missing file is an AssertionError, as in the reference."""
    f = Path(vae_path) / "pytorch_model.pt"
    assert f.exists(), f"VAE checkpoint not found: {f}"
    sd = torch.load(f, map_location=map_location, weights_only=False)
    sd = sd.get("state_dict", sd)
    if any(k.startswith("vae.") for k in sd):
        sd = {k.replace("vae.", ""): v for k, v in sd.items() if k.startswith("vae.")}
    return sd


def load_vae(vae_type: str = "884-16c-hy", vae_precision: str = None, sample_size: tuple = None, vae_path: str = None,
             logger=None, device=None, t_ops_config_path: str = None, test: bool = False):
    """Same signature, checkpoint format, ordering of side effects and return tuple as the reference's
    load_vae (hyvideo/vae/__init__.py:70-127):
    -> (vae, vae_path, spatial_compression_ratio, time_compression_ratio)."""
    log = logger.info if logger is not None else (lambda *_: None)
    vae_path = vae_path if vae_path is not None else VAE_PATH[vae_type]
    log(f"Loading 3D VAE model ({vae_type}) from: {vae_path}")
    overrides = {"sample_size": sample_size} if sample_size else {}
    vae = AutoencoderKLCausal3D.from_config(AutoencoderKLCausal3D.load_config(vae_path), **overrides)
    vae.load_state_dict(_read_state_dict(vae_path, vae.device))
    ratios = (vae.config.spatial_compression_ratio, vae.config.time_compression_ratio)
    if vae_precision is not None:
        vae = vae.to(dtype=PRECISION_TO_TYPE[vae_precision])
    vae.requires_grad_(False)
    log(f"VAE to dtype: {vae.dtype}")
    if device is not None:
        vae = vae.to(device)
    vae.eval()
    if t_ops_config_path is not None and test:  # the experiment hooks are only armed in test mode (:121)
        log("Applying T-pool/pad configs to the loaded VAE.")
        _apply_t_ops_config_to_vae(vae, load_t_ops_config(t_ops_config_path))
    return (vae, vae_path) + ratios
