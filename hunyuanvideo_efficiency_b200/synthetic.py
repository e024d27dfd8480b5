"""Deterministic synthetic weights and inputs for the VAE, keyed by the reference's state-dict names.

There are no checkpoints or datasets offline, so bench.py, smoke(), the tests and oracle/make_golden.py all
draw parameters and clips from here (this module only GENERATES tensors; it contains no VAE arithmetic).  The weights are
NOT the reference's nn.Module default initialisation (that depends on module construction order);
they are a pure function of (config, seed, key name), so the reference (through load_state_dict),
the oracle and the CUDA path can all be given bit-identical parameters anywhere, with no checkpoint.
Distribution: conv / linear weights and biases ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (the bound
PyTorch's default initialisers use), GroupNorm weight = 1 + 0.1 N(0,1), GroupNorm bias = 0.1 N(0,1).
"""
from __future__ import annotations

import hashlib
import math
from collections import OrderedDict
from typing import Dict, Tuple

import torch

HY_VAE_CONFIG = dict(  # ckpts/hunyuan-video-t2v-720p/vae/config.json values
    in_channels=3, out_channels=3, latent_channels=16,
    down_block_types=["DownEncoderBlockCausal3D"] * 4, up_block_types=["UpDecoderBlockCausal3D"] * 4,
    block_out_channels=[128, 256, 512, 512], layers_per_block=2, act_fn="silu", norm_num_groups=32,
    sample_size=256, sample_tsize=64, scaling_factor=0.476986, spatial_compression_ratio=8,
    time_compression_ratio=4, mid_block_add_attention=True,
)

SMALL_CONFIG = dict(HY_VAE_CONFIG, block_out_channels=[32, 64, 128, 128], sample_size=32, sample_tsize=16)


def _strides(cfg):
    n = len(cfg["block_out_channels"])
    ns = int(math.log2(cfg.get("spatial_compression_ratio", 8)))
    nt = int(math.log2(cfg.get("time_compression_ratio", 4)))
    return [(i < ns) or ((i >= n - 1 - nt) and i != n - 1) for i in range(n)]


def state_dict_spec(cfg) -> "OrderedDict[str, Tuple[int, ...]]":
    """Key -> shape for every parameter of AutoencoderKLCausal3D(cfg) (248 keys for the HY config)."""
    spec: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    boc, L, lc = list(cfg["block_out_channels"]), cfg.get("layers_per_block", 2), cfg["latent_channels"]

    def conv(p, ci, co, k):
        spec[p + "conv.weight"] = (co, ci, k, k, k)
        spec[p + "conv.bias"] = (co,)

    def norm(p, c):
        spec[p + "weight"] = (c,)
        spec[p + "bias"] = (c,)

    def resnet(p, ci, co):
        norm(p + "norm1.", ci)
        conv(p + "conv1.", ci, co, 3)
        norm(p + "norm2.", co)
        conv(p + "conv2.", co, co, 3)
        if ci != co:
            conv(p + "conv_shortcut.", ci, co, 1)

    def mid(p, c):
        if cfg.get("mid_block_add_attention", True):
            norm(p + "attentions.0.group_norm.", c)
            for n in ("to_q", "to_k", "to_v", "to_out.0"):
                spec[f"{p}attentions.0.{n}.weight"] = (c, c)
                spec[f"{p}attentions.0.{n}.bias"] = (c,)
        resnet(p + "resnets.0.", c, c)
        resnet(p + "resnets.1.", c, c)

    has_sampler = _strides(cfg)
    conv("encoder.conv_in.", cfg["in_channels"], boc[0], 3)
    co = boc[0]
    for i in range(len(boc)):
        ci, co = co, boc[i]
        for j in range(L):
            resnet(f"encoder.down_blocks.{i}.resnets.{j}.", ci if j == 0 else co, co)
        if has_sampler[i]:
            conv(f"encoder.down_blocks.{i}.downsamplers.0.conv.", co, co, 3)
    mid("encoder.mid_block.", boc[-1])
    norm("encoder.conv_norm_out.", boc[-1])
    conv("encoder.conv_out.", boc[-1], 2 * lc, 3)

    rev = boc[::-1]
    conv("decoder.conv_in.", lc, rev[0], 3)
    mid("decoder.mid_block.", rev[0])
    co = rev[0]
    for i in range(len(rev)):
        ci, co = co, rev[i]
        for j in range(L + 1):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}.", ci if j == 0 else co, co)
        if has_sampler[i]:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv.", co, co, 3)
    norm("decoder.conv_norm_out.", rev[-1])
    conv("decoder.conv_out.", rev[-1], cfg["out_channels"], 3)

    spec["quant_conv.weight"] = (2 * lc, 2 * lc, 1, 1, 1)
    spec["quant_conv.bias"] = (2 * lc,)
    spec["post_quant_conv.weight"] = (lc, lc, 1, 1, 1)
    spec["post_quant_conv.bias"] = (lc,)
    return spec


def _gen(seed: int, key: str) -> torch.Generator:
    h = int.from_bytes(hashlib.sha256(f"{seed}:{key}".encode()).digest()[:7], "little")
    g = torch.Generator(device="cpu")
    g.manual_seed(h)
    return g


def make_state_dict(cfg, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    sd: Dict[str, torch.Tensor] = OrderedDict()
    spec = state_dict_spec(cfg)
    for key, shape in spec.items():
        g = _gen(seed, key)
        is_norm = ("norm" in key.split(".")[-2]) if len(key.split(".")) > 1 else False
        if is_norm:
            t = torch.randn(shape, generator=g) * 0.1
            if key.endswith("weight"):
                t = t + 1.0
        else:
            wkey = key[: -len("bias")] + "weight" if key.endswith("bias") else key
            wshape = spec[wkey]
            fan_in = 1
            for d in wshape[1:]:
                fan_in *= d
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        sd[key] = t.to(dtype)
    return sd


def make_video(shape, seed: int = 1234) -> torch.Tensor:
    """Synthetic clip in the dataset's value range [-1, 1] (dataset_processor/mp42tensor.py:78)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.rand(shape, generator=g) * 2 - 1


def make_latent(shape, seed: int = 4321) -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.randn(shape, generator=g)
