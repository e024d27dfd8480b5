"""AutoencoderKLCausal3D on libhyvae.so: the reference's module API over a CUDA tile schedule.

Surface kept from /root/reference/hyvideo/vae/autoencoder_kl_causal_3d.py and vae.py:
constructor/config keys, `encode / decode / forward`, the tiling and slicing switches and public tiling
attributes, `spatial_tiled_* / temporal_tiled_*`, `blend_v/h/t`, `DiagonalGaussianDistribution`,
`DecoderOutput`, the sub-module tree and its 248 state-dict keys.  `tiled_decode()` (named by the
north-star, absent from the reference) is provided as decode-with-tiling.

Underneath, a sub-model call is: NCTHW slice -> channels-last volume -> kernel schedule (blocks.py) ->
NCTHW tile; tile assembly is the fused blend+crop+scatter kernel applied in the reference's raster
order (the blend chain is in place and order dependent, autoencoder_kl_causal_3d.py:403-409,457-462).
"""
from __future__ import annotations

import inspect
import json
import math
import os
from typing import Optional, Tuple, Union

import numpy as np
import torch
from torch import nn

from .. import _native as N
from .._native import Vol
from .blocks import CausalConv3d, UNetMidBlockCausal3D, _GroupNorm, _publish, get_down_block3d, get_up_block3d


# --------------------------------------------------------------------------------------- outputs
class _Output(dict):
    """Attribute + index access like diffusers' BaseOutput (`.sample`, `["sample"]`, `[0]`)."""

    def __init__(self, **kw):
        super().__init__({k: v for k, v in kw.items() if v is not None})
        self.__dict__.update(kw)

    def __getitem__(self, k):
        return super().__getitem__(k) if isinstance(k, str) else self.to_tuple()[k]

    def to_tuple(self):
        return tuple(self.values())


class DecoderOutput(_Output):
    def __init__(self, sample):
        super().__init__(sample=sample)


class DecoderOutput2(_Output):
    def __init__(self, sample, posterior=None):
        super().__init__(sample=sample, posterior=posterior)


class AutoencoderKLOutput(_Output):
    def __init__(self, latent_dist, tiles_ci=None):  # `tiles_ci` is passed at autoencoder_kl_causal_3d.py:296
        super().__init__(latent_dist=latent_dist)
        self.tiles_ci = tiles_ci


class DiagonalGaussianDistribution(object):
    """vae.py:297-358.  A handful of tiny elementwise torch ops on the moments."""

    def __init__(self, parameters: torch.Tensor, deterministic: bool = False):
        if parameters.ndim == 3:
            dim = 2
        elif parameters.ndim in (4, 5):
            dim = 1
        else:
            raise NotImplementedError
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=dim)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.deterministic = deterministic
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)
        if deterministic:
            self.var = self.std = torch.zeros_like(self.mean)

    def sample(self, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        dev = self.parameters.device
        gdev = generator.device if generator is not None else dev
        noise = torch.randn(self.mean.shape, generator=generator, device=gdev, dtype=self.parameters.dtype).to(dev)
        return self.mean + self.std * noise

    def kl(self, other: "DiagonalGaussianDistribution" = None) -> torch.Tensor:
        if self.deterministic:
            return torch.Tensor([0.0])
        dims = list(range(1, self.mean.ndim))
        if other is None:
            return 0.5 * torch.sum(torch.pow(self.mean, 2) + self.var - 1.0 - self.logvar, dim=dims)
        return 0.5 * torch.sum(torch.pow(self.mean - other.mean, 2) / other.var + self.var / other.var - 1.0
                               - self.logvar + other.logvar, dim=dims)

    def nll(self, sample: torch.Tensor, dims=(1, 2, 3)) -> torch.Tensor:
        if self.deterministic:
            return torch.Tensor([0.0])
        return 0.5 * torch.sum(np.log(2.0 * np.pi) + self.logvar + torch.pow(sample - self.mean, 2) / self.var, dim=list(dims))

    def mode(self) -> torch.Tensor:
        return self.mean


# --------------------------------------------------------------------------------------- sub-models
def _sampler_plan(n_blocks: int, spatial_ratio: int, time_ratio: int):
    """Per block (has_sampler, (t, h, w) factor) — vae.py:58-81 (encoder) and :176-201 (decoder)."""
    if time_ratio != 4:
        raise ValueError(f"Unsupported time_compression_ratio: {time_ratio}.")
    ns, nt = int(np.log2(spatial_ratio)), int(np.log2(time_ratio))
    plan = []
    for i in range(n_blocks):
        sp = i < ns
        tm = (i >= n_blocks - 1 - nt) and i != n_blocks - 1
        plan.append((sp or tm, ((2 if tm else 1), (2 if sp else 1), (2 if sp else 1))))
    return plan


class EncoderCausal3D(nn.Module):
    """vae.py:32-136."""

    def __init__(self, in_channels=3, out_channels=3, down_block_types=("DownEncoderBlockCausal3D",),
                 block_out_channels=(64,), layers_per_block=2, norm_num_groups=32, act_fn="silu", double_z=True,
                 mid_block_add_attention=True, time_compression_ratio=4, spatial_compression_ratio=8):
        super().__init__()
        self.layers_per_block = layers_per_block
        self.conv_in = CausalConv3d(in_channels, block_out_channels[0], kernel_size=3, stride=1)
        self.conv_in.emit_gn_groups = norm_num_groups
        self.down_blocks = nn.ModuleList([])
        plan = _sampler_plan(len(block_out_channels), spatial_compression_ratio, time_compression_ratio)
        oc = block_out_channels[0]
        for i, t in enumerate(down_block_types):
            ic, oc = oc, block_out_channels[i]
            self.down_blocks.append(get_down_block3d(
                t, num_layers=layers_per_block, in_channels=ic, out_channels=oc, add_downsample=plan[i][0],
                downsample_stride=plan[i][1], resnet_eps=1e-6, downsample_padding=0, resnet_act_fn=act_fn,
                resnet_groups=norm_num_groups, attention_head_dim=oc, temb_channels=None))
        self.mid_block = UNetMidBlockCausal3D(
            in_channels=block_out_channels[-1], resnet_eps=1e-6, resnet_act_fn=act_fn, output_scale_factor=1,
            resnet_time_scale_shift="default", attention_head_dim=block_out_channels[-1], resnet_groups=norm_num_groups,
            temb_channels=None, add_attention=mid_block_add_attention)
        self.conv_norm_out = _GroupNorm(norm_num_groups, block_out_channels[-1], 1e-6)
        self.conv_out = CausalConv3d(block_out_channels[-1], 2 * out_channels if double_z else out_channels, kernel_size=3)

    def forward_vol(self, x: Vol) -> Vol:
        x = self.conv_in.forward_vol(x)
        for blk in self.down_blocks:
            x = blk.forward_vol(x)
        x = self.mid_block.forward_vol(x)
        x = self.conv_norm_out.forward_vol(x, True, self.conv_out.wants_halo(x.dtype))
        return self.conv_out.forward_vol(x)

    def forward(self, sample: torch.Tensor) -> torch.Tensor:
        assert len(sample.shape) == 5, "The input tensor should have 5 dimensions"
        return self.forward_vol(Vol.from_ncthw(sample)).to_ncthw()


class DecoderCausal3D(nn.Module):
    """vae.py:139-294."""

    def __init__(self, in_channels=3, out_channels=3, up_block_types=("UpDecoderBlockCausal3D",),
                 block_out_channels=(64,), layers_per_block=2, norm_num_groups=32, act_fn="silu", norm_type="group",
                 mid_block_add_attention=True, time_compression_ratio=4, spatial_compression_ratio=8):
        super().__init__()
        if norm_type != "group":
            raise NotImplementedError("norm_type 'spatial'")
        self.layers_per_block = layers_per_block
        self.conv_in = CausalConv3d(in_channels, block_out_channels[-1], kernel_size=3, stride=1)
        self.conv_in.emit_gn_groups = norm_num_groups
        self.mid_block = UNetMidBlockCausal3D(
            in_channels=block_out_channels[-1], resnet_eps=1e-6, resnet_act_fn=act_fn, output_scale_factor=1,
            resnet_time_scale_shift="default", attention_head_dim=block_out_channels[-1], resnet_groups=norm_num_groups,
            temb_channels=None, add_attention=mid_block_add_attention)
        self.up_blocks = nn.ModuleList([])
        rev = list(reversed(block_out_channels))
        plan = _sampler_plan(len(block_out_channels), spatial_compression_ratio, time_compression_ratio)
        oc = rev[0]
        for i, t in enumerate(up_block_types):
            pc, oc = oc, rev[i]
            self.up_blocks.append(get_up_block3d(
                t, num_layers=layers_per_block + 1, in_channels=pc, out_channels=oc, prev_output_channel=None,
                add_upsample=plan[i][0], upsample_scale_factor=plan[i][1], resnet_eps=1e-6, resnet_act_fn=act_fn,
                resnet_groups=norm_num_groups, attention_head_dim=oc, temb_channels=None,
                resnet_time_scale_shift="default"))
        self.conv_norm_out = _GroupNorm(norm_num_groups, block_out_channels[0], 1e-6)
        self.conv_out = CausalConv3d(block_out_channels[0], out_channels, kernel_size=3)
        self.gradient_checkpointing = False

    def forward_vol(self, x: Vol) -> Vol:
        x = self.conv_in.forward_vol(x)
        x = self.mid_block.forward_vol(x)
        for blk in self.up_blocks:
            x = blk.forward_vol(x)
        x = self.conv_norm_out.forward_vol(x, True, self.conv_out.wants_halo(x.dtype))
        return self.conv_out.forward_vol(x)

    def forward(self, sample: torch.Tensor, latent_embeds=None) -> torch.Tensor:
        assert len(sample.shape) == 5, "The input tensor should have 5 dimensions."
        if latent_embeds is not None:
            raise NotImplementedError("latent_embeds")
        return self.forward_vol(Vol.from_ncthw(sample)).to_ncthw()


# --------------------------------------------------------------------------------------- config
class FrozenConfig(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class _ConvParamsOnly(nn.Module):
    """quant_conv / post_quant_conv: an nn.Conv3d(k=1) in the reference (:114-115); here weights + a
    1x1x1 launch of the conv kernel."""

    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(c, c, 1, 1, 1))
        self.bias = nn.Parameter(torch.empty(c))
        b = 1.0 / math.sqrt(c)
        nn.init.uniform_(self.weight, -b, b)
        nn.init.uniform_(self.bias, -b, b)
        self._packed = None

    def forward_vol(self, x: Vol, affine=None) -> Vol:
        """affine = (scale, shift): computes conv(x * scale + shift) with the affine map folded into the packed weights and
        bias (W' = W * scale, b' = b + shift * W.sum(in)), in fp32, rounded once — the pipeline's `latents / scaling_factor
        (+ shift_factor)` (pipeline_hunyuan_video.py:1060-1069) then costs no pass over the latents."""
        co, ci = self.weight.shape[0], self.weight.shape[1]
        if x.C < ci or x.c_valid > ci:
            raise N.HyvaeError(f"1x1x1 conv: input has {x.c_valid} channels (stored {x.C}), the weight expects {ci}")
        key = (self.weight._version, self.bias._version, self.weight.data_ptr(), x.dtype, x.C, affine)
        if self._packed is None or self._packed[0] != key:
            w32, b32 = self.weight.detach().float().reshape(co, ci), self.bias.detach().float()
            if affine is not None:
                scale, shift = float(affine[0]), float(affine[1])
                b32 = b32 + shift * w32.sum(1)
                w32 = w32 * scale
            if x.C > ci:   # producer padded its channel count (e.g. Cout of a tensor-core conv_out to a multiple of 8): zero columns
                w32 = torch.cat([w32, torch.zeros(co, x.C - ci, device=w32.device)], 1)
            self._packed = (key, w32.reshape(1, co, x.C).to(x.dtype).contiguous(), b32.contiguous())
            _publish(self._packed[1])   # read from every tile stream afterwards (run_tiles)
        y = N.conv3d_direct(x, self._packed[1], self._packed[2], 1, (1, 1, 1), co, round_like_ref=False)
        return y

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.forward_vol(Vol.from_ncthw(x)).to_ncthw()


_TILE_STREAMS = {}


def run_tiles(thunks, n_streams: int = 1):
    """Evaluate independent tile sub-model calls (thunks returning a tensor), optionally dealt round-robin over
    `n_streams` CUDA streams.  Every conv kernel is a persistent grid with one CTA per SM, so a second stream cannot
    steal SMs from a running conv; what it does is start the NEXT kernel's CTAs on the SMs that the tail wave of the
    current one leaves idle (the 17 x 32 x 32 mid-block layers fill only 136 of 148 SMs' worth of tiles), and let the
    HBM-bound GroupNorm / pad passes of one tile run under the tensor-bound convs of another.  Results do not depend
    on the interleaving: each kernel's tile schedule is static and the GroupNorm partial buffers are per stream.
    Lazily derived parameter tensors (packed weights, fp32 norm parameters) are created by whichever thunk needs them
    first and published with a stream synchronisation before any later thunk is enqueued (blocks._publish), so all lanes
    start at once."""
    thunks = list(thunks)
    if n_streams <= 1 or len(thunks) <= 2 or not torch.cuda.is_available():
        return [f() for f in thunks]
    main = torch.cuda.current_stream()
    dev = torch.cuda.current_device()
    side = _TILE_STREAMS.setdefault((dev, n_streams), [torch.cuda.Stream(device=dev) for _ in range(n_streams - 1)])
    for st in side:
        st.wait_stream(main)
    lanes = [main] + side
    outs = []
    for k, f in enumerate(thunks):
        st = lanes[k % n_streams]
        with torch.cuda.stream(st):
            o = f()
        if st is not main:
            o.record_stream(main)  # allocated in the side stream's pool, consumed (blend / gather) on the caller's stream
        outs.append(o)
    for st in side:
        main.wait_stream(st)
    return outs


class AutoencoderKLCausal3D(nn.Module):
    """autoencoder_kl_causal_3d.py:53-616 (the parts callers use)."""

    config_name = "config.json"
    _supports_gradient_checkpointing = False

    def __init__(self, in_channels: int = 3, out_channels: int = 3,
                 down_block_types: Tuple[str] = ("DownEncoderBlockCausal3D",),
                 up_block_types: Tuple[str] = ("UpDecoderBlockCausal3D",), block_out_channels: Tuple[int] = (64,),
                 layers_per_block: int = 1, act_fn: str = "silu", latent_channels: int = 4, norm_num_groups: int = 32,
                 sample_size: int = 32, sample_tsize: int = 64, scaling_factor: float = 0.18215,
                 force_upcast: float = True, spatial_compression_ratio: int = 8, time_compression_ratio: int = 4,
                 mid_block_add_attention: bool = True):
        super().__init__()
        sig = inspect.signature(AutoencoderKLCausal3D.__init__)
        loc = locals()
        self._internal_dict = FrozenConfig({k: loc[k] for k in list(sig.parameters)[1:]})
        self.time_compression_ratio = time_compression_ratio
        self.encoder = EncoderCausal3D(
            in_channels=in_channels, out_channels=latent_channels, down_block_types=down_block_types,
            block_out_channels=block_out_channels, layers_per_block=layers_per_block, act_fn=act_fn,
            norm_num_groups=norm_num_groups, double_z=True, time_compression_ratio=time_compression_ratio,
            spatial_compression_ratio=spatial_compression_ratio, mid_block_add_attention=mid_block_add_attention)
        self.decoder = DecoderCausal3D(
            in_channels=latent_channels, out_channels=out_channels, up_block_types=up_block_types,
            block_out_channels=block_out_channels, layers_per_block=layers_per_block, norm_num_groups=norm_num_groups,
            act_fn=act_fn, time_compression_ratio=time_compression_ratio,
            spatial_compression_ratio=spatial_compression_ratio, mid_block_add_attention=mid_block_add_attention)
        self.quant_conv = _ConvParamsOnly(2 * latent_channels)
        self.post_quant_conv = _ConvParamsOnly(latent_channels)
        self.use_slicing = self.use_spatial_tiling = self.use_temporal_tiling = False
        self.tile_sample_min_tsize = sample_tsize
        self.tile_latent_min_tsize = sample_tsize // time_compression_ratio
        self.tile_sample_min_size = sample_size
        ss = sample_size[0] if isinstance(sample_size, (list, tuple)) else sample_size
        self.tile_latent_min_size = int(ss / (2 ** (len(block_out_channels) - 1)))
        self.tile_overlap_factor = 0.25
        self.bf16_compute = "fp16"  # see _act_dtype()
        # fp16 operands have range 65 504 where the bf16 the model was built in has 3e38: a tile whose result is not finite
        # is re-run with bf16 operands (see _guard_tiles)
        self.fp16_range_guard = os.environ.get("HYVAE_FP16_GUARD", "1") == "1"
        self.range_guard_reruns = 0
        self._force_bf16 = False
        self._latent_affine = None   # (scale, shift) folded into post_quant_conv by decode_to_image
        # sub-model calls of different tiles are independent: run them on this many CUDA streams (run_tiles)
        self.tile_streams = int(os.environ.get("HYVAE_TILE_STREAMS", "2"))

    # ---- diffusers-style config / nn.Module conveniences
    @property
    def config(self):
        return self._internal_dict

    @classmethod
    def load_config(cls, path, **kw):
        p = path if str(path).endswith(".json") else os.path.join(path, cls.config_name)
        with open(p) as f:
            return json.load(f)

    @classmethod
    def from_config(cls, config, **kwargs):
        sig = inspect.signature(cls.__init__)
        cfg = {k: v for k, v in dict(config).items() if k in sig.parameters}
        cfg.update(kwargs)
        return cls(**cfg)

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    # ---- switches (:138-179)
    def enable_temporal_tiling(self, use_tiling: bool = True):
        self.use_temporal_tiling = use_tiling

    def disable_temporal_tiling(self):
        self.enable_temporal_tiling(False)

    def enable_spatial_tiling(self, use_tiling: bool = True):
        self.use_spatial_tiling = use_tiling

    def disable_spatial_tiling(self):
        self.enable_spatial_tiling(False)

    def enable_tiling(self, use_tiling: bool = True):
        self.enable_spatial_tiling(use_tiling)
        self.enable_temporal_tiling(use_tiling)

    def disable_tiling(self):
        self.disable_spatial_tiling()
        self.disable_temporal_tiling()

    def enable_slicing(self):
        self.use_slicing = True

    def disable_slicing(self):
        self.use_slicing = False

    # ---- one sub-model call on one tile (NCTHW in, NCTHW out)
    def _act_dtype(self):
        """Storage / tensor-core operand type of the activations inside a sub-model call.

        fp32 and fp16 models compute in their own type.  A bf16 model computes with FP16 operands and
        fp32 accumulation by default (`bf16_compute = "fp16"`, or env HYVAE_BF16_COMPUTE=fp16): the bf16
        weights convert to fp16 exactly, tcgen05 kind::f16 runs both types at the same rate, and fp16's
        three extra mantissa bits are what brings the result within BASELINE.json's 2e-2 / 45 dB of the
        exact evaluation (the reference's own all-bf16 run is 3-8e-2 away; DESIGN.md "Precision").  fp16 is
        the reference's default VAE precision (hyvideo/config.py:67-73), so its range is known to suffice.
        `bf16_compute = "bf16"` keeps every operand and intermediate in bf16 like the reference."""
        dt = self.dtype
        if dt == torch.bfloat16 and not self._force_bf16 and os.environ.get("HYVAE_BF16_COMPUTE", self.bf16_compute) == "fp16":
            return torch.float16
        return dt

    def _guard_tiles(self, outs, rerun):
        """Range guard of the fp16-operand mode of a bf16 model.  An activation beyond fp16's 65 504 becomes inf at the store
        that rounds it and reaches the tile's output as inf / NaN (the next GroupNorm turns it into NaN), so `all finite` on
        the outputs detects it: one fp32 sum per tile, ONE host read per batch of tiles.  Flagged tiles are recomputed by
        `rerun(k)` with every operand in bf16, exactly as `bf16_compute = "bf16"` evaluates them."""
        if not (self.fp16_range_guard and outs and self.dtype == torch.bfloat16 and self._act_dtype() == torch.float16):
            return outs
        flags = torch.stack([o.sum(dtype=torch.float32) for o in outs])
        bad = (~torch.isfinite(flags)).nonzero().flatten().tolist()
        if bad:
            self._force_bf16 = True
            try:
                for k in bad:
                    outs[k] = rerun(k)
                    self.range_guard_reruns += 1
            finally:
                self._force_bf16 = False
        return outs

    def _guarded(self, fn, x):
        return self._guard_tiles([fn(x)], lambda k: fn(x))[0]

    def _encode_tile(self, x: torch.Tensor) -> torch.Tensor:
        act = self._act_dtype()
        pad, ch = self.encoder.conv_in.input_layout(act)   # tensor-core conv_in: halo + 3->8 channels written here
        v = Vol.from_ncthw(x, dtype=act, pad=pad, channels=ch, kw_pack=self.encoder.conv_in.wants_kw_pack(act))
        return self.quant_conv.forward_vol(self.encoder.forward_vol(v)).to_ncthw(dtype=self.dtype)

    def _decode_tile(self, z: torch.Tensor) -> torch.Tensor:
        v = Vol.from_ncthw(z, dtype=self._act_dtype())
        return self.decoder.forward_vol(self.post_quant_conv.forward_vol(v, affine=self._latent_affine)).to_ncthw(dtype=self.dtype)

    # ---- encode / decode (:259-342)
    def encode(self, x: torch.Tensor, return_dict: bool = True):
        assert len(x.shape) == 5, "The input tensor should have 5 dimensions."
        if self.use_temporal_tiling and x.shape[2] > self.tile_sample_min_tsize:
            return self.temporal_tiled_encode(x, return_dict=return_dict)
        if self.use_spatial_tiling and (x.shape[-1] > self.tile_sample_min_size or x.shape[-2] > self.tile_sample_min_size):
            return self.spatial_tiled_encode(x, return_dict=return_dict)
        if self.use_slicing and x.shape[0] > 1:
            moments = torch.cat([self._guarded(self._encode_tile, s) for s in x.split(1)])
        else:
            moments = self._guarded(self._encode_tile, x)
        posterior = DiagonalGaussianDistribution(moments)
        if not return_dict:
            return (posterior,)
        return AutoencoderKLOutput(latent_dist=posterior, tiles_ci=None)

    def _decode(self, z: torch.Tensor, return_dict: bool = True, _post: bool = False):
        assert len(z.shape) == 5, "The input tensor should have 5 dimensions."
        if self.use_temporal_tiling and z.shape[2] > self.tile_latent_min_tsize:
            return self.temporal_tiled_decode(z, return_dict=return_dict, _post=_post)
        if self.use_spatial_tiling and (z.shape[-1] > self.tile_latent_min_size or z.shape[-2] > self.tile_latent_min_size):
            return self.spatial_tiled_decode(z, return_dict=return_dict, _post=_post)
        dec = self._guarded(self._decode_tile, z)
        if _post:
            dec = N.image_postprocess(dec)
        if not return_dict:
            return (dec,)
        return DecoderOutput(sample=dec)

    def decode(self, z: torch.Tensor, return_dict: bool = True, generator=None):
        if self.use_slicing and z.shape[0] > 1:
            decoded = torch.cat([self._decode(s).sample for s in z.split(1)])
        else:
            decoded = self._decode(z).sample
        if not return_dict:
            return (decoded,)
        return DecoderOutput(sample=decoded)

    def decode_to_image(self, z: torch.Tensor, latent_scale: float = 1.0, latent_shift: float = 0.0) -> torch.Tensor:
        """The pipeline tail in one call (pipeline_hunyuan_video.py:1060-1092): decode(z * latent_scale + latent_shift)
        followed by float((x / 2 + 0.5).clamp(0, 1)), returned as an fp32 tensor on the device.  The affine map of the
        latents is folded into post_quant_conv's packed weights, and the image post-processing is the epilogue of the
        last tile-assembly kernel (hyvae_blend_crop_scatter, post = 1), so neither costs a pass over the data."""
        self._latent_affine = None if (latent_scale == 1.0 and latent_shift == 0.0) else (float(latent_scale), float(latent_shift))
        try:
            if self.use_slicing and z.shape[0] > 1:
                return torch.cat([self._decode(s, _post=True).sample for s in z.split(1)])
            return self._decode(z, _post=True).sample
        finally:
            self._latent_affine = None

    def tiled_decode(self, z: torch.Tensor, return_dict: bool = True):
        """decode() with spatial + temporal tiling switched on for this call."""
        saved = (self.use_spatial_tiling, self.use_temporal_tiling)
        self.enable_tiling(True)
        try:
            return self.decode(z, return_dict=return_dict)
        finally:
            self.use_spatial_tiling, self.use_temporal_tiling = saved

    # ---- blends (:344-360), public like the reference's; in place on b, one kernel each
    @staticmethod
    def _blend(a: torch.Tensor, b: torch.Tensor, extent: int, axis: int) -> torch.Tensor:
        e = min(a.shape[axis], b.shape[axis], extent)
        if e <= 0:
            return b
        if a.ndim != 5 or b.ndim != 5 or a.dtype != b.dtype or a.device != b.device:
            raise ValueError("blend_*: a and b must be 5-D tensors of one dtype on one device")
        ax = axis % 5
        if any(a.shape[d] != b.shape[d] for d in range(5) if d != ax):   # the reference's tensor ops would raise as well
            raise ValueError(f"blend_*: shapes {tuple(a.shape)} and {tuple(b.shape)} differ outside the blended axis")
        assert a.is_contiguous() and b.is_contiguous(), "blend_* operate on contiguous tile tensors"
        B, C, T, H, W = b.shape
        if axis == -2:
            N.blend_crop_scatter(b, a, None, B * C * T, H, W, a.shape[-2], 0, e, 0, None, 0, 0, 0, 0, 0, 0)
        elif axis == -1:
            N.blend_crop_scatter(b, None, a, B * C * T, H, W, 0, a.shape[-1], 0, e, None, 0, 0, 0, 0, 0, 0)
        else:
            N.blend_crop_scatter(b, a, None, B * C, T, H * W, a.shape[-3], 0, e, 0, None, 0, 0, 0, 0, 0, 0)
        return b

    def blend_v(self, a, b, blend_extent):
        return self._blend(a, b, blend_extent, -2)

    def blend_h(self, a, b, blend_extent):
        return self._blend(a, b, blend_extent, -1)

    def blend_t(self, a, b, blend_extent):
        return self._blend(a, b, blend_extent, -3)

    # ---- spatial tiling (:362-469)
    @staticmethod
    def _spatial_cuts(x: torch.Tensor, tile: int, stride: int):
        """Tile origins of the reference's spatial loops (:387-396,441-450) and the grid's row count."""
        ii, jj = list(range(0, x.shape[-2], stride)), list(range(0, x.shape[-1], stride))
        return [(i, j) for i in ii for j in jj], len(ii), len(jj)

    def _spatial_tiled(self, x: torch.Tensor, fn, tile: int, stride: int, extent: int, limit: int, post: bool = False) -> torch.Tensor:
        cuts, ni, nj = self._spatial_cuts(x, tile, stride)
        outs = run_tiles([(lambda i=i, j=j: fn(x[:, :, :, i:i + tile, j:j + tile])) for i, j in cuts],
                         self.tile_streams if x.is_cuda else 1)
        outs = self._guard_tiles(outs, lambda k: fn(x[:, :, :, cuts[k][0]:cuts[k][0] + tile, cuts[k][1]:cuts[k][1] + tile]))
        rows = [outs[r * nj:(r + 1) * nj] for r in range(ni)]
        return self._assemble_spatial(rows, extent, limit, post)

    def _assemble_spatial(self, rows, extent: int, limit: int, post: bool = False) -> torch.Tensor:
        """Raster-order in-place blend chain + crop + scatter of a grid of tile tensors (one kernel per tile).
        post: the scattered result is the fp32 image float((v / 2 + 0.5).clamp(0, 1)) (decode_to_image)."""
        hs = [min(r[0].shape[-2], limit) for r in rows]
        ws = [min(t.shape[-1], limit) for t in rows[0]]
        B, C, T = rows[0][0].shape[:3]
        out = torch.empty((B, C, T, sum(hs), sum(ws)), dtype=torch.float32 if post else rows[0][0].dtype, device=rows[0][0].device)
        Yo, Xo, n = out.shape[-2], out.shape[-1], B * C * T
        y0 = 0
        for i, row in enumerate(rows):
            x0 = 0
            for j, t in enumerate(row):
                above = rows[i - 1][j] if i > 0 else None
                left = row[j - 1] if j > 0 else None
                ev = min(above.shape[-2], t.shape[-2], extent) if above is not None else 0
                eh = min(left.shape[-1], t.shape[-1], extent) if left is not None else 0
                if ev <= 0:
                    above = None
                if eh <= 0:
                    left = None
                N.blend_crop_scatter(t, above, left, n, t.shape[-2], t.shape[-1],
                                     above.shape[-2] if above is not None else 0, left.shape[-1] if left is not None else 0,
                                     ev, eh, out, Yo, Xo, y0, x0, hs[i], ws[j], post=post)
                x0 += ws[j]
            y0 += hs[i]
        return out

    def spatial_tiled_encode(self, x: torch.Tensor, return_dict: bool = True, return_moments: bool = False):
        stride = int(self.tile_sample_min_size * (1 - self.tile_overlap_factor))
        extent = int(self.tile_latent_min_size * self.tile_overlap_factor)
        moments = self._spatial_tiled(x, self._encode_tile, self.tile_sample_min_size, stride, extent,
                                      self.tile_latent_min_size - extent)
        if return_moments:
            return moments
        posterior = DiagonalGaussianDistribution(moments)
        if not return_dict:
            return (posterior,)
        return AutoencoderKLOutput(latent_dist=posterior)

    def spatial_tiled_decode(self, z: torch.Tensor, return_dict: bool = True, _post: bool = False):
        stride = int(self.tile_latent_min_size * (1 - self.tile_overlap_factor))
        extent = int(self.tile_sample_min_size * self.tile_overlap_factor)
        dec = self._spatial_tiled(z, self._decode_tile, self.tile_latent_min_size, stride, extent,
                                  self.tile_sample_min_size - extent, post=_post)
        if not return_dict:
            return (dec,)
        return DecoderOutput(sample=dec)

    # ---- temporal tiling (:471-541)
    def _temporal_tiled(self, x, fn_plain, spatial, tile_t, stride, extent, limit, min_size, post: bool = False) -> torch.Tensor:
        """spatial = (tile, stride, extent, limit) of the spatial split applied inside every temporal tile (:487-490,523-526).
        The sub-model calls of ALL temporal tiles are independent, so they are dealt over the tile streams in one batch
        (one join per direction instead of one per temporal tile); assembly then follows the reference's order: the
        spatial grid of each temporal tile, then the temporal chain."""
        starts = list(range(0, x.shape[2], stride))
        s_tile, s_stride, s_extent, s_limit = spatial
        thunks, reruns, layout = [], [], []   # layout: per temporal tile (first thunk index, rows, columns) or (index, 0, 0) when not split
        for i in starts:
            t = x[:, :, i:i + tile_t + 1]
            if self.use_spatial_tiling and (t.shape[-1] > min_size or t.shape[-2] > min_size):
                cuts, ni, nj = self._spatial_cuts(t, s_tile, s_stride)
                layout.append((len(thunks), ni, nj))
                for (a, b) in cuts:
                    thunks.append(lambda t=t, a=a, b=b: fn_plain(t[:, :, :, a:a + s_tile, b:b + s_tile]))
            else:
                layout.append((len(thunks), 0, 0))
                thunks.append(lambda t=t: fn_plain(t))
        outs = run_tiles(thunks, self.tile_streams if x.is_cuda else 1)
        outs = self._guard_tiles(outs, lambda k: thunks[k]())
        row = []  # (tensor, first_frame_offset): tiles i>0 drop their first output frame (:491,527)
        for n, (k0, ni, nj) in enumerate(layout):
            if ni == 0:
                t = outs[k0]
            else:
                t = self._assemble_spatial([outs[k0 + r * nj:k0 + (r + 1) * nj] for r in range(ni)], s_extent, s_limit)
            row.append((t, 1 if n > 0 else 0))
        return self._assemble_temporal(row, extent, limit, post)

    def _assemble_temporal(self, row, extent: int, limit: int, post: bool = False) -> torch.Tensor:
        """`row` = [(tile tensor, leading frames to drop)]: blend_t chain + crop + concatenate along T.
        post: as in _assemble_spatial (this is the last assembly pass of a temporally tiled decode)."""
        lens = [t.shape[2] - off for t, off in row]
        keep = [min(lens[i], limit + 1 if i == 0 else limit) for i in range(len(row))]
        B, C, _, H, W = row[0][0].shape
        hw = H * W
        out = torch.empty((B, C, sum(keep), H, W), dtype=torch.float32 if post else row[0][0].dtype, device=row[0][0].device)
        y0 = 0
        for i, (t, off) in enumerate(row):
            cur = t[:, :, off:]
            if i > 0:
                pt, poff = row[i - 1]
                e = min(lens[i - 1], lens[i], extent)
            else:
                pt, poff, e = None, 0, 0
            above = pt[:, :, poff:] if (pt is not None and e > 0) else None
            ns = (t.shape[2] * hw, (pt.shape[2] * hw) if pt is not None else 0, 0, out.shape[2] * hw)
            N.blend_crop_scatter(cur, above, None, B * C, lens[i], hw, lens[i - 1] if above is not None else 0, 0,
                                 e if above is not None else 0, 0, out, out.shape[2], hw, y0, 0, keep[i], hw, n_strides=ns, post=post)
            y0 += keep[i]
        return out

    def temporal_tiled_encode(self, x: torch.Tensor, return_dict: bool = True):
        stride = int(self.tile_sample_min_tsize * (1 - self.tile_overlap_factor))
        extent = int(self.tile_latent_min_tsize * self.tile_overlap_factor)
        s_extent = int(self.tile_latent_min_size * self.tile_overlap_factor)
        spatial = (self.tile_sample_min_size, int(self.tile_sample_min_size * (1 - self.tile_overlap_factor)), s_extent,
                   self.tile_latent_min_size - s_extent)                         # as spatial_tiled_encode
        moments = self._temporal_tiled(x, self._encode_tile, spatial,
                                       self.tile_sample_min_tsize, stride, extent, self.tile_latent_min_tsize - extent,
                                       self.tile_sample_min_size)
        posterior = DiagonalGaussianDistribution(moments)
        if not return_dict:
            return (posterior,)
        return AutoencoderKLOutput(latent_dist=posterior)

    def temporal_tiled_decode(self, z: torch.Tensor, return_dict: bool = True, _post: bool = False):
        stride = int(self.tile_latent_min_tsize * (1 - self.tile_overlap_factor))
        extent = int(self.tile_sample_min_tsize * self.tile_overlap_factor)
        s_extent = int(self.tile_sample_min_size * self.tile_overlap_factor)
        spatial = (self.tile_latent_min_size, int(self.tile_latent_min_size * (1 - self.tile_overlap_factor)), s_extent,
                   self.tile_sample_min_size - s_extent)                         # as spatial_tiled_decode
        dec = self._temporal_tiled(z, self._decode_tile, spatial,
                                   self.tile_latent_min_tsize, stride, extent, self.tile_sample_min_tsize - extent,
                                   self.tile_latent_min_size, post=_post)
        if not return_dict:
            return (dec,)
        return DecoderOutput(sample=dec)

    # ---- forward (:543-578)
    def forward(self, sample: torch.Tensor, sample_posterior: bool = False, return_dict: bool = True,
                return_posterior: bool = False, generator: Optional[torch.Generator] = None):
        posterior = self.encode(sample).latent_dist
        z = posterior.sample(generator=generator) if sample_posterior else posterior.mode()
        dec = self.decode(z).sample
        if not return_dict:
            return (dec, posterior) if return_posterior else (dec,)
        return DecoderOutput2(sample=dec, posterior=posterior if return_posterior else None)
