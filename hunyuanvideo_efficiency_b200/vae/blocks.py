"""Building blocks of the causal 3D VAE, scheduled onto libhyvae.so kernels.

Same class names, constructor arguments, sub-module / parameter names (hence the same 248 state-dict
keys) and t-ops hooks as /root/reference/hyvideo/vae/unet_causal_3d_blocks.py, so callers and
checkpoints are interchangeable.  What differs is everything underneath: activations travel between
blocks as channels-last `Vol`s in HBM, each `forward_vol` is a fixed schedule of C-ABI kernel calls
(include/hyvae.h), and nothing here computes with torch ops.  Public `forward(tensor)` wrappers take
and return NCTHW tensors like the reference modules do.
"""
from __future__ import annotations

import math
import os
from typing import Optional, Tuple

import torch
from torch import nn

from .. import _native as N
from .._native import Vol

_16BIT = (torch.bfloat16, torch.float16)


def _publish(t: torch.Tensor):
    """Derived parameter tensors (packed weights, fp32 norm parameters, summed biases) are created lazily on whatever stream
    first needs them and then read from EVERY tile stream (model.run_tiles) without further ordering: wait here, once per
    parameter version, until the kernels that produced them have finished."""
    if t.is_cuda:
        torch.cuda.current_stream(t.device).synchronize()


def _triple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v, v)


def tc_eligible(dtype, cin: int, cout: int, stride, k: int) -> bool:
    """Tensor-core (tcgen05) path iff the shape fits it; everything else runs the CUDA-core kernel."""
    if os.environ.get("HYVAE_FORCE_DIRECT", "0") == "1":
        return False
    del cin, cout, k  # any channel count works: operands are zero-padded to multiples of 8 (TMA fills the rest)
    return N.device_supports_tc() and dtype in _16BIT and stride[1] <= 2 and stride[2] <= 2


def prepare_causal_attention_mask(n_frame: int, n_hw: int, dtype, device, batch_size: int = None):
    """API parity with unet_causal_3d_blocks.py:38-46.  The CUDA attention path never builds this
    mask (hyvae_softmax_frame_causal derives it from indices); kept for callers that import it."""
    f = torch.arange(n_frame * n_hw, device=device) // n_hw
    mask = torch.zeros((n_frame * n_hw, n_frame * n_hw), dtype=dtype, device=device)
    mask.masked_fill_(f[None, :] > f[:, None], float("-inf"))
    if batch_size is not None:
        mask = mask.unsqueeze(0).expand(batch_size, -1, -1)
    return mask


class _Conv3dParams(nn.Module):
    """Parameter holder standing where the reference keeps an nn.Conv3d (`CausalConv3d.conv`):
    same `weight` [Cout,Cin,k,k,k] / `bias` names and a mutable `stride` (unet_causal_3d_blocks.py:741)."""

    def __init__(self, cin, cout, k, stride=1, bias=True):
        super().__init__()
        self.in_channels, self.out_channels = cin, cout
        self.kernel_size, self.stride = _triple(k), _triple(stride)
        self.weight = nn.Parameter(torch.empty(cout, cin, *self.kernel_size))
        self.bias = nn.Parameter(torch.empty(cout)) if bias else None
        bound = 1.0 / math.sqrt(cin * self.kernel_size[0] * self.kernel_size[1] * self.kernel_size[2])
        nn.init.uniform_(self.weight, -bound, bound)
        if bias:
            nn.init.uniform_(self.bias, -bound, bound)
        self._packed = None

    def cin_padded(self) -> int:
        """Input channels the tensor-core kernels see: multiples of 8 (16-byte rows); a thin stride-1 3x3x3 layer with a
        wide output (conv_in, 3 -> 128) is stored as 16 channels = one K = 16 MMA slice per 32-byte row (conv_halo.cu, THIN)."""
        ci, co = self.in_channels, self.out_channels
        if ci < 16 and 64 < co <= 128 and self.kernel_size[0] == 3 and tuple(int(v) for v in self.stride) == (1, 1, 1):
            return 16
        return -(-ci // 8) * 8

    def packed(self, dtype, pad8: bool = False, tfold: bool = False):
        """[taps][Cout][Cin] weights in the activation dtype + fp32 bias (kernel layout), cached.  pad8 zero-pads
        Cout up to a multiple of 8 and Cin up to cin_padded() (16-byte rows) for the tensor-core kernel.  tfold (3x3x3 only)
        appends the 18 folded first-frame tap slices the halo / kh-trick kernels use for output frames 0 and 1 (csrc/tcgen05.cuh
        tfold_class): [27..35] = W[kt=0] + W[1] + W[2], [36..44] = W[0] + W[1], summed in fp32 and rounded once."""
        w = self.weight
        key = (w._version, w.data_ptr(), dtype, w.device, None if self.bias is None else self.bias._version, pad8,
               tuple(int(v) for v in self.stride), tfold)
        if self._packed is None or self._packed[0] != key:
            k = self.kernel_size[0]
            co, ci = self.out_channels, self.in_channels
            w32 = w.detach().float().permute(2, 3, 4, 0, 1)                      # [kt][kh][kw][Cout][Cin]
            if tfold:
                assert k == 3
                w32 = torch.cat([w32, w32.sum(0, keepdim=True), w32[0:1] + w32[1:2]], 0)
            pw = w32.reshape(-1, co, ci).to(dtype)
            pb = None if self.bias is None else self.bias.detach().float()
            if pad8 and (co % 8 or ci != self.cin_padded()):
                cop, cip = -(-co // 8) * 8, self.cin_padded()
                full = torch.zeros((pw.shape[0], cop, cip), dtype=dtype, device=w.device)
                full[:, :co, :ci] = pw
                pw = full
                if pb is not None:
                    pb = torch.cat([pb, torch.zeros(cop - co, device=w.device)])
            self._packed = (key, pw.contiguous(), None if pb is None else pb.contiguous())
            _publish(self._packed[1])
        return self._packed[1], self._packed[2]

    def wino_packed(self, dtype):
        """[5 * 9][Cout][Cin] tap groups of the Winograd-T form of this stride-1 3x3x3 conv (csrc/conv_wino.cu):
        g0, (g0 + g1 + g2) / 2, (g0 - g1 + g2) / 2, g2 and g0 + g1 + g2 (g_kt = W[:, :, kt]), each [(kh, kw)][Cout][Cin];
        formed in fp32 and rounded once.  Cached per parameter version."""
        w = self.weight
        key = ("wino", w._version, w.data_ptr(), dtype, w.device, None if self.bias is None else self.bias._version)
        cached = getattr(self, "_wino_packed", None)
        if cached is None or cached[0] != key:
            assert self.kernel_size[0] == 3
            w32 = w.detach().float().permute(2, 3, 4, 0, 1)                      # [kt][kh][kw][Cout][Cin]
            g0, g1, g2 = w32[0], w32[1], w32[2]
            u = torch.stack([g0, (g0 + g1 + g2) / 2, (g0 - g1 + g2) / 2, g2, g0 + g1 + g2], 0)
            pw = u.reshape(45, self.out_channels, self.in_channels).to(dtype).contiguous()
            pb = None if self.bias is None else self.bias.detach().float().contiguous()
            self._wino_packed = cached = (key, pw, pb)
            _publish(pw)
        return cached[1], cached[2]

    def kw_packable(self) -> bool:
        """conv_in-like layer served by the thin halo kernel whose three kw taps fit one 16-channel row (3 * Cin <= 16)."""
        return (self.cin_padded() == 16 and 3 * self.in_channels <= 16 and self.kernel_size[0] == 3
                and os.environ.get("HYVAE_KWPACK", "1") == "1")

    def packed_kw(self, dtype):
        """[9 = kt*3+kh][Cout][16] weights for a kw-packed input (Vol.from_ncthw(kw_pack=True)): column kw * Cin + c of tap
        (kt, kh) is W[:, c, kt, kh, kw]; the other columns are zero."""
        w = self.weight
        key = ("kw", w._version, w.data_ptr(), dtype, w.device, None if self.bias is None else self.bias._version)
        cached = getattr(self, "_packed_kw", None)
        if cached is None or cached[0] != key:
            co, ci = self.out_channels, self.in_channels
            pw = torch.zeros((9, co, 16), dtype=dtype, device=w.device)
            pw[:, :, :3 * ci] = w.detach().permute(2, 3, 0, 4, 1).reshape(9, co, 3 * ci).to(dtype)   # [kt][kh][Cout][kw][Cin]
            pb = None if self.bias is None else self.bias.detach().float().contiguous()
            self._packed_kw = cached = (key, pw.contiguous(), pb)
            _publish(cached[1])
        return cached[1], cached[2]

    def phase_packed(self, dtype, up):
        """Weights of the sub-pixel phases of `nearest-upsample(up) -> this 3x3x3 conv` (hyvae_conv3d_upphase_tc):
        {(pt, ph, pw): [nkt*2*2][Cout][Cin]}.  Along an upsampled axis the three taps fold onto two low-res taps,
        even outputs use (W0, W1+W2), odd outputs (W0+W1, W2); sums are formed in fp32 and rounded once."""
        w = self.weight
        key = ("phase", w._version, w.data_ptr(), dtype, w.device, tuple(up), None if self.bias is None else self.bias._version)
        cached = getattr(self, "_phase_packed", None)
        if cached is None or cached[0] != key:
            w32 = w.detach().float()
            fold = (torch.tensor([[1., 0., 0.], [0., 1., 1.]], device=w.device), torch.tensor([[1., 1., 0.], [0., 0., 1.]], device=w.device))
            eye = torch.eye(3, device=w.device)
            out = {}
            for pt in range(2 if up[0] == 2 else 1):
                mt = fold[pt] if up[0] == 2 else eye
                for ph in range(2):
                    for pw in range(2):
                        wp = torch.einsum("at,bh,cw,oithw->abcoi", mt, fold[ph], fold[pw], w32)
                        out[(pt, ph, pw)] = wp.reshape(-1, self.out_channels, self.in_channels).to(dtype).contiguous()
            pb = None if self.bias is None else self.bias.detach().float().contiguous()
            self._phase_packed = cached = (key, out, pb)
            _publish(w)
        return cached[1], cached[2]


class CausalConv3d(nn.Module):
    """unet_causal_3d_blocks.py:49-75.  `time_causal_padding` is kept as an attribute; the padding
    itself is a halo of the input volume (tensor-core path) or index clamping (CUDA-core path)."""

    def __init__(self, chan_in, chan_out, kernel_size, stride=1, dilation=1, pad_mode="replicate", **kwargs):
        super().__init__()
        if _triple(dilation) != (1, 1, 1):
            raise NotImplementedError("dilation != 1")
        if pad_mode != "replicate":
            raise NotImplementedError(f"pad_mode {pad_mode}")
        if not isinstance(kernel_size, int) or kernel_size not in (1, 3):
            raise NotImplementedError(f"kernel_size {kernel_size}")
        self.pad_mode = pad_mode
        k = kernel_size
        self.time_causal_padding = (k // 2, k // 2, k // 2, k // 2, k - 1, 0)
        self.conv = _Conv3dParams(chan_in, chan_out, k, stride, bias=kwargs.get("bias", True))
        self.emit_gn_groups = 0  # >0: the consumer is a GroupNorm with that many groups -> stats from the epilogue

    @property
    def halo(self) -> Tuple[int, int, int]:
        k = self.conv.kernel_size[0]
        return (k - 1, k // 2, k // 2)

    def wants_halo(self, dtype) -> Tuple[int, int, int]:
        """Halo the producer should write so this conv can read its input through TMA."""
        c = self.conv
        return self.halo if tc_eligible(dtype, c.in_channels, c.out_channels, c.stride, c.kernel_size[0]) else (0, 0, 0)

    def wants_wino(self, dtype) -> bool:
        """True when the GroupNorm that feeds this conv should write the Winograd-T plane volume (csrc/conv_wino.cu): stride-1
        3x3x3, fp16 operands (like the sub-pixel phases, the transformed taps need fp16's mantissa), Cin % 64 == 0 and
        Cout % 128 == 0.  HYVAE_WINO=0 switches it off (A/B measurements, parity tests against the plain path)."""
        c = self.conv
        return (os.environ.get("HYVAE_WINO", "1") == "1" and dtype == torch.float16 and c.kernel_size[0] == 3
                and tuple(int(v) for v in c.stride) == (1, 1, 1) and c.in_channels % 64 == 0 and c.out_channels % 128 == 0
                and c.bias is not None and tc_eligible(dtype, c.in_channels, c.out_channels, (1, 1, 1), 3))

    def wants_kw_pack(self, dtype) -> bool:
        """True when the producer (the NCTHW -> volume layout pass) should write the kw-packed operand (conv_in)."""
        c = self.conv
        return (c.kw_packable() and tuple(int(v) for v in c.stride) == (1, 1, 1) and 64 < c.out_channels <= 128
                and tc_eligible(dtype, c.in_channels, c.out_channels, c.stride, 3))

    def input_layout(self, dtype):
        """(halo, channel count) a producer should write so that this conv needs no extra pad pass."""
        c = self.conv
        if tc_eligible(dtype, c.in_channels, c.out_channels, c.stride, c.kernel_size[0]):
            return self.halo, c.cin_padded()
        return (0, 0, 0), c.in_channels

    def forward_vol(self, x: Vol, residual: Optional[Vol] = None, up=(1, 1, 1), out_dtype=None, out_pad=(0, 0, 0)) -> Vol:
        """out_pad (tensor-core path only): write the result into the interior of a volume with that halo and replicate
        the halo afterwards (N.halo_fill), so that the next CausalConv3d needs no full-tensor pad pass."""
        c = self.conv
        k, stride = c.kernel_size[0], tuple(int(s) for s in c.stride)
        rl = False  # one rounding per stored tensor: conv + bias + residual are summed in fp32, then stored
        if x.wino_T:   # the producer (GroupNorm) wrote the Winograd-T planes: 4 instead of 6 plane-GEMMs per frame pair
            assert self.wants_wino(x.dtype) and x.C == c.in_channels and up == (1, 1, 1) and out_dtype in (None, x.dtype)
            w, b = c.wino_packed(x.dtype)
            y = N.conv3d_wino(x, w, b, c.out_channels, residual, out_pad, gn_groups=self.emit_gn_groups)
            N.halo_fill(y)
            return y
        if x.kw_packed:   # conv_in on the operand the layout pass packed for it: 9 (kt, kh) taps of K = 16
            assert self.wants_kw_pack(x.dtype) and x.pad == self.halo and x.C == 16 and residual is None and up == (1, 1, 1)
            w, b = c.packed_kw(x.dtype)
            y = N.conv3d_tc(x, w, b, 3, (1, 1, 1), c.out_channels, None, out_dtype, rl, gn_groups=self.emit_gn_groups,
                            variant=N.VARIANT_KWPACK)
            y.c_valid = c.out_channels
            return y
        if tc_eligible(x.dtype, c.in_channels, c.out_channels, stride, k):
            # first-frame temporal fold: fp16 operands only (like the sub-pixel phases, the folded sums need fp16's mantissa)
            tfold = (k == 3 and stride == (1, 1, 1) and x.dtype == torch.float16 and out_dtype in (None, torch.float16)
                     and os.environ.get("HYVAE_TFOLD", "1") == "1")
            w, b = c.packed(x.dtype, pad8=True, tfold=tfold)
            cin_p, cout_p = w.shape[2], w.shape[1]
            if x.pad != self.halo or up != (1, 1, 1) or x.C != cin_p:
                x = N.pad_upsample(x, up, self.halo, channels=cin_p)
            out = None
            if tuple(out_pad) != (0, 0, 0):
                To, Ho, Wo = N.conv_out_dims(x.T, x.H, x.W, stride)
                out = Vol(x.B, To, Ho, Wo, cout_p, out_dtype or x.dtype, x.device, tuple(out_pad))
            y = N.conv3d_tc(x, w, b, k, stride, cout_p, residual, out_dtype, rl, gn_groups=self.emit_gn_groups,
                            variant=N.VARIANT_TFOLD if tfold else 0, out=out)
            N.halo_fill(y)
            y.c_valid = c.out_channels
            return y
        w, b = c.packed(x.dtype)
        if x.C != c.in_channels:
            raise N.HyvaeError(f"CausalConv3d: input has {x.C} channels, expected {c.in_channels}")
        return N.conv3d_direct(x, w, b, k, stride, c.out_channels, residual, up, out_dtype, rl)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.forward_vol(Vol.from_ncthw(x)).to_ncthw()


class _GroupNorm(nn.Module):
    """nn.GroupNorm-compatible parameters (`weight`, `bias`); the compute is hyvae_groupnorm_*."""

    def __init__(self, num_groups, num_channels, eps=1e-6):
        super().__init__()
        self.num_groups, self.num_channels, self.eps = num_groups, num_channels, eps
        self.weight = nn.Parameter(torch.ones(num_channels))
        self.bias = nn.Parameter(torch.zeros(num_channels))
        self._f32 = None

    def _params(self):
        key = (self.weight._version, self.bias._version, self.weight.data_ptr(), self.weight.device)
        if self._f32 is None or self._f32[0] != key:
            self._f32 = (key, self.weight.detach().float().contiguous(), self.bias.detach().float().contiguous())
            _publish(self._f32[1])
        return self._f32[1], self._f32[2]

    def forward_vol(self, x: Vol, silu: bool, pad=(0, 0, 0), wino: bool = False) -> Vol:
        """wino: write the Winograd-T plane volume the consumer conv asked for (CausalConv3d.wants_wino) instead of `pad`."""
        g, b = self._params()
        if wino:
            return N.groupnorm_wino(x, g, b, self.num_groups, self.eps, silu)
        return N.groupnorm(x, g, b, self.num_groups, self.eps, silu, pad, False)


class UpsampleCausal3D(nn.Module):
    """unet_causal_3d_blocks.py:78-183: nearest upsample (frame 0 spatial only) + CausalConv3d.  The
    upsample is folded into the conv's gather (CUDA-core path) or into the pad pass that builds the
    TMA halo (tensor-core path); no fp32 round trip is needed because nothing is interpolated."""

    def __init__(self, channels, use_conv=False, use_conv_transpose=False, out_channels=None, name="conv",
                 kernel_size=None, padding=1, norm_type=None, eps=None, elementwise_affine=None, bias=True,
                 interpolate=True, upsample_factor=(2, 2, 2)):
        super().__init__()
        if use_conv_transpose or norm_type is not None:
            raise NotImplementedError
        self.channels, self.out_channels = channels, out_channels or channels
        self.use_conv, self.name, self.interpolate = use_conv, name, interpolate
        self.upsample_factor = tuple(upsample_factor)
        # fp16 activations only: the folded weights (W0+W1, ...) need fp16's mantissa; bf16 operands keep the plain path
        self.phase_decomposition = os.environ.get("HYVAE_UPSAMPLE_PHASES", "1") == "1"
        if any(f not in (1, 2) for f in self.upsample_factor):
            raise NotImplementedError(f"upsample_factor {upsample_factor}")
        conv = CausalConv3d(self.channels, self.out_channels, kernel_size=kernel_size or 3, bias=bias) if use_conv else None
        if conv is not None:
            conv.emit_gn_groups = 32  # next consumer: the following resnet's norm1 (ignored unless its groups match)
        if name == "conv":
            self.conv = conv
        else:
            self.Conv2d_0 = conv

    def _uses_phases(self, dtype) -> bool:
        up = self.upsample_factor if self.interpolate else (1, 1, 1)
        conv = self.conv if self.name == "conv" else self.Conv2d_0
        if conv is None:
            return False
        c = conv.conv
        return (self.phase_decomposition and dtype == torch.float16 and up[1:] == (2, 2) and c.kernel_size[0] == 3
                and tuple(int(v) for v in c.stride) == (1, 1, 1) and c.in_channels % 8 == 0 and c.out_channels % 8 == 0
                and c.out_channels >= 64 and tc_eligible(dtype, c.in_channels, c.out_channels, (1, 1, 1), 3))

    def wants_halo(self, dtype) -> Tuple[int, int, int]:
        """Halo of the LOW-res input the sub-pixel phase path reads through TMA ((0,0,0): the pad+upsample pass builds it)."""
        return (1 if self.upsample_factor[0] == 2 else 2, 1, 1) if self._uses_phases(dtype) else (0, 0, 0)

    def forward_vol(self, x: Vol) -> Vol:
        assert x.C == self.channels
        up = self.upsample_factor if self.interpolate else (1, 1, 1)
        conv = self.conv if self.name == "conv" else self.Conv2d_0
        if conv is None:
            return N.pad_upsample(x, up)
        c = conv.conv
        if self._uses_phases(x.dtype):
            # sub-pixel phases over the low-res tensor: 3.4x (2.25x) fewer MACs, no 8x upsampled intermediate
            pw, pb = c.phase_packed(x.dtype, up)
            halo = self.wants_halo(x.dtype)
            xp = x if x.pad == halo else N.pad_upsample(x, (1, 1, 1), halo)
            return N.conv3d_upsample_phases(xp, pw, pb, up, c.out_channels, gn_groups=conv.emit_gn_groups)
        return conv.forward_vol(x, up=up)

    def forward(self, hidden_states, output_size=None, scale: float = 1.0):
        if output_size is not None:
            raise NotImplementedError
        return self.forward_vol(Vol.from_ncthw(hidden_states)).to_ncthw()


class DownsampleCausal3D(nn.Module):
    """unet_causal_3d_blocks.py:186-247: one strided CausalConv3d."""

    def __init__(self, channels, use_conv=False, out_channels=None, padding=1, name="conv", kernel_size=3,
                 norm_type=None, eps=None, elementwise_affine=None, bias=True, stride=2):
        super().__init__()
        if not use_conv or norm_type is not None:
            raise NotImplementedError
        self.channels, self.out_channels = channels, out_channels or channels
        self.use_conv, self.padding, self.name = use_conv, padding, name
        self.conv = CausalConv3d(self.channels, self.out_channels, kernel_size=kernel_size, stride=stride, bias=bias)
        self.conv.emit_gn_groups = 32
        if name == "conv":
            self.Conv2d_0 = self.conv

    def forward_vol(self, x: Vol) -> Vol:
        assert x.C == self.channels
        return self.conv.forward_vol(x)

    def forward(self, hidden_states, scale: float = 1.0):
        return self.forward_vol(Vol.from_ncthw(hidden_states)).to_ncthw()


class ResnetBlockCausal3D(nn.Module):
    """unet_causal_3d_blocks.py:250-417 for the configuration the VAE uses (no temb, GroupNorm, swish).
    Schedule: GN-stats, GN+SiLU (writes conv1's halo), conv1, GN-stats, GN+SiLU, [1x1x1 shortcut],
    conv2 with the residual add fused into its epilogue."""

    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout=0.0, temb_channels=512,
                 groups=32, groups_out=None, pre_norm=True, eps=1e-6, non_linearity="swish", skip_time_act=False,
                 time_embedding_norm="default", kernel=None, output_scale_factor=1.0, use_in_shortcut=None,
                 up=False, down=False, conv_shortcut_bias=True, conv_3d_out_channels=None):
        super().__init__()
        if temb_channels is not None or time_embedding_norm != "default" or up or down:
            raise NotImplementedError("only the temb-free GroupNorm resnet of the VAE is implemented")
        if non_linearity not in ("swish", "silu"):
            raise NotImplementedError(non_linearity)
        if dropout != 0.0:
            raise NotImplementedError("dropout")
        out_channels = in_channels if out_channels is None else out_channels
        self.in_channels, self.out_channels = in_channels, out_channels
        self.output_scale_factor = output_scale_factor
        self.norm1 = _GroupNorm(groups, in_channels, eps)
        self.conv1 = CausalConv3d(in_channels, out_channels, kernel_size=3, stride=1)
        self.norm2 = _GroupNorm(groups_out or groups, out_channels, eps)
        c3 = conv_3d_out_channels or out_channels
        self.conv2 = CausalConv3d(out_channels, c3, kernel_size=3, stride=1)
        self.use_in_shortcut = (in_channels != c3) if use_in_shortcut is None else use_in_shortcut
        self.conv_shortcut = (CausalConv3d(in_channels, c3, kernel_size=1, stride=1, bias=conv_shortcut_bias)
                              if self.use_in_shortcut else None)
        # both convs feed a GroupNorm (norm2; the next block's norm1 / conv_norm_out / the attention's group_norm)
        self.conv1.emit_gn_groups = self.norm2.num_groups
        self.conv2.emit_gn_groups = groups

    def forward_vol(self, x: Vol, out_pad=(0, 0, 0)) -> Vol:
        """out_pad: halo the block's consumer (a down / upsampler conv) wants on the result; see CausalConv3d.forward_vol."""
        if self.output_scale_factor != 1.0:
            raise NotImplementedError("output_scale_factor != 1")
        h = self.norm1.forward_vol(x, True, self.conv1.wants_halo(x.dtype), wino=self.conv1.wants_wino(x.dtype))
        h = self.conv1.forward_vol(h)
        w2 = self.conv2.wants_wino(x.dtype) and (self.conv_shortcut is None or self._can_fuse_wino_shortcut(x))
        h = self.norm2.forward_vol(h, True, self.conv2.wants_halo(x.dtype), wino=w2)
        if w2 and self.conv_shortcut is not None:
            return self._conv2_wino_with_shortcut(h, x, out_pad)
        if self.conv_shortcut is not None and self._can_fuse_shortcut(h, x):
            return self._conv2_with_shortcut(h, x, out_pad)
        skip = x if self.conv_shortcut is None else self.conv_shortcut.forward_vol(x)
        return self.conv2.forward_vol(h, residual=skip, out_pad=out_pad)

    def _can_fuse_shortcut(self, h: Vol, x: Vol) -> bool:
        """The 1x1x1 conv_shortcut runs as extra K chunks of conv2's accumulator (hyvae_conv3d_causal_tc_shortcut) when
        conv2 is served by the halo kernel or the kh-trick pair kernel: 16-bit tensor-core path, Cout > 64."""
        c2, cs = self.conv2.conv, self.conv_shortcut.conv
        return (os.environ.get("HYVAE_FUSE_SHORTCUT", "1") == "1" and h.dtype in _16BIT and 64 < c2.out_channels
                and (c2.out_channels <= 128 or (h.H + 15) // 16 * ((h.W + 7) // 8) * h.T * h.B >= 2)
                and c2.out_channels % 8 == 0 and cs.in_channels % 8 == 0 and x.C == cs.in_channels and h.pad == self.conv2.halo
                and h.C == c2.in_channels and c2.bias is not None and cs.bias is not None
                and tc_eligible(h.dtype, c2.in_channels, c2.out_channels, (1, 1, 1), 3))

    def _can_fuse_wino_shortcut(self, x: Vol) -> bool:
        """The 1x1x1 conv_shortcut as extra K chunks of the Winograd-T kernel's accumulators 0 and 3 (csrc/conv_wino.cu)."""
        c2, cs = self.conv2.conv, self.conv_shortcut.conv
        return (os.environ.get("HYVAE_FUSE_SHORTCUT", "1") == "1" and cs.in_channels % 8 == 0 and x.C == cs.in_channels
                and c2.bias is not None and cs.bias is not None and cs.kernel_size[0] == 1)

    def _conv2_wino_with_shortcut(self, h: Vol, x: Vol, out_pad=(0, 0, 0)) -> Vol:
        c2, cs = self.conv2.conv, self.conv_shortcut.conv
        w2, b2 = c2.wino_packed(h.dtype)
        ws, bs = cs.packed(h.dtype, pad8=True)
        key = ("wino", c2.bias._version, cs.bias._version, cs.weight._version, b2.data_ptr(), ws.data_ptr())
        if getattr(self, "_wino_sc", None) is None or self._wino_sc[0] != key:
            self._wino_sc = (key, (b2 + bs).contiguous(), torch.cat([ws, -ws], 0).contiguous())   # bias sum; (+Ws, -Ws)
            _publish(self._wino_sc[2])
        y = N.conv3d_wino(h, w2, self._wino_sc[1], c2.out_channels, None, tuple(out_pad), gn_groups=self.conv2.emit_gn_groups,
                          sc_x=x, sc_w=self._wino_sc[2])
        return N.halo_fill(y)

    def _conv2_with_shortcut(self, h: Vol, x: Vol, out_pad=(0, 0, 0)) -> Vol:
        c2, cs = self.conv2.conv, self.conv_shortcut.conv
        tfold = h.dtype == torch.float16 and os.environ.get("HYVAE_TFOLD", "1") == "1"   # as in CausalConv3d.forward_vol
        w2, b2 = c2.packed(h.dtype, pad8=True, tfold=tfold)
        ws, bs = cs.packed(h.dtype, pad8=True)
        key = (c2.bias._version, cs.bias._version, b2.data_ptr(), bs.data_ptr())
        if getattr(self, "_bias_sum", None) is None or self._bias_sum[0] != key:
            self._bias_sum = (key, (b2 + bs).contiguous())
            _publish(self._bias_sum[1])
        try:
            return N.halo_fill(N.conv3d_tc_shortcut(h, w2, self._bias_sum[1], x, ws, c2.out_channels, gn_groups=self.conv2.emit_gn_groups,
                                                    tfold=tfold, out_pad=tuple(out_pad)))
        except N.HyvaeUnsupported:   # tile shape the fused kernels do not take: run the shortcut as its own k=1 conv
            return self.conv2.forward_vol(h, residual=self.conv_shortcut.forward_vol(x), out_pad=out_pad)

    def forward(self, input_tensor, temb=None, scale: float = 1.0):
        return self.forward_vol(Vol.from_ncthw(input_tensor)).to_ncthw()


class _Linear(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.in_features, self.out_features = cin, cout
        self.weight = nn.Parameter(torch.empty(cout, cin))
        self.bias = nn.Parameter(torch.empty(cout))
        bound = 1.0 / math.sqrt(cin)
        nn.init.uniform_(self.weight, -bound, bound)
        nn.init.uniform_(self.bias, -bound, bound)
        self._packed = None

    def packed(self, dtype):
        key = (self.weight._version, self.bias._version, self.weight.data_ptr(), dtype)
        if self._packed is None or self._packed[0] != key:
            self._packed = (key, self.weight.detach().to(dtype).contiguous(), self.bias.detach().float().contiguous())
            _publish(self._packed[1])
        return self._packed[1], self._packed[2]


def _gemm_nt(x: Vol, w: torch.Tensor, bias, cout: int, residual: Optional[Vol] = None, out_dtype=None,
             out: Optional[Vol] = None, gn_groups: int = 0) -> Vol:
    """y[m][n] = sum_k x[m][k] * w[n][k] (+bias[n]) (+residual): a 1x1x1 'conv' on either kernel.  gn_groups > 0 (tensor-core
    path): GroupNorm statistics of y from the epilogue (y.gn_sums)."""
    rl = False
    if tc_eligible(x.dtype, x.C, cout, (1, 1, 1), 1) and x.pad == (0, 0, 0) and x.C % 8 == 0 and cout % 8 == 0:
        return N.conv3d_tc(x, w, bias, 1, (1, 1, 1), cout, residual, out_dtype, rl, out=out, gn_groups=gn_groups)
    return N.conv3d_direct(x, w, bias, 1, (1, 1, 1), cout, residual, (1, 1, 1), out_dtype, rl, out=out)


class Attention(nn.Module):
    """The diffusers==0.31.0 `Attention` the reference instantiates at unet_causal_3d_blocks.py:580-592
    (one head of width C, GroupNorm(32), biased q/k/v/out projections, residual), with the frame-causal
    mask of :38-46 computed from indices.  Parameter names match diffusers' so checkpoints load.
    Schedule per batch item: GN -> Q,K = X Wq^T, X Wk^T -> V^T = Wv X^T -> S = Q K^T (fp32) ->
    masked softmax -> O = P V + bv -> out-proj (+bias +residual in the epilogue)."""

    def __init__(self, query_dim, heads=1, dim_head=None, rescale_output_factor=1.0, eps=1e-6, norm_num_groups=32,
                 spatial_norm_dim=None, residual_connection=True, bias=True, upcast_softmax=True,
                 _from_deprecated_attn_block=True, **unused):
        super().__init__()
        dim_head = dim_head or query_dim
        if heads != 1 or dim_head != query_dim or spatial_norm_dim is not None or not bias or rescale_output_factor != 1.0:
            raise NotImplementedError("only the single-head VAE attention block is implemented")
        self.heads, self.inner_dim, self.scale = heads, dim_head, dim_head ** -0.5
        self.residual_connection, self.rescale_output_factor = residual_connection, rescale_output_factor
        self.group_norm = _GroupNorm(norm_num_groups, query_dim, eps)
        self.to_q, self.to_k, self.to_v = _Linear(query_dim, dim_head), _Linear(query_dim, dim_head), _Linear(query_dim, dim_head)
        self.to_out = nn.ModuleList([_Linear(dim_head, query_dim), nn.Identity()])
        self.emit_gn_groups = norm_num_groups or 0   # the consumer is the mid block's second resnet (same group count)

    @staticmethod
    def fused_eligible(dtype, L: int, Cn: int) -> bool:
        """The flash-style tcgen05 kernel takes 16-bit operands, head widths 128/256/512 and L % 8 == 0 (TMA row pitch of
        V^T); HYVAE_ATTN_UNFUSED=1 forces the GEMM -> softmax -> GEMM schedule (A/B measurements, parity tests)."""
        if os.environ.get("HYVAE_ATTN_UNFUSED", "0") == "1" or os.environ.get("HYVAE_FORCE_DIRECT", "0") == "1":
            return False
        return N.device_supports_tc() and dtype in _16BIT and Cn in (128, 256, 512) and L % 8 == 0

    def forward_vol(self, x: Vol) -> Vol:
        B, T, H, W, Cn = x.dims
        L, n_hw = T * H * W, H * W
        hn = self.group_norm.forward_vol(x, False)
        out = x.like()
        wq, bq = self.to_q.packed(x.dtype)
        wk, bk = self.to_k.packed(x.dtype)
        wv, bv = self.to_v.packed(x.dtype)
        wo, bo = self.to_out[0].packed(x.dtype)
        for b in range(B):
            xb = Vol(1, 1, 1, L, Cn, x.dtype, x.device, tensor=hn.t[b].reshape(1, 1, 1, L, Cn))
            with N.profile_class("attn_proj"):   # Linear layers of the attention, not nn.Conv3d FLOPs of the reference
                q = _gemm_nt(xb, wq, bq, Cn)
                k = _gemm_nt(xb, wk, bk, Cn)
                wv_rows = Vol(1, 1, 1, Cn, Cn, x.dtype, x.device, tensor=wv.reshape(1, 1, 1, Cn, Cn))
                vt = _gemm_nt(wv_rows, xb.t.reshape(L, Cn), None, L)                  # V^T [C][L], bias folded below
            o = None
            if self.fused_eligible(x.dtype, L, Cn):
                try:  # one kernel: S and P stay in TMEM / shared memory (attn_fused.cu)
                    ot = N.attn_block_causal(q.t.reshape(L, Cn), k.t.reshape(L, Cn), vt.t.reshape(Cn, L), bv, n_hw, self.scale)
                    o = Vol(1, 1, 1, L, Cn, x.dtype, x.device, tensor=ot.reshape(1, 1, 1, L, Cn))
                except N.HyvaeUnsupported:
                    o = None
            if o is None:
                with N.profile_class("attn"):
                    s = _gemm_nt(q, k.t.reshape(L, Cn), None, L, out_dtype=torch.float32)     # S [L][L] fp32
                    p = N.softmax_frame_causal(s.t.reshape(1, L, L), n_hw, self.scale, x.dtype)
                    pv = Vol(1, 1, 1, L, L, x.dtype, x.device, tensor=p.reshape(1, 1, 1, L, L))
                    o = _gemm_nt(pv, vt.t.reshape(Cn, L), bv, Cn)                             # rows of P sum to 1 => + bv
            res = Vol(1, 1, 1, L, Cn, x.dtype, x.device, tensor=x.t[b].reshape(1, 1, 1, L, Cn)) if self.residual_connection else None
            ob = Vol(1, 1, 1, L, Cn, x.dtype, x.device, tensor=out.t[b].reshape(1, 1, 1, L, Cn))
            # B == 1 (every tiled call): the out-projection's epilogue also emits the GroupNorm statistics that the next
            # resnet's norm1 needs, so no separate statistics pass reads the result back
            with N.profile_class("attn_proj"):
                yb = _gemm_nt(o, wo, bo, Cn, residual=res, out=ob, gn_groups=self.emit_gn_groups if B == 1 else 0)
            if B == 1 and yb.gn_sums is not None:
                out.gn_sums, out.gn_groups = yb.gn_sums, yb.gn_groups
        return out


class UNetMidBlockCausal3D(nn.Module):
    """unet_causal_3d_blocks.py:525-678: resnet, then (attention, resnet) x num_layers, with the
    optional per-resnet temporal avg-pool hooks of the stride/pool experiments."""

    def __init__(self, in_channels, temb_channels, dropout=0.0, num_layers=1, resnet_eps=1e-6,
                 resnet_time_scale_shift="default", resnet_act_fn="swish", resnet_groups=32, attn_groups=None,
                 resnet_pre_norm=True, add_attention=True, attention_head_dim=1, output_scale_factor=1.0):
        super().__init__()
        self.add_attention = add_attention
        resnet_groups = resnet_groups if resnet_groups is not None else min(in_channels // 4, 32)
        if attn_groups is None:
            attn_groups = resnet_groups if resnet_time_scale_shift == "default" else None
        if attention_head_dim is None:
            attention_head_dim = in_channels

        def res():
            return ResnetBlockCausal3D(in_channels=in_channels, out_channels=in_channels, temb_channels=temb_channels,
                                       eps=resnet_eps, groups=resnet_groups, dropout=dropout,
                                       time_embedding_norm=resnet_time_scale_shift, non_linearity=resnet_act_fn,
                                       output_scale_factor=output_scale_factor, pre_norm=resnet_pre_norm)

        resnets, attentions = [res()], []
        for _ in range(num_layers):
            attentions.append(Attention(in_channels, heads=in_channels // attention_head_dim, dim_head=attention_head_dim,
                                        rescale_output_factor=output_scale_factor, eps=resnet_eps,
                                        norm_num_groups=attn_groups, residual_connection=True, bias=True)
                              if add_attention else None)
            resnets.append(res())
        self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)
        self.num_resblocks = num_layers + 1
        self.resnet_pool_configs = [None] * self.num_resblocks
        self.resnet_pad_configs = [None] * self.num_resblocks

    def apply_t_ops_config_midblock(self, config: dict):
        if not isinstance(config, dict):
            return
        epb = config.get("enable_t_pool_before_block", [])
        epa = config.get("enable_t_pool_after_block", [])
        if any(len(lst) != self.num_resblocks for lst in (epb, epa)):
            raise ValueError(f"[UNetMidBlockCausal3D] T-ops config mismatch: we have {self.num_resblocks} ResnetBlock(s), "
                             f"but got list lengths: {list(map(len, [epb, epa]))}")
        k, s = config.get("pool_t_kernel", 2), config.get("pool_t_stride", 2)
        for i in range(self.num_resblocks):
            self.resnet_pool_configs[i] = {"enable_before": epb[i], "enable_after": epa[i], "kernel": k, "stride": s}

    def forward_vol(self, x: Vol) -> Vol:
        for i in range(self.num_resblocks):
            if i > 0 and self.attentions[i - 1] is not None:
                x = self.attentions[i - 1].forward_vol(x)
            pc = self.resnet_pool_configs[i] or {}
            if pc.get("enable_before", False):
                x = N.avgpool_t(x, pc["kernel"], pc["stride"])
            x = self.resnets[i].forward_vol(x)
            if pc.get("enable_after", False):
                x = N.avgpool_t(x, pc["kernel"], pc["stride"])
        return x

    def forward(self, hidden_states, temb=None):
        return self.forward_vol(Vol.from_ncthw(hidden_states)).to_ncthw()


class DownEncoderBlockCausal3D(nn.Module):
    """unet_causal_3d_blocks.py:680-790."""

    def __init__(self, in_channels, out_channels, dropout=0.0, num_layers=1, resnet_eps=1e-6,
                 resnet_time_scale_shift="default", resnet_act_fn="swish", resnet_groups=32, resnet_pre_norm=True,
                 output_scale_factor=1.0, add_downsample=True, downsample_stride=2, downsample_padding=1):
        super().__init__()
        self.resnets = nn.ModuleList([
            ResnetBlockCausal3D(in_channels=in_channels if i == 0 else out_channels, out_channels=out_channels,
                                temb_channels=None, eps=resnet_eps, groups=resnet_groups, dropout=dropout,
                                time_embedding_norm=resnet_time_scale_shift, non_linearity=resnet_act_fn,
                                output_scale_factor=output_scale_factor, pre_norm=resnet_pre_norm)
            for i in range(num_layers)])
        self.downsamplers = (nn.ModuleList([DownsampleCausal3D(out_channels, use_conv=True, out_channels=out_channels,
                                                               padding=downsample_padding, name="op", stride=downsample_stride)])
                             if add_downsample else None)
        self.resnet_pool_configs = [None] * num_layers

    def apply_t_ops_config(self, block_config: dict):
        if "downsample_stride" in block_config and self.downsamplers is not None:
            for ds in self.downsamplers:
                ds.conv.conv.stride = tuple(block_config["downsample_stride"])
        n = len(self.resnets)
        epb = block_config.get("enable_t_pool_before_block", [])
        epa = block_config.get("enable_t_pool_after_block", [])
        if any(len(x) != n for x in (epb, epa)):
            raise ValueError(f"[DownEncoderBlockCausal3D] config mismatch: expecting {n} bools in each list.")
        k, s = block_config.get("pool_t_kernel", 2), block_config.get("pool_t_stride", 2)
        for i in range(n):
            self.resnet_pool_configs[i] = {"enable_before": epb[i], "enable_after": epa[i], "kernel": k, "stride": s}

    def forward_vol(self, x: Vol) -> Vol:
        for i, resnet in enumerate(self.resnets):
            pc = self.resnet_pool_configs[i] or {}
            if pc.get("enable_before", False):
                x = N.avgpool_t(x, pc["kernel"], pc["stride"])
            last = i == len(self.resnets) - 1 and self.downsamplers is not None and not pc.get("enable_after", False)
            # the last resnet writes straight into the halo'd operand of the downsampler conv (no pad pass in between)
            x = resnet.forward_vol(x, out_pad=self.downsamplers[0].conv.wants_halo(x.dtype) if last else (0, 0, 0))
            if pc.get("enable_after", False):
                x = N.avgpool_t(x, pc["kernel"], pc["stride"])
        if self.downsamplers is not None:
            for ds in self.downsamplers:
                x = ds.forward_vol(x)
        return x

    def forward(self, hidden_states, scale: float = 1.0, index: int = None):
        return self.forward_vol(Vol.from_ncthw(hidden_states)).to_ncthw()


class UpDecoderBlockCausal3D(nn.Module):
    """unet_causal_3d_blocks.py:792-917."""

    def __init__(self, in_channels, out_channels, resolution_idx=None, dropout=0.0, num_layers=1, resnet_eps=1e-6,
                 resnet_time_scale_shift="default", resnet_act_fn="swish", resnet_groups=32, resnet_pre_norm=True,
                 output_scale_factor=1.0, add_upsample=True, upsample_scale_factor=(2, 2, 2), temb_channels=None):
        super().__init__()
        self.resnets = nn.ModuleList([
            ResnetBlockCausal3D(in_channels=in_channels if i == 0 else out_channels, out_channels=out_channels,
                                temb_channels=temb_channels, eps=resnet_eps, groups=resnet_groups, dropout=dropout,
                                time_embedding_norm=resnet_time_scale_shift, non_linearity=resnet_act_fn,
                                output_scale_factor=output_scale_factor, pre_norm=resnet_pre_norm)
            for i in range(num_layers)])
        self.upsamplers = (nn.ModuleList([UpsampleCausal3D(out_channels, use_conv=True, out_channels=out_channels,
                                                           upsample_factor=upsample_scale_factor)])
                           if add_upsample else None)
        self.resolution_idx, self.num_layers = resolution_idx, num_layers
        self.resnet_interp_configs = [None] * num_layers

    def apply_t_ops_config(self, block_config: dict):
        n = len(self.resnets)
        eib = block_config.get("enable_t_interp_before_block", [])
        eia = block_config.get("enable_t_interp_after_block", [])
        if any(len(x) != n for x in (eib, eia)):
            raise ValueError(f"[UpDecoderBlockCausal3D] config mismatch: expecting {n} bools in each list.")
        sc, mode = block_config.get("interp_t_scale_factor", 2), block_config.get("interp_mode", "nearest")
        for i in range(n):
            self.resnet_interp_configs[i] = {"enable_before": eib[i], "enable_after": eia[i], "scale_factor": sc, "mode": mode}

    @staticmethod
    def _interp(x: Vol, conf) -> Vol:
        if x.T <= 0:
            return x
        if conf["mode"] == "nearest":
            return N.interp_t_nearest(x, conf["scale_factor"])
        if conf["mode"] not in N.INTERP_MODES:   # F.interpolate raises for 'linear' / 'bilinear' / 'bicubic' on a 5-D tensor as well
            raise NotImplementedError(f"interp_mode {conf['mode']!r} is not defined for 5-D tensors")
        return N.interp_t(x, conf["scale_factor"], conf["mode"])

    def forward_vol(self, x: Vol) -> Vol:
        for i, resnet in enumerate(self.resnets):
            ic = self.resnet_interp_configs[i] or {}
            if ic.get("enable_before", False):
                x = self._interp(x, ic)
            last = i == len(self.resnets) - 1 and self.upsamplers is not None and not ic.get("enable_after", False)
            x = resnet.forward_vol(x, out_pad=self.upsamplers[0].wants_halo(x.dtype) if last else (0, 0, 0))
            if ic.get("enable_after", False):
                x = self._interp(x, ic)
        if self.upsamplers is not None:
            for up in self.upsamplers:
                x = up.forward_vol(x)
        return x

    def forward(self, hidden_states, temb=None, scale: float = 1.0):
        return self.forward_vol(Vol.from_ncthw(hidden_states)).to_ncthw()


def get_down_block3d(down_block_type, num_layers, in_channels, out_channels, temb_channels, add_downsample,
                     downsample_stride, resnet_eps, resnet_act_fn, resnet_groups=None, downsample_padding=None,
                     resnet_time_scale_shift="default", dropout=0.0, **unused):
    t = down_block_type[7:] if down_block_type.startswith("UNetRes") else down_block_type
    if t == "DownEncoderBlockCausal3D":
        return DownEncoderBlockCausal3D(num_layers=num_layers, in_channels=in_channels, out_channels=out_channels,
                                        dropout=dropout, add_downsample=add_downsample, downsample_stride=downsample_stride,
                                        resnet_eps=resnet_eps, resnet_act_fn=resnet_act_fn, resnet_groups=resnet_groups,
                                        downsample_padding=downsample_padding, resnet_time_scale_shift=resnet_time_scale_shift)
    raise ValueError(f"{down_block_type} does not exist.")


def get_up_block3d(up_block_type, num_layers, in_channels, out_channels, prev_output_channel, temb_channels,
                   add_upsample, upsample_scale_factor, resnet_eps, resnet_act_fn, resolution_idx=None,
                   resnet_groups=None, resnet_time_scale_shift="default", dropout=0.0, **unused):
    t = up_block_type[7:] if up_block_type.startswith("UNetRes") else up_block_type
    if t == "UpDecoderBlockCausal3D":
        return UpDecoderBlockCausal3D(num_layers=num_layers, in_channels=in_channels, out_channels=out_channels,
                                      resolution_idx=resolution_idx, dropout=dropout, add_upsample=add_upsample,
                                      upsample_scale_factor=upsample_scale_factor, resnet_eps=resnet_eps,
                                      resnet_act_fn=resnet_act_fn, resnet_groups=resnet_groups,
                                      resnet_time_scale_shift=resnet_time_scale_shift, temb_channels=temb_channels)
    raise ValueError(f"{up_block_type} does not exist.")
