"""ctypes binding of libhyvae.so (include/hyvae.h) and the channels-last volume container.

There is NO CPU fallback: every op below enqueues a hand-written sm_100a kernel on the current CUDA
stream, and loading fails loudly when the shared library has not been built
(`make -C hunyuanvideo_efficiency_b200/csrc`, or `python -c "import __graft_entry__ as g; g.build()"`).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhyvae.so")

BF16, F32, F16 = 0, 1, 2
_DT = {torch.bfloat16: BF16, torch.float32: F32, torch.float16: F16}


class HyvaeError(RuntimeError):
    pass


class HyvaeUnsupported(HyvaeError):
    """HYVAE_EUNSUPPORTED (-3): the entry point does not take this shape; the caller picks another schedule."""


class _CVol(C.Structure):
    _fields_ = [("data", C.c_void_p), ("dtype", C.c_int32), ("B", C.c_int32), ("T", C.c_int32), ("H", C.c_int32),
                ("W", C.c_int32), ("C", C.c_int32), ("pt", C.c_int32), ("ph", C.c_int32), ("pw", C.c_int32)]


_VP = C.POINTER(_CVol)
_i32, _i64, _f32, _vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p

# name -> argtypes; every entry returns int status (include/hyvae.h)
_SIGNATURES = {
    "hyvae_ncthw_to_vol": [_vp, _i32, _i32, C.POINTER(_i64), _VP, _vp],
    "hyvae_ncthw_to_vol_kw3": [_vp, _i32, _i32, C.POINTER(_i64), _VP, _vp],
    "hyvae_vol_to_ncthw": [_VP, _vp, _i32, _i32, _vp],
    "hyvae_conv3d_causal_direct": [_VP, _vp, _vp, _VP, _VP, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp],
    "hyvae_conv3d_causal_tc": [_VP, _vp, _vp, _VP, _VP, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp],
    "hyvae_conv3d_causal_tc_shortcut": [_VP, _vp, _vp, _VP, _vp, _VP, _vp, _i32, _i32, _vp],
    "hyvae_conv3d_upphase_tc": [_VP, _vp, _vp, _VP, _i32, _i32, _i32, _i32, _vp, _i32, _vp],
    "hyvae_groupnorm_finalize": [_vp, _i32, _i64, _i32, _vp, _vp],
    "hyvae_groupnorm_apply_wino": [_VP, _vp, _vp, _vp, _i32, _f32, _i32, _VP, _vp],
    "hyvae_conv3d_causal_wino": [_VP, _i32, _vp, _vp, _VP, _VP, _vp, _VP, _vp, _i32, _vp, _vp],
    "hyvae_groupnorm_stats": [_VP, _i32, _vp, _vp, _i64, _vp],
    "hyvae_groupnorm_apply": [_VP, _vp, _vp, _vp, _i32, _f32, _i32, _i32, _VP, _vp],
    "hyvae_pad_upsample": [_VP, _VP, _i32, _i32, _i32, _vp],
    "hyvae_halo_fill": [_VP, _vp],
    "hyvae_softmax_frame_causal": [_vp, _vp, _i32, _i32, _i32, _i32, _f32, _vp],
    "hyvae_attn_block_causal": [_vp, _vp, _vp, _vp, _vp, _i32, _i64, _i32, _i32, _f32, _vp],
    "hyvae_video_to_frames_u8": [_vp, _i32, C.POINTER(_i64), _i32, _i32, _i32, _i32, _i32, _vp, _vp],
    "hyvae_frame_metrics_u8": [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _i64, _vp],
    "hyvae_avgpool_t": [_VP, _VP, _i32, _i32, _vp],
    "hyvae_interp_t_nearest": [_VP, _VP, _f32, _vp],
    "hyvae_interp_t": [_VP, _VP, _i32, _f32, _vp],
    "hyvae_image_postprocess": [_vp, _i32, _vp, _i64, _vp],
    "hyvae_blend_crop_scatter": [_vp, _vp, _vp, _i32, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _i32,
                                 _i32, _i32, _i32, _i32, C.POINTER(_i64), _i32, _vp],
    "hyvae_peer_copy": [_vp, _vp, _i64, _vp],
    "hyvae_profile_class_override": [_i32],
}
EXPORTS = sorted(list(_SIGNATURES) + ["hyvae_version", "hyvae_last_error", "hyvae_device_supports_tc", "hyvae_launch_count",
                                       "hyvae_groupnorm_workspace_bytes", "hyvae_profile_begin", "hyvae_profile_end", "hyvae_profile_executed_flops",
                                       "hyvae_conv3d_tc_gn_rows", "hyvae_gn_partials_doubles", "hyvae_frame_metrics_workspace_bytes", "hyvae_wino_planes"])

_lib = None


def lib():
    """Load libhyvae.so once.  Raises HyvaeError if it is missing — the product has no other path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HyvaeError(f"{LIB_PATH} not found: build the CUDA extension first "
                             f"(make -C {os.path.join(_HERE, 'csrc')}); there is no CPU fallback")
        l = C.CDLL(LIB_PATH)
        for name, args in _SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = args
            fn.restype = C.c_int
        l.hyvae_version.restype = C.c_int
        if l.hyvae_version() != ABI_VERSION:   # a stale build would be called with the wrong argument lists
            raise HyvaeError(f"{LIB_PATH} is ABI version {l.hyvae_version()}, this package binds {ABI_VERSION}: rebuild it "
                             f"(make -C {os.path.join(_HERE, 'csrc')})")
        l.hyvae_last_error.restype = C.c_char_p
        l.hyvae_device_supports_tc.restype = C.c_int
        l.hyvae_launch_count.restype = C.c_int64
        l.hyvae_groupnorm_workspace_bytes.restype = C.c_int64
        l.hyvae_frame_metrics_workspace_bytes.restype = C.c_int64
        l.hyvae_frame_metrics_workspace_bytes.argtypes = [_i32, _i32, _i32]
        l.hyvae_groupnorm_workspace_bytes.argtypes = [_VP, _i32]
        l.hyvae_profile_executed_flops.restype = C.c_double
        l.hyvae_profile_executed_flops.argtypes = []
        l.hyvae_wino_planes.restype = C.c_int32
        l.hyvae_wino_planes.argtypes = [_i32]
        l.hyvae_conv3d_tc_gn_rows.restype = C.c_int64
        l.hyvae_conv3d_tc_gn_rows.argtypes = []
        l.hyvae_gn_partials_doubles.restype = C.c_int64
        l.hyvae_gn_partials_doubles.argtypes = [_i32, _i32]
        _lib = l
    return _lib


def _check(status: int, what: str):
    if status != 0:
        cls = HyvaeUnsupported if status == -3 else HyvaeError
        raise cls(f"{what} failed ({status}): {lib().hyvae_last_error().decode()}")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


PROFILE_CLASSES = ("conv_tc", "conv_direct", "gn_stats", "gn_apply", "pad_upsample", "softmax", "layout", "blend", "temporal", "attn", "attn_proj")


def profile_begin():
    _check(lib().hyvae_profile_begin(), "profile_begin")


def profile_end() -> dict:
    """{class: {"ms", "work" (flops or bytes, algorithmic), "launches"}} for the bracketed region."""
    n = len(PROFILE_CLASSES)
    ms, work, cnt = (C.c_double * n)(), (C.c_double * n)(), (C.c_int64 * n)()
    fn = lib().hyvae_profile_end
    fn.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int32]
    _check(fn(ms, work, cnt, n), "profile_end")
    out = {k: {"ms": ms[i], "work": work[i], "launches": int(cnt[i])} for i, k in enumerate(PROFILE_CLASSES)}
    out["conv_tc"]["executed"] = float(lib().hyvae_profile_executed_flops())
    return out


class profile_class:
    """with profile_class("attn_proj"): conv launches inside are booked under that profile class (measurement only)."""

    def __init__(self, name: str):
        self.idx = PROFILE_CLASSES.index(name)

    def __enter__(self):
        lib().hyvae_profile_class_override(self.idx)

    def __exit__(self, *exc):
        lib().hyvae_profile_class_override(-1)
        return False


def launch_count() -> int:
    return int(lib().hyvae_launch_count())


_tc_ok: Optional[bool] = None


def device_supports_tc() -> bool:
    global _tc_ok
    if _tc_ok is None:
        _tc_ok = bool(lib().hyvae_device_supports_tc())
    return _tc_ok


class Vol:
    """Channels-last activation volume [B][T+pt][H+2ph][W+2pw][C] in HBM (see hyvae_vol)."""

    __slots__ = ("t", "B", "T", "H", "W", "C", "pad", "_c", "c_valid", "gn_sums", "gn_groups", "kw_packed", "wino_T")

    def __init__(self, B, T, H, W, Cn, dtype, device, pad: Tuple[int, int, int] = (0, 0, 0), tensor=None):
        pt, ph, pw = pad
        shape = (B, T + pt, H + 2 * ph, W + 2 * pw, Cn)
        if tensor is None:
            tensor = torch.empty(shape, dtype=dtype, device=device)
        else:
            assert tuple(tensor.shape) == shape and tensor.is_contiguous()
        if tensor.device.type != "cuda":
            raise HyvaeError("hyvae volumes live in CUDA memory; there is no CPU execution path")
        self.t, self.B, self.T, self.H, self.W, self.C, self.pad = tensor, B, T, H, W, Cn, tuple(pad)
        self._c = _CVol(tensor.data_ptr(), _DT[tensor.dtype], B, T, H, W, Cn, pt, ph, pw)
        self.c_valid = Cn        # channels that carry data (C may be zero-padded up to a multiple of 8)
        self.gn_sums = None      # [B][groups][2] fp64 GroupNorm statistics emitted by the producing conv, if any
        self.gn_groups = 0
        self.wino_T = 0          # > 0: this is the Winograd-T PLANE volume of a tensor with that many frames (groupnorm_wino)
        self.kw_packed = False   # channels are (kw, c) of a thin source (from_ncthw(kw_pack=True)): only conv_in reads such a volume

    @property
    def dtype(self):
        return self.t.dtype

    @property
    def device(self):
        return self.t.device

    @property
    def dims(self):
        return (self.B, self.T, self.H, self.W, self.C)

    def ref(self):
        return C.byref(self._c)

    def like(self, C_=None, pad=(0, 0, 0), dtype=None, T=None, H=None, W=None) -> "Vol":
        return Vol(self.B, T or self.T, H or self.H, W or self.W, C_ or self.C, dtype or self.dtype, self.device, pad)

    def interior(self) -> torch.Tensor:
        """Logical [B,T,H,W,C] view (no halo)."""
        pt, ph, pw = self.pad
        return self.t[:, pt:, ph:ph + self.H, pw:pw + self.W, :]

    # ---- NCTHW <-> volume -------------------------------------------------------------------
    @staticmethod
    def from_ncthw(x: torch.Tensor, dtype=None, pad=(0, 0, 0), channels: Optional[int] = None, kw_pack: bool = False) -> "Vol":
        """`channels` > x.shape[1] zero-pads the channel axis (3 -> 8 for the tensor-core conv_in).  kw_pack: channel
        kw * C + c of a voxel holds source channel c of its (w + kw - 1) neighbour (hyvae_ncthw_to_vol_kw3)."""
        assert x.ndim == 5
        B, Cn, T, H, W = x.shape
        v = Vol(B, T, H, W, channels or Cn, dtype or x.dtype, x.device, pad)
        v.c_valid = Cn
        strides = (_i64 * 5)(*x.stride())
        if kw_pack:
            _check(lib().hyvae_ncthw_to_vol_kw3(x.data_ptr(), _DT[x.dtype], Cn, strides, v.ref(), _stream()), "ncthw_to_vol_kw3")
            v.kw_packed = True
        else:
            _check(lib().hyvae_ncthw_to_vol(x.data_ptr(), _DT[x.dtype], Cn, strides, v.ref(), _stream()), "ncthw_to_vol")
        return v

    def to_ncthw(self, dtype=None) -> torch.Tensor:
        out = torch.empty((self.B, self.c_valid, self.T, self.H, self.W), dtype=dtype or self.dtype, device=self.device)
        _check(lib().hyvae_vol_to_ncthw(self.ref(), out.data_ptr(), _DT[out.dtype], self.c_valid, _stream()), "vol_to_ncthw")
        return out


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


# ------------------------------------------------------------------------------------------------ ops
def conv_out_dims(T, H, W, stride, up=(1, 1, 1)):
    Tl = 1 + 2 * (T - 1) if up[0] == 2 else T
    Hl, Wl = H * up[1], W * up[2]
    return (Tl - 1) // stride[0] + 1, (Hl - 1) // stride[1] + 1, (Wl - 1) // stride[2] + 1


def conv3d_direct(x: Vol, w: torch.Tensor, bias, k: int, stride, cout: int, residual: Optional[Vol] = None,
                  up=(1, 1, 1), out_dtype=None, round_like_ref=True, out: Optional[Vol] = None) -> Vol:
    To, Ho, Wo = conv_out_dims(x.T, x.H, x.W, stride, up)
    y = out if out is not None else Vol(x.B, To, Ho, Wo, cout, out_dtype or x.dtype, x.device)
    _check(lib().hyvae_conv3d_causal_direct(x.ref(), w.data_ptr(), _ptr(bias), residual.ref() if residual else None, y.ref(),
                                            k, stride[0], stride[1], stride[2], up[0], up[1], up[2],
                                            int(round_like_ref), _stream()), "conv3d_causal_direct")
    return y


ABI_VERSION = 121      # HYVAE_VERSION of include/hyvae.h this module binds
VARIANT_KWPACK = 0x200  # hyvae_conv3d_causal_tc: x is a kw-packed thin volume (Vol.from_ncthw(kw_pack=True)), w is [9][Cout][16]
VARIANT_TFOLD = 0x100  # hyvae_conv3d_causal_tc: `w` carries the 18 folded first-frame tap slices after the 27 (include/hyvae.h)
_GN_PART = {}


def _gn_partials(B: int, rows: int, groups: int, device) -> torch.Tensor:
    """fp64 scratch for the conv epilogue's GroupNorm partial sums: [B][rows][groups][2] warp rows, then the per-CTA rows and
    the ticket of the convs that finish the statistics themselves (hyvae_gn_partials_doubles).  Zeroed once: every
    hyvae_groupnorm_finalize / fused finalize leaves it zeroed again, and conv + finalize pairs are stream ordered."""
    key = (device, torch.cuda.current_stream(device).cuda_stream, B, rows, groups)
    buf = _GN_PART.get(key)
    if buf is None:
        n = int(lib().hyvae_gn_partials_doubles(B, groups))
        assert n >= B * rows * groups * 2
        buf = _GN_PART[key] = torch.zeros((n,), dtype=torch.float64, device=device)
    return buf


class _GnEpilogue:
    """GroupNorm statistics of a conv's output from its epilogue.  `sums` is allocated BEFORE any launch, and a failure
    between the first conv launch and hyvae_groupnorm_finalize (which re-zeroes the partial buffer) drops the cached
    buffer, so a failed call (e.g. an out-of-memory error the caller catches before retrying with tiling) cannot leave
    partial sums behind that would poison every later statistic on that stream."""

    def __init__(self, x: Vol, cout: int, gn_groups: int, min_cpg: int):
        self.part = self.sums = None
        self.rows, self.groups, self.B, self.device = 0, 0, x.B, x.device
        if gn_groups > 0 and cout % gn_groups == 0 and (cout // gn_groups) in (1, 2, 4, 8, 16, 32) and cout // gn_groups >= min_cpg:
            self.sums = torch.empty((x.B, gn_groups, 2), dtype=torch.float64, device=x.device)
            self.rows = int(lib().hyvae_conv3d_tc_gn_rows())
            self.groups = gn_groups
            self.part = _gn_partials(x.B, self.rows, gn_groups, x.device)

    def args(self):
        return _ptr(self.part), self.groups

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc, tb):
        if exc_type is not None and self.part is not None:
            _GN_PART.pop((self.device, torch.cuda.current_stream(self.device).cuda_stream, self.B, self.rows, self.groups), None)
        return False

    def attach(self, y: Vol):
        """The conv finished the statistics itself (fused finalize): they travel with y, no launch."""
        if self.part is not None:
            y.gn_sums, y.gn_groups = self.sums, self.groups
        return y

    def finalize(self, y: Vol):
        if self.part is not None:
            _check(lib().hyvae_groupnorm_finalize(self.part.data_ptr(), self.B, self.rows, self.groups, self.sums.data_ptr(), _stream()),
                   "groupnorm_finalize")
            y.gn_sums, y.gn_groups = self.sums, self.groups
        return y


def conv3d_tc(x: Vol, w: torch.Tensor, bias, k: int, stride, cout: int, residual: Optional[Vol] = None,
              out_dtype=None, round_like_ref=True, variant=0, out: Optional[Vol] = None, gn_groups: int = 0) -> Vol:
    """gn_groups > 0: the epilogue also emits GroupNorm partial statistics of y; they are reduced here
    (hyvae_groupnorm_finalize) and travel with the returned volume (y.gn_sums), sparing the consumer's stats pass."""
    To, Ho, Wo = conv_out_dims(x.T, x.H, x.W, stride)
    y = out if out is not None else Vol(x.B, To, Ho, Wo, cout, out_dtype or x.dtype, x.device)
    with _GnEpilogue(x, cout, gn_groups, 1) as gn:
        part, groups = gn.args()
        _check(lib().hyvae_conv3d_causal_tc(x.ref(), w.data_ptr(), _ptr(bias), residual.ref() if residual else None, y.ref(),
                                            k, stride[0], stride[1], stride[2], int(round_like_ref), variant, part, groups, _stream()),
               "conv3d_causal_tc")
        return gn.finalize(y)


def conv3d_tc_shortcut(x: Vol, w: torch.Tensor, bias: torch.Tensor, sc_x: Vol, sc_w: torch.Tensor, cout: int, gn_groups: int = 0,
                       tfold: bool = False, out_pad=(0, 0, 0)) -> Vol:
    """y = conv3x3x3(x) + conv1x1x1(sc_x) + bias in one launch (the resnet block's conv2 with its conv_shortcut).
    out_pad: y is allocated with that halo and only its interior is written (the caller runs halo_fill)."""
    y = Vol(x.B, x.T, x.H, x.W, cout, x.dtype, x.device, out_pad)
    with _GnEpilogue(x, cout, gn_groups, 2) as gn:
        part, groups = gn.args()
        _check(lib().hyvae_conv3d_causal_tc_shortcut(x.ref(), w.data_ptr(), _ptr(bias), sc_x.ref(), sc_w.data_ptr(), y.ref(),
                                                     part, groups, int(tfold), _stream()), "conv3d_causal_tc_shortcut")
        return gn.finalize(y)


def conv3d_upsample_phases(x: Vol, phase_w, bias, up, cout: int, gn_groups: int = 0) -> Vol:
    """UpsampleCausal3D's nearest upsample + 3x3x3 conv as 4 (up[0] == 1) or 8 phase convolutions over the LOW-res
    volume `x`, which must carry the halo (nkt-1, 1, 1).  phase_w: {(pt, ph, pw): [nkt*4][Cout][Cin] tensor}."""
    assert up[1] == 2 and up[2] == 2 and up[0] in (1, 2)
    T = 2 * x.T - 1 if up[0] == 2 else x.T
    y = Vol(x.B, T, 2 * x.H, 2 * x.W, cout, x.dtype, x.device)
    with _GnEpilogue(x, cout, gn_groups, 2) as gn:
        part, groups = gn.args()
        for (pt, ph, pw), w in phase_w.items():
            _check(lib().hyvae_conv3d_upphase_tc(x.ref(), w.data_ptr(), _ptr(bias), y.ref(), up[0], pt, ph, pw, part, groups, _stream()),
                   "conv3d_upphase_tc")
        return gn.finalize(y)


def groupnorm(x: Vol, gamma: torch.Tensor, beta: torch.Tensor, groups: int, eps: float, silu: bool,
              pad=(0, 0, 0), round_like_ref=True) -> Vol:
    if x.gn_sums is not None and x.gn_groups == groups:
        sums = x.gn_sums   # statistics came with the tensor, from the conv epilogue that produced it
    else:
        sums = torch.empty((x.B, groups, 2), dtype=torch.float64, device=x.device)
        nbytes = lib().hyvae_groupnorm_workspace_bytes(x.ref(), groups)
        if nbytes < 0:
            raise HyvaeError(f"GroupNorm: unsupported channel count C={x.C} (needs C % 8 == 0)")
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=x.device)
        _check(lib().hyvae_groupnorm_stats(x.ref(), groups, sums.data_ptr(), ws.data_ptr(), nbytes, _stream()), "groupnorm_stats")
    y = x.like(pad=pad)
    _check(lib().hyvae_groupnorm_apply(x.ref(), sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(), groups, eps, int(silu),
                                       int(round_like_ref), y.ref(), _stream()), "groupnorm_apply")
    return y


def wino_planes(T: int) -> int:
    """Planes of the Winograd-T operand of a T-frame tensor (hyvae_wino_planes)."""
    return 0 if T <= 0 else 1 + 4 * ((T - 1) // 2) + (3 if T % 2 == 0 else 0)


def groupnorm_wino(x: Vol, gamma: torch.Tensor, beta: torch.Tensor, groups: int, eps: float, silu: bool) -> Vol:
    """GroupNorm (+SiLU) written as the Winograd-T plane volume of the following stride-1 3x3x3 conv
    (hyvae_groupnorm_apply_wino): [B][wino_planes(T)][H+2][W+2][C]; statistics as in groupnorm()."""
    if x.gn_sums is not None and x.gn_groups == groups:
        sums = x.gn_sums
    else:
        sums = torch.empty((x.B, groups, 2), dtype=torch.float64, device=x.device)
        nbytes = lib().hyvae_groupnorm_workspace_bytes(x.ref(), groups)
        if nbytes < 0:
            raise HyvaeError(f"GroupNorm: unsupported channel count C={x.C} (needs C % 8 == 0)")
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=x.device)
        _check(lib().hyvae_groupnorm_stats(x.ref(), groups, sums.data_ptr(), ws.data_ptr(), nbytes, _stream()), "groupnorm_stats")
    y = Vol(x.B, wino_planes(x.T), x.H, x.W, x.C, x.dtype, x.device, (0, 1, 1))
    y.wino_T = x.T
    _check(lib().hyvae_groupnorm_apply_wino(x.ref(), sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(), groups, eps, int(silu),
                                            y.ref(), _stream()), "groupnorm_apply_wino")
    return y


def conv3d_wino(planes: Vol, uw: torch.Tensor, bias, cout: int, residual: Optional[Vol] = None, out_pad=(0, 0, 0),
                gn_groups: int = 0, sc_x: Optional[Vol] = None, sc_w: Optional[torch.Tensor] = None) -> Vol:
    """Stride-1 3x3x3 causal conv of the tensor whose Winograd-T planes are `planes` (hyvae_conv3d_causal_wino);
    uw: [45][Cout][Cin] = the five tap groups of _Conv3dParams.wino_packed.  sc_x / sc_w ([2][Cout][Csc] = +Ws, -Ws): the
    resnet block's 1x1x1 conv_shortcut fused into the accumulators (`bias` is then the sum of both convs' biases)."""
    T = planes.wino_T
    assert T > 0, "conv3d_wino takes the plane volume written by groupnorm_wino"
    y = Vol(planes.B, T, planes.H, planes.W, cout, planes.dtype, planes.device, tuple(out_pad))
    with _GnEpilogue(planes, cout, gn_groups, 4) as gn:
        part, groups = gn.args()
        _check(lib().hyvae_conv3d_causal_wino(planes.ref(), T, uw.data_ptr(), _ptr(bias), residual.ref() if residual else None,
                                              sc_x.ref() if sc_x is not None else None, _ptr(sc_w), y.ref(), part, groups,
                                              _ptr(gn.sums) if part is not None else None, _stream()),
               "conv3d_causal_wino")
        return gn.attach(y)   # the kernel's last CTA summed the rows: no hyvae_groupnorm_finalize launch


def halo_fill(y: Vol) -> Vol:
    """Replicate halo of a padded volume whose interior a conv has just written (hyvae_halo_fill)."""
    if y.pad != (0, 0, 0):
        _check(lib().hyvae_halo_fill(y.ref(), _stream()), "halo_fill")
    return y


def pad_upsample(x: Vol, up=(1, 1, 1), pad=(0, 0, 0), channels: Optional[int] = None) -> Vol:
    T = 1 + 2 * (x.T - 1) if up[0] == 2 else x.T
    y = Vol(x.B, T, x.H * up[1], x.W * up[2], channels or x.C, x.dtype, x.device, pad)
    y.c_valid = x.c_valid
    _check(lib().hyvae_pad_upsample(x.ref(), y.ref(), up[0], up[1], up[2], _stream()), "pad_upsample")
    return y


def softmax_frame_causal(S: torch.Tensor, n_hw: int, scale: float, p_dtype) -> torch.Tensor:
    B, L, L2 = S.shape
    assert L == L2 and S.dtype == torch.float32 and S.is_contiguous()
    P = torch.empty((B, L, L), dtype=p_dtype, device=S.device)
    _check(lib().hyvae_softmax_frame_causal(S.data_ptr(), P.data_ptr(), _DT[p_dtype], B, L, n_hw, scale, _stream()),
           "softmax_frame_causal")
    return P


def attn_block_causal(q: torch.Tensor, k: torch.Tensor, vt: torch.Tensor, bv, n_hw: int, scale: float) -> torch.Tensor:
    """Fused softmax(scale Q K^T + frame-causal mask) V + bv (hyvae_attn_block_causal).  q, k: [L][D]; vt: [D][L].
    Raises HyvaeUnsupported for shapes the fused kernel does not take."""
    L, D = q.shape
    assert k.shape == (L, D) and vt.shape == (D, L) and q.dtype == k.dtype == vt.dtype
    assert q.is_contiguous() and k.is_contiguous() and vt.is_contiguous()
    o = torch.empty((L, D), dtype=q.dtype, device=q.device)
    _check(lib().hyvae_attn_block_causal(q.data_ptr(), k.data_ptr(), vt.data_ptr(), _ptr(bv), o.data_ptr(), _DT[q.dtype],
                                         L, n_hw, D, scale, _stream()), "attn_block_causal")
    return o


def video_to_frames_u8(video: torch.Tensor, rescale: bool = True) -> torch.Tensor:
    """(C, T, H, W) or (1, C, T, H, W) float video -> [T][H][W][C] uint8 frames, quantised like save_videos_grid."""
    if video.ndim == 5:
        assert video.shape[0] == 1, "one video per call (save_videos_grid tiles a batch into a grid; not needed here)"
        video = video[0]
    Cn, T, H, W = video.shape
    out = torch.empty((T, H, W, Cn), dtype=torch.uint8, device=video.device)
    st = (_i64 * 4)(*video.stride())
    _check(lib().hyvae_video_to_frames_u8(video.data_ptr(), _DT[video.dtype], st, Cn, T, H, W, int(rescale), out.data_ptr(), _stream()),
           "video_to_frames_u8")
    return out


def frame_metrics_u8(a: torch.Tensor, b: torch.Tensor):
    """Per-frame (ssd int64 [N], min_a, max_a, min_b, max_b int32 [N], ssim float64 [N]) of two [N][H][W][C] uint8 frame stacks."""
    assert a.shape == b.shape and a.dtype == b.dtype == torch.uint8 and a.is_contiguous() and b.is_contiguous()
    n, H, W, Cn = a.shape
    nb = int(lib().hyvae_frame_metrics_workspace_bytes(n, H, W))
    if nb < 0:
        raise HyvaeError(f"frame_metrics_u8: frames of {H}x{W} are smaller than the 7x7 SSIM window")
    ws = torch.empty((nb + 32 * n,), dtype=torch.uint8, device=a.device)
    raw = torch.empty((n, 8), dtype=torch.int32, device=a.device)
    ssim = torch.empty((n,), dtype=torch.float64, device=a.device)
    _check(lib().hyvae_frame_metrics_u8(a.data_ptr(), b.data_ptr(), n, H, W, Cn, raw.data_ptr(), ssim.data_ptr(), ws.data_ptr(),
                                        ws.numel(), _stream()), "frame_metrics_u8")
    ssd = raw.view(torch.int64)[:, 0]
    return ssd, raw[:, 2], raw[:, 3], raw[:, 4], raw[:, 5], ssim


def avgpool_t(x: Vol, k: int, s: int) -> Vol:
    y = x.like(T=(x.T - 1) // s + 1)
    _check(lib().hyvae_avgpool_t(x.ref(), y.ref(), k, s, _stream()), "avgpool_t")
    return y


def interp_t_nearest(x: Vol, scale: float) -> Vol:
    import math
    y = x.like(T=int(math.floor(x.T * scale)))
    _check(lib().hyvae_interp_t_nearest(x.ref(), y.ref(), float(1.0 / scale), _stream()), "interp_t_nearest")
    return y


INTERP_MODES = {"nearest": 0, "trilinear": 1, "area": 2, "nearest-exact": 3}


def interp_t(x: Vol, scale: float, mode: str) -> Vol:
    """F.interpolate(x, scale_factor=(scale, 1, 1), mode=mode) for the modes a 5-D tensor accepts (hyvae_interp_t)."""
    import math
    if mode not in INTERP_MODES:
        raise HyvaeError(f"interp_mode {mode!r}: F.interpolate takes {sorted(INTERP_MODES)} for a 5-D tensor")
    y = x.like(T=int(math.floor(x.T * scale)))
    _check(lib().hyvae_interp_t(x.ref(), y.ref(), INTERP_MODES[mode], float(1.0 / scale), _stream()), "interp_t")
    return y


def blend_crop_scatter(cur: torch.Tensor, above, left, N: int, Yc: int, Xc: int, Ya: int, Xl: int, ev: int, eh: int,
                       out, Yo: int, Xo: int, y0: int, x0: int, crop_y: int, crop_x: int, n_strides=None, post: bool = False):
    """post: `out` is an fp32 tensor that receives float((v / 2 + 0.5).clamp(0, 1)) (the pipeline tail's image)."""
    ns = (_i64 * 4)(*n_strides) if n_strides is not None else None
    if post and (out is None or out.dtype != torch.float32):
        raise HyvaeError("blend_crop_scatter(post=True) writes an fp32 image")
    if not post and out is not None and out.dtype != cur.dtype:
        raise HyvaeError("blend_crop_scatter: out must have the tile dtype")
    _check(lib().hyvae_blend_crop_scatter(cur.data_ptr(), _ptr(above), _ptr(left), _DT[cur.dtype], N, Yc, Xc, Ya, Xl, ev, eh,
                                          _ptr(out), Yo, Xo, y0, x0, crop_y, crop_x, ns, int(post), _stream()), "blend_crop_scatter")


def peer_copy(dst: torch.Tensor, src: torch.Tensor):
    """dst (possibly peer-GPU memory mapped into this process) <- src on the CURRENT stream of src's device (hyvae_peer_copy)."""
    assert dst.is_contiguous() and src.is_contiguous() and dst.dtype == src.dtype and dst.numel() == src.numel()
    _check(lib().hyvae_peer_copy(dst.data_ptr(), src.data_ptr(), src.numel() * src.element_size(),
                                 torch.cuda.current_stream(src.device).cuda_stream), "peer_copy")


def image_postprocess(image: torch.Tensor) -> torch.Tensor:
    """float((image / 2 + 0.5).clamp(0, 1)) in one pass (pipeline_hunyuan_video.py:1090-1092); returns fp32 on the device."""
    if not image.is_cuda:
        raise HyvaeError("image_postprocess runs on CUDA tensors; there is no CPU execution path")
    image = image.contiguous()
    out = torch.empty(image.shape, dtype=torch.float32, device=image.device)
    _check(lib().hyvae_image_postprocess(image.data_ptr(), _DT[image.dtype], out.data_ptr(), image.numel(), _stream()), "image_postprocess")
    return out
