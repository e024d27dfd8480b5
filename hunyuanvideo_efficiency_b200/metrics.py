"""PSNR / SSIM of a reconstruction against its original on the GPU — the evaluation step of the fork's stride / pool /
bucket sweeps (run_experiments_{pool,stride}.sh -> evaluation/compute_metrics.py) without the mp4 round trip.

The reference writes both videos to mp4 with `save_videos_grid(rescale=True)` (hyvideo/utils/file_utils.py:58-66), reads
the uint8 frames back and averages `compute_psnr` / `compute_ssim` (compute_metrics.py:31-41) over all frames of all
videos (:129-152).  Here the same quantisation and the same per-frame metrics run as CUDA kernels
(csrc/metrics.cu) on the tensors the VAE already holds in HBM; the lossy codec in between is the only step left out.
LPIPS (an AlexNet feature distance, compute_metrics.py:43-61) is outside the VAE hot path and not provided.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Tuple

import torch

from . import _native as N


def video_to_frames_u8(video: torch.Tensor, rescale: bool = True) -> torch.Tensor:
    """(C, T, H, W) or (1, C, T, H, W) float video on the device -> [T][H][W][C] uint8 frames (save_videos_grid's rule:
    x = (x + 1) / 2 if rescale; clamp(0, 1); (x * 255) truncated)."""
    if not video.is_cuda:
        raise N.HyvaeError("metrics run on the GPU: the video must be a CUDA tensor (there is no CPU path)")
    return N.video_to_frames_u8(video, rescale)


def frame_metrics(frames1: torch.Tensor, frames2: torch.Tensor) -> Tuple[List[float], List[float]]:
    """Per-frame (PSNR, SSIM) lists of two uint8 frame stacks [T][H][W][C]; like `zip(vid1_frames, vid2_frames)`
    (compute_metrics.py:129) the shorter stack decides how many frames are compared."""
    n = min(frames1.shape[0], frames2.shape[0])
    if n == 0:
        return [], []
    a, b = frames1[:n].contiguous(), frames2[:n].contiguous()
    ssd, mn_a, mx_a, mn_b, mx_b, ssim = (t.cpu() for t in N.frame_metrics_u8(a, b))
    per_frame = a.shape[1] * a.shape[2] * a.shape[3]
    psnr, ssims = [], []
    for i in range(n):
        mse = (float(ssd[i]) / per_frame) / (255.0 * 255.0)   # np.mean((img1/255 - img2/255)**2), :32
        psnr.append(100 if mse < 1.0e-10 else 20 * math.log10(1 / math.sqrt(mse)))   # :33-36
        const = bool(mn_a[i] == mx_a[i]) or bool(mn_b[i] == mx_b[i])                  # :39-40
        ssims.append(1.0 if const else float(ssim[i]))
    return psnr, ssims


def compare_videos(pairs: Iterable[Tuple[torch.Tensor, torch.Tensor]], rescale: bool = True) -> Dict[str, float]:
    """`pairs` of (original, reconstruction) videos, each (C, T, H, W) / (1, C, T, H, W) in [-1, 1] (rescale=True) or
    [0, 1] on the device.  Returns {"PSNR", "SSIM"} as compute_metrics.py's main loop reports them, plus the frame count."""
    ps: List[float] = []
    ss: List[float] = []
    for ref, rec in pairs:
        p, s = frame_metrics(video_to_frames_u8(ref, rescale), video_to_frames_u8(rec, rescale))
        ps += p
        ss += s
    out: Dict[str, float] = {"frames": len(ps)}
    if ps:
        out["PSNR"] = sum(ps) / len(ps)
        out["SSIM"] = sum(ss) / len(ss)
    return out
