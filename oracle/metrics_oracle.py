"""CPU oracle (TEST INFRASTRUCTURE, never on the product path) for the reconstruction metrics of the stride / pool /
bucket experiments: numpy restatement of

* `save_videos_grid`'s frame quantisation, /root/reference/hyvideo/utils/file_utils.py:58-66;
* `compute_psnr` and `compute_ssim`, /root/reference/evaluation/compute_metrics.py:31-41, and the averaging over all
  frames of all videos of its main loop (:129-152).

`compute_ssim` calls scikit-image's `structural_similarity`, a third-party dependency that is neither vendored in
/root/reference nor pinned in its requirements.txt, and is not installed here: its published algorithm (Wang et al.
2004 as implemented by skimage.metrics._structural_similarity: uniform 7x7 window through scipy.ndimage.uniform_filter,
sample covariance, K1 = 0.01, K2 = 0.03, crop of (win_size - 1) // 2 border pixels, float64 for uint8 input, per-channel
mean then mean over channels) is restated below on top of the same scipy filter.  PARITY UNPINNED for that function:
the reference holds no test or golden value for it; `ssim_frame_bruteforce` (explicit window loops) cross-checks the
restatement.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.ndimage import uniform_filter


def video_to_frames_u8(video: np.ndarray, rescale: bool = True) -> np.ndarray:
    """file_utils.py:58-66 for a batch of one: (C, T, H, W) float32 -> [T][H][W][C] uint8."""
    x = np.asarray(video, dtype=np.float32).transpose(1, 2, 3, 0)  # 'c t h w -> t h w c' (make_grid of one image is the image)
    if rescale:
        x = (x + np.float32(1.0)) / np.float32(2.0)
    x = np.clip(x, 0, 1)
    return (x * np.float32(255)).astype(np.uint8)


def psnr_frame(img1: np.ndarray, img2: np.ndarray) -> float:
    """compute_metrics.py:31-36."""
    mse = np.mean((img1 / 255.0 - img2 / 255.0) ** 2)
    if mse < 1.0e-10:
        return 100
    return 20 * math.log10(1 / math.sqrt(mse))


def _ssim_channel(im1: np.ndarray, im2: np.ndarray, data_range: float, win_size: int = 7) -> float:
    K1, K2 = 0.01, 0.03
    im1, im2 = im1.astype(np.float64), im2.astype(np.float64)
    NP = win_size ** im1.ndim
    cov_norm = NP / (NP - 1)  # use_sample_covariance=True
    ux, uy = uniform_filter(im1, size=win_size), uniform_filter(im2, size=win_size)
    uxx, uyy, uxy = uniform_filter(im1 * im1, size=win_size), uniform_filter(im2 * im2, size=win_size), uniform_filter(im1 * im2, size=win_size)
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    R = data_range
    C1, C2 = (K1 * R) ** 2, (K2 * R) ** 2
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
    pad = (win_size - 1) // 2
    return float(S[pad:S.shape[0] - pad, pad:S.shape[1] - pad].mean(dtype=np.float64))


def ssim_frame(img1: np.ndarray, img2: np.ndarray) -> float:
    """compute_metrics.py:38-41: HWC uint8 frames; constant frames score 1; data_range from the FIRST image."""
    if np.all(img1 == img1[0, 0, 0]) or np.all(img2 == img2[0, 0, 0]):
        return 1.0
    data_range = float(img1.max() - img1.min())
    return float(np.mean([_ssim_channel(img1[..., c], img2[..., c], data_range) for c in range(img1.shape[-1])]))


def ssim_frame_bruteforce(img1: np.ndarray, img2: np.ndarray) -> float:
    """Same quantity with explicit 7x7 window sums over the window centres that survive the crop (small frames only)."""
    if np.all(img1 == img1[0, 0, 0]) or np.all(img2 == img2[0, 0, 0]):
        return 1.0
    R = float(img1.max() - img1.min())
    C1, C2, NP = (0.01 * R) ** 2, (0.03 * R) ** 2, 49.0
    H, W, C = img1.shape
    a, b = img1.astype(np.float64), img2.astype(np.float64)
    per_ch = []
    for c in range(C):
        tot = 0.0
        for y in range(3, H - 3):
            for x in range(3, W - 3):
                p, q = a[y - 3:y + 4, x - 3:x + 4, c], b[y - 3:y + 4, x - 3:x + 4, c]
                ux, uy = p.sum() / NP, q.sum() / NP
                vx = NP / (NP - 1) * ((p * p).sum() / NP - ux * ux)
                vy = NP / (NP - 1) * ((q * q).sum() / NP - uy * uy)
                vxy = NP / (NP - 1) * ((p * q).sum() / NP - ux * uy)
                tot += ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2))
        per_ch.append(tot / ((H - 6) * (W - 6)))
    return float(np.mean(per_ch))


def compare_videos(pairs) -> dict:
    """Main loop of compute_metrics.py:129-152: `pairs` = iterable of (frames1, frames2) uint8 [T][H][W][C]; frames are
    zipped (the shorter video decides), every frame contributes one PSNR and one SSIM, the result is the plain mean."""
    ps, ss = [], []
    for f1, f2 in pairs:
        for a, b in zip(f1, f2):
            ps.append(psnr_frame(a, b))
            ss.append(ssim_frame(a, b))
    out = {}
    if ps:
        out["PSNR"] = sum(ps) / len(ps)
    if ss:
        out["SSIM"] = sum(ss) / len(ss)
    return out
