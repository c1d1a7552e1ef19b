"""oracle/iqa_oracle.py -- TEST INFRASTRUCTURE.

Float64 restatement of the reference's full-reference metrics:
  preprocess  /root/reference/quality_assessment_module.py:169-195  (max()<=1 -> *255 truncating cast)
  PSNR        /root/reference/quality_assessment_module.py:277-320  -> skimage.metrics.peak_signal_noise_ratio
              (= 10*log10(R^2 / mean((a-b)^2)) over every element, float64; manual fallback 313-318 identical)
  SSIM        /root/reference/quality_assessment_module.py:322-389  -> cv2.cvtColor(RGB2GRAY) then
              skimage.metrics.structural_similarity(gaussian_weights=True, sigma=1.5,
              use_sample_covariance=False, data_range=255)
scikit-image (requirements.txt:13, >=0.21.0) is absent from the reference tree and from this image; its published
algorithm is restated here: 11 taps exp(-k^2/(2*1.5^2)) normalised (radius int(3.5*1.5+0.5)=5), filtered
ux,uy,uxx,uyy,uxy, S = ((2uxuy+C1)(2vxy+C2))/((ux^2+uy^2+C1)(vx+vy+C2)), mean over S[5:-5,5:-5] -- so only
'valid' filter outputs are used and the border mode never matters.
"""
from __future__ import annotations

import numpy as np

RADIUS = 5
SIGMA = 1.5


def gaussian_taps() -> np.ndarray:
    x = np.arange(-RADIUS, RADIUS + 1)
    phi = np.exp(-0.5 / (SIGMA * SIGMA) * x ** 2)
    return phi / phi.sum()


def preprocess(img: np.ndarray) -> np.ndarray:
    if img.max() <= 1.0:
        img = (img * 255).astype(np.uint8)
    return img


def common_crop(a, b):
    if a.shape != b.shape:
        h, w = min(a.shape[0], b.shape[0]), min(a.shape[1], b.shape[1])
        a, b = a[:h, :w], b[:h, :w]
    return a, b


def rgb_to_gray(img: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(RGB2GRAY) on uint8: 15-bit fixed point (probed exact against cv2 4.13)."""
    r, g, b = (img[..., i].astype(np.int64) for i in range(3))
    return ((r * 9798 + g * 19235 + b * 3735 + 16384) >> 15).astype(np.uint8)


def sq_err_sum(a, b) -> int:
    a, b = common_crop(preprocess(a), preprocess(b))
    d = a.astype(np.int64) - b.astype(np.int64)
    return int((d * d).sum())


def psnr(a, b, data_range=255.0) -> float:
    a, b = common_crop(preprocess(a), preprocess(b))
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    if mse == 0:
        return float("inf")
    return float(10 * np.log10((data_range ** 2) / mse))


def _valid_filter(x: np.ndarray, taps: np.ndarray) -> np.ndarray:
    """separable 'valid' correlation, axis 0 then axis 1 (scipy.ndimage.gaussian_filter's axis order)."""
    n = len(taps)
    h, w = x.shape
    t = np.zeros((h - n + 1, w))
    for k in range(n):
        t += taps[k] * x[k:k + h - n + 1, :]
    o = np.zeros((h - n + 1, w - n + 1))
    for k in range(n):
        o += taps[k] * t[:, k:k + w - n + 1]
    return o


def ssim_map_interior(ga: np.ndarray, gb: np.ndarray, data_range=255.0) -> np.ndarray:
    x = ga.astype(np.float64)
    y = gb.astype(np.float64)
    t = gaussian_taps()
    ux, uy = _valid_filter(x, t), _valid_filter(y, t)
    uxx, uyy, uxy = _valid_filter(x * x, t), _valid_filter(y * y, t), _valid_filter(x * y, t)
    vx, vy, vxy = uxx - ux * ux, uyy - uy * uy, uxy - ux * uy
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    return ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux ** 2 + uy ** 2 + c1) * (vx + vy + c2))


def ssim(a, b, data_range=255.0) -> float:
    a, b = common_crop(preprocess(a), preprocess(b))
    if a.ndim == 3:
        a, b = rgb_to_gray(a), rgb_to_gray(b)
    if min(a.shape) < 2 * RADIUS + 1:
        raise ValueError("image smaller than the 11x11 window")
    return float(ssim_map_interior(a, b, data_range).mean(dtype=np.float64))


def ssim_uniform(a, b, data_range=255.0) -> float:
    """scikit-image's DEFAULT structural_similarity (what calculate_ssim(multiscale=False) calls,
    quality_assessment_module.py:378-384): 7x7 uniform window (scipy.ndimage.uniform_filter), SAMPLE covariance
    (cov_norm = 49/48), mean over the map cropped by 3.  Parity for this key is pinned to the published algorithm and
    to SciPy only -- the reference has no fallback of its own for it."""
    a, b = common_crop(preprocess(a), preprocess(b))
    if a.ndim == 3:
        a, b = rgb_to_gray(a), rgb_to_gray(b)
    if min(a.shape) < 7:
        raise ValueError("image smaller than the 7x7 window")
    x, y = a.astype(np.float64), b.astype(np.float64)
    t = np.full(7, 1.0 / 7.0)
    ux, uy = _valid_filter(x, t), _valid_filter(y, t)
    uxx, uyy, uxy = _valid_filter(x * x, t), _valid_filter(y * y, t), _valid_filter(x * y, t)
    cn = 49.0 / 48.0
    vx, vy, vxy = cn * (uxx - ux * ux), cn * (uyy - uy * uy), cn * (uxy - ux * uy)
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux ** 2 + uy ** 2 + c1) * (vx + vy + c2))
    return float(s.mean(dtype=np.float64))


def ssim_simple(a, b, data_range=255.0) -> float:
    """_calculate_ssim_simple (/root/reference/quality_assessment_module.py:391-417), the path calculate_ssim takes when
    scikit-image raises or is absent -- i.e. what the reference ACTUALLY returns in this image: cv2.GaussianBlur((11, 11),
    1.5) in float64 with BORDER_REFLECT_101 and the mean over the FULL map.  Filtering with that border is filtering the
    image reflect-padded by the window radius, so the full-map mean is the interior mean of the padded pair: equal to the
    reference's own outputs (tests/golden: iqa_ssim_fallback, iqa_demo_ssim_fallback) to 1e-16."""
    a, b = common_crop(preprocess(np.asarray(a)), preprocess(np.asarray(b)))
    if min(a.shape[:2]) <= RADIUS:
        raise ValueError("image smaller than the window radius")
    pad = ((RADIUS, RADIUS), (RADIUS, RADIUS)) + (((0, 0),) if a.ndim == 3 else ())
    return ssim(np.pad(a, pad, mode="reflect"), np.pad(b, pad, mode="reflect"), data_range)
