"""oracle/resize_oracle.py -- TEST INFRASTRUCTURE.

NumPy restatement of `cv2.resize(img, (dw, dh), interpolation=cv2.INTER_CUBIC)` on uint8 images, the arithmetic of
  downsample_bicubic   /root/reference/quality_assessment_module.py:226-253
  upsample_bicubic     /root/reference/quality_assessment_module.py:255-275   (also the bicubic SR stub of SURVEY 8d)
  _evaluate_downsample_comparison  /root/reference/quality_assessment_module.py:518-555

Third-party arithmetic: opencv-python (requirements.txt, ">=4.8", probed build 4.13.0, "Intel IPP 2022.2 / Custom HAL:
ipp").  For uint8 INTER_CUBIC the wheel does NOT run OpenCV's own 11-bit fixed-point code (that path differs on ~4 %
of the pixels); it calls Intel IPP (closed source), whose result is the real-valued Keys cubic (a = -0.75), evaluated at
float64 sample positions, rounded half-to-even and saturated.  Restated from that published definition:

  scale = 1 / (dst / src)                      (OpenCV's inv_scale -> scale, float64)
  f     = (d + 0.5) * scale - 0.5,  s = floor(f),  t = f - s
  taps  = src[clamp(s - 1 .. s + 2)],  weights = Keys(t) (OpenCV's interpolateCubic formula, in float64 -> float32)
  out   = rint(sum_x wx * (sum_y wy * src))    float32, rows first, products and sums rounded separately

PARITY: pinned to cv2 4.13 itself (tests/golden/make_golden.py stores cv2's outputs): identical on all but <= 2e-5 of
the pixels, never more than 1 LSB apart (IPP's internal op order is unknown; the stragglers are rounding ties).  The
GPU kernel follows THIS restatement bit for bit; PSNR / SSIM of a down-sampled pair move by < 1e-6 relative.
"""
from __future__ import annotations

import numpy as np

A = -0.75


def cubic_weights(t: np.ndarray) -> np.ndarray:
    """OpenCV interpolateCubic (imgproc, a = -0.75) in float64; t in [0, 1) -> (n, 4)"""
    x = np.asarray(t, np.float64)
    c0 = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A
    c1 = ((A + 2) * x - (A + 3)) * x * x + 1
    c2 = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1
    c3 = 1.0 - c0 - c1 - c2
    return np.stack([c0, c1, c2, c3], -1)


def axis_table(src: int, dst: int):
    """per destination index: the 4 clamped source indices and their float32 weights"""
    inv = dst / src
    scale = 1.0 / inv
    d = np.arange(dst, dtype=np.float64)
    f = (d + 0.5) * scale - 0.5
    s = np.floor(f)
    idx = np.clip(s.astype(np.int64)[:, None] + np.arange(-1, 3)[None, :], 0, src - 1)
    return idx, cubic_weights(f - s).astype(np.float32)


def resize_cubic_u8(img: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(img, (dw, dh), interpolation=cv2.INTER_CUBIC) for uint8 (H, W) or (H, W, C)"""
    if img.dtype != np.uint8:
        raise TypeError("uint8 only")
    squeeze = img.ndim == 2
    im = (img[:, :, None] if squeeze else img).astype(np.float32)
    h, w = im.shape[:2]
    ix, cx = axis_table(w, dw)
    iy, cy = axis_table(h, dh)
    t = np.zeros((dh, w, im.shape[2]), np.float32)
    for k in range(4):  # rows first: t += src[iy_k] * wy_k (float32 product, then float32 sum)
        t = t + im[iy[:, k]] * cy[:, k][:, None, None]
    o = np.zeros((dh, dw, im.shape[2]), np.float32)
    for k in range(4):
        o = o + t[:, ix[:, k]] * cx[:, k][None, :, None]
    out = np.clip(np.rint(o), 0, 255).astype(np.uint8)
    return out[:, :, 0] if squeeze else out


def linear_table(src: int, dst: int, vertical: bool = False):
    """OpenCV's fixed-point bilinear taps (imgproc resize, INTER_LINEAR, uint8): source index pair and the two 11-bit
    coefficients per destination index; the sample position is float32.  The two axes treat the ends differently:
    the HORIZONTAL table resets the fraction to 0 where the sample falls before the first or on / after the last source
    column; the VERTICAL pass keeps the fraction and clamps each of the two row indices on its own, so above the first
    (below the last) source row both taps read the same row with coefficients a0 + a1 = 2048 -- and each product is
    truncated separately, which is visible in the last bit.  Only enlarging reaches those rows."""
    inv = dst / src
    scale = 1.0 / inv
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    fr = (f - s.astype(np.float32)).astype(np.float32)
    if not vertical:
        lo, hi = s < 0, s >= src - 1
        fr[lo | hi] = 0
        s[lo] = 0
        s[hi] = src - 1
    a1 = np.rint(fr * np.float32(2048)).astype(np.int32)                    # saturate_cast<short>(cvRound)
    a0 = np.rint((np.float32(1) - fr) * np.float32(2048)).astype(np.int32)
    return np.clip(s, 0, src - 1), np.clip(s + 1, 0, src - 1), a0, a1


def resize_linear_u8(img: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(img, (dw, dh)) (INTER_LINEAR, the default) for uint8, shrinking or enlarging -- what
    TilingModule.merge_tiles applies to a tile whose data is not (output_h, output_w)
    (/root/reference/tiling_module.py:1107-1109).  OpenCV's own integer code (IPP is not used for 8-bit linear unless
    its "not exact" mode is on): horizontal pass in int32 with 11-bit coefficients, vertical pass
    ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2.  Bit-exact against cv2 4.13 on 1 500 random shapes
    in every direction (tests/test_oracle.py)."""
    if img.dtype != np.uint8:
        raise TypeError("uint8 only")
    h, w = img.shape[:2]
    x0, x1, ax0, ax1 = linear_table(w, dw)
    y0, y1, ay0, ay1 = linear_table(h, dh, vertical=True)
    im = (img[:, :, None] if img.ndim == 2 else img).astype(np.int32)
    hz = im[:, x0, :] * ax0[None, :, None] + im[:, x1, :] * ax1[None, :, None]
    s0, s1 = hz[y0], hz[y1]
    out = (((ay0[:, None, None] * (s0 >> 4)) >> 16) + ((ay1[:, None, None] * (s1 >> 4)) >> 16) + 2) >> 2
    out = np.clip(out, 0, 255).astype(np.uint8)
    return out[:, :, 0] if img.ndim == 2 else out


def downsample_bicubic(img: np.ndarray, scale_factor: float) -> np.ndarray:
    """quality_assessment_module.py:226-253"""
    if scale_factor >= 1.0 or scale_factor <= 0:
        raise ValueError(f"scale_factor must lie in (0, 1), got {scale_factor}")
    h, w = img.shape[:2]
    return resize_cubic_u8(img, int(w * scale_factor), int(h * scale_factor))


def downsample_comparison(original: np.ndarray, upscaled: np.ndarray, scale_factors=(0.1, 0.2, 0.4), names=None):
    """quality_assessment_module.py:518-555 with the oracle's PSNR / 7x7-uniform SSIM"""
    from . import iqa_oracle

    names = names or {0.1: "structure_color", 0.2: "mid_frequency", 0.4: "high_frequency"}
    out = {}
    for s in scale_factors:
        name = names.get(s, f"scale_{s}")
        sr, hr = downsample_bicubic(upscaled, s), downsample_bicubic(original, s)
        mh, mw = min(sr.shape[0], hr.shape[0]), min(sr.shape[1], hr.shape[1])
        sr, hr = sr[:mh, :mw], hr[:mh, :mw]
        out[f"psnr_{name}"] = iqa_oracle.psnr(hr, sr)
        out[f"ssim_{name}"] = iqa_oracle.ssim_uniform(hr, sr)
    return out
