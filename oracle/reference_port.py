"""oracle/reference_port.py -- TEST / BASELINE INFRASTRUCTURE.

The reference's CPU path with the reference's own COST structure, for bench.py's `cpu_baseline` leg and
`--impl reference` arm (the reference is Python and cannot travel to the GPU box; this port can).  Same
arithmetic as oracle/fusion_oracle.py and oracle/iqa_oracle.py -- results are asserted equal in
tests/test_oracle.py -- but nothing is cached or shared that the reference recomputes:
  * per tile: astype(float32), Gaussian pyramid (with the two full-size copies of blending_module.py:245-246),
    Laplacian pyramid, the float64 trigonometric weight map and ITS pyramid (blending_module.py:448-453 rebuilds
    both for every tile), per-level multiply, collapse, accumulate (436-492);
  * then maximum / divide / clip / cast over the canvas (496-506);
  * PSNR and SSIM as two separate passes (quality_assessment_module.py:499, 504), SSIM in float64 through
    scipy.ndimage.gaussian_filter when SciPy is importable (that is what scikit-image calls), else the NumPy taps.
`threads` > 1 runs tiles concurrently (NumPy and OpenCV release the GIL) and accumulates in list order, which is
more parallelism than the reference itself uses: the baseline gets every host core it can use.
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import iqa_oracle, pyr


def _weight_map_trig(h, w, kind):
    fw = min(h, w) // 8
    y = np.arange(h).reshape(-1, 1)
    x = np.arange(w).reshape(1, -1)
    d = np.minimum(np.minimum(y, h - 1 - y), np.minimum(x, w - 1 - x))
    nd = np.clip(d / fw, 0, 1)
    if kind == "cosine":
        wgt = 0.5 * (1 - np.cos(np.pi * nd))
    elif kind == "sigmoid":
        wgt = 1 / (1 + np.exp(-10 * (nd - 0.5)))
    else:
        wgt = nd
    return wgt.astype(np.float32)


def _gauss(img, levels, be):
    if img.dtype != np.float32:
        img = img.astype(np.float32)
    out = [img.copy()]
    cur = img.copy()
    for _ in range(levels - 1):
        if cur.shape[0] < 2 or cur.shape[1] < 2:
            break
        cur = pyr.pyr_down(cur, be)
        out.append(cur)
    return out


def _tile_contribution(img, levels, kind, be):
    f = img.astype(np.float32) if img.dtype != np.float32 else img
    g = _gauss(f, levels, be)
    lap = [g[i] - pyr.pyr_up(g[i + 1], g[i].shape[:2], be) for i in range(len(g) - 1)] + [g[-1]]
    w0 = _weight_map_trig(img.shape[0], img.shape[1], kind)
    wp = _gauss(w0, levels, be)
    wl = [l * (w[..., None] if l.ndim == 3 else w) for l, w in zip(lap, wp)]
    cur = wl[-1].copy()
    for i in range(len(wl) - 2, -1, -1):
        cur = pyr.pyr_up(cur, wl[i].shape[:2], be) + wl[i]
    return cur, w0


def laplacian_fuse(tiles, positions_yx, out_hw, levels=6, kind="cosine", threads=1):
    be = "cv2" if pyr.have_cv2() else "c"
    H, W = out_hw
    ch = tiles[0].shape[2:]
    acc = np.zeros((H, W) + ch, np.float32)
    wacc = np.zeros((H, W), np.float32)

    def work(i):
        return _tile_contribution(tiles[i], levels, kind, be)

    if threads > 1:
        with ThreadPoolExecutor(threads) as ex:
            results = ex.map(work, range(len(tiles)))
            results = list(results)
    else:
        results = (work(i) for i in range(len(tiles)))
    for (c0, w0), img, (y, x) in zip(results, tiles, positions_yx):
        h, w = img.shape[:2]
        ye, xe = min(y + h, H), min(x + w, W)
        if ye <= y or xe <= x:
            continue
        acc[y:ye, x:xe] += c0[:ye - y, :xe - x]
        wacc[y:ye, x:xe] += w0[:ye - y, :xe - x]
    wacc = np.maximum(wacc, 1e-6)
    res = acc / (wacc[..., None] if ch else wacc)
    return np.clip(res, 0, 255).astype(np.uint8)


def psnr(a, b):
    return iqa_oracle.psnr(a, b)


def ssim(a, b):
    try:
        from scipy import ndimage
    except Exception:
        return iqa_oracle.ssim(a, b)
    a, b = iqa_oracle.common_crop(iqa_oracle.preprocess(a), iqa_oracle.preprocess(b))
    if a.ndim == 3:
        try:
            import cv2
            ga, gb = cv2.cvtColor(a, cv2.COLOR_RGB2GRAY), cv2.cvtColor(b, cv2.COLOR_RGB2GRAY)
        except Exception:
            ga, gb = iqa_oracle.rgb_to_gray(a), iqa_oracle.rgb_to_gray(b)
    else:
        ga, gb = a, b
    x, y = ga.astype(np.float64), gb.astype(np.float64)

    def f(z):
        return ndimage.gaussian_filter(z, iqa_oracle.SIGMA, truncate=3.5, mode="reflect")

    ux, uy, uxx, uyy, uxy = f(x), f(y), f(x * x), f(y * y), f(x * y)
    vx, vy, vxy = uxx - ux * ux, uyy - uy * uy, uxy - ux * uy
    c1, c2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux ** 2 + uy ** 2 + c1) * (vx + vy + c2))
    r = iqa_oracle.RADIUS
    return float(s[r:-r, r:-r].mean(dtype=np.float64))
