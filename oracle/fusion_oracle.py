"""oracle/fusion_oracle.py -- TEST INFRASTRUCTURE.

NumPy restatement of the reference's Laplacian-pyramid tile fusion, in the reference's float32 op order:
  Gaussian pyramid     /root/reference/blending_module.py:217-269
  Laplacian pyramid    /root/reference/blending_module.py:271-318
  collapse             /root/reference/blending_module.py:320-363
  fusion               /root/reference/blending_module.py:369-506
  distance weight map  /root/reference/blending_module.py:508-561

Per tile t (independent pyramids, tile-local borders):
    G_0 = f32(tile);  G_{l+1} = pyrDown(G_l)            (stops early if a side < 2)
    W_0 = weight map; W_{l+1} = pyrDown(W_l)
    C_top = G_top * W_top;   C_l = pyrUp(C_{l+1}) + (G_l - pyrUp(G_{l+1})) * W_l
    acc[y:y+h, x:x+w] += C_0 (cropped to the canvas);  wacc += W_0
out = uint8(trunc(clip(acc / max(wacc, 1e-6), 0, 255)))
"""
from __future__ import annotations

import numpy as np

from . import pyr

WEIGHT_KINDS = ("linear", "cosine", "sigmoid")


def weight_profile(feather: int, kind: str) -> np.ndarray:
    """W as a function of d = min distance to the tile border, d = 0..feather (float64 math, then f32),
    i.e. blending_module.py:543-561 evaluated on the only values it can take."""
    d = np.arange(feather + 1)
    nd = np.clip(d / feather, 0, 1)
    if kind == "cosine":
        w = 0.5 * (1 - np.cos(np.pi * nd))
    elif kind == "sigmoid":
        w = 1 / (1 + np.exp(-10 * (nd - 0.5)))
    else:
        w = nd
    return w.astype(np.float32)


def weight_map(h: int, w: int, kind: str = "cosine", feather=None) -> np.ndarray:
    if feather is None:
        feather = min(h, w) // 8
    y = np.arange(h).reshape(-1, 1)
    x = np.arange(w).reshape(1, -1)
    d = np.minimum(np.minimum(y, h - 1 - y), np.minimum(x, w - 1 - x))
    return weight_profile(feather, kind)[np.minimum(d, feather)]


def gaussian_pyramid(img: np.ndarray, levels: int, backend="c"):
    cur = img.astype(np.float32)
    out = [cur]
    for _ in range(levels - 1):
        if cur.shape[0] < 2 or cur.shape[1] < 2:
            break
        cur = pyr.pyr_down(cur, backend)
        out.append(cur)
    return out


def weighted_collapse(g, wp, backend="c"):
    """C_0 of one tile from its Gaussian pyramid g and weight pyramid wp."""
    def ex(a):
        return a[..., None] if g[0].ndim == 3 else a
    cur = g[-1] * ex(wp[-1])
    for l in range(len(g) - 2, -1, -1):
        hw = g[l].shape[:2]
        lap = g[l] - pyr.pyr_up(g[l + 1], hw, backend)
        cur = pyr.pyr_up(cur, hw, backend) + lap * ex(wp[l])
    return cur


def laplacian_fuse(tiles, positions_yx, out_hw, levels=6, kind="cosine", backend="c"):
    """tiles: list of uint8 (h,w,C) or (h,w); positions_yx: list of (y, x); returns uint8 canvas."""
    H, W = out_hw
    ch = tiles[0].shape[2:]
    acc = np.zeros((H, W) + ch, np.float32)
    wacc = np.zeros((H, W), np.float32)
    cache = {}
    for img, (y, x) in zip(tiles, positions_yx):
        h, w = img.shape[:2]
        key = (h, w)
        if key not in cache:  # identical for same-shape tiles; the reference recomputes it per tile
            w0 = weight_map(h, w, kind)
            cache[key] = gaussian_pyramid(w0, levels, backend)
        wp = cache[key]
        g = gaussian_pyramid(img, levels, backend)
        c0 = weighted_collapse(g, wp, backend)
        ye, xe = min(y + h, H), min(x + w, W)
        if ye <= y or xe <= x:
            continue
        acc[y:ye, x:xe] += c0[:ye - y, :xe - x]
        wacc[y:ye, x:xe] += wp[0][:ye - y, :xe - x]
    wacc = np.maximum(wacc, np.float32(1e-6))
    res = acc / (wacc[..., None] if ch else wacc)
    return np.clip(res, 0, 255).astype(np.uint8)


def weighted_average_custom(tiles, positions_yx, weights, out_hw=None, kind="cosine"):
    """weighted_average_fusion WITH the `weights=` argument (/root/reference/blending_module.py:688-757), every NumPy
    promotion written out: a float32 map multiplies and accumulates in float32; a float64 map makes the product float64
    and the in-place `+=` on the float32 accumulators run in float64 before the result is stored back as float32.
    Tiles past the end of `weights` get the generated distance map (float32).  (h, w, 1) maps count as planes."""
    if out_hw is None:
        out_hw = (max(y + t.shape[0] for t, (y, x) in zip(tiles, positions_yx)), max(x + t.shape[1] for t, (y, x) in zip(tiles, positions_yx)))
    H, W = out_hw
    ch = tiles[0].shape[2:]
    acc = np.zeros((H, W) + ch, np.float32)
    wacc = np.zeros((H, W), np.float32)
    for i, (img, (y, x)) in enumerate(zip(tiles, positions_yx)):
        h, w = img.shape[:2]
        wm = np.asarray(weights[i]) if i < len(weights) else weight_map(h, w, kind)
        ye, xe = min(y + h, H), min(x + w, W)
        if ye <= y or xe <= x:
            continue
        t = img[:ye - y, :xe - x].astype(np.float32)
        m = wm[:ye - y, :xe - x]
        m = m.reshape(m.shape[:2])
        wide = np.result_type(np.float32, m.dtype) == np.float64
        m = m.astype(np.float64 if wide else np.float32)
        mc = m[..., None] if ch else m
        if wide:
            acc[y:ye, x:xe] = (acc[y:ye, x:xe].astype(np.float64) + t.astype(np.float64) * mc).astype(np.float32)
            wacc[y:ye, x:xe] = (wacc[y:ye, x:xe].astype(np.float64) + m).astype(np.float32)
        else:
            acc[y:ye, x:xe] = acc[y:ye, x:xe] + t * mc
            wacc[y:ye, x:xe] = wacc[y:ye, x:xe] + m
    wacc = np.maximum(wacc, np.float32(1e-6))
    res = acc / (wacc[..., None] if ch else wacc)
    return np.clip(res, 0, 255).astype(np.uint8)


# ---- seam detection (blending_module.py:765-968): per-window whole-patch SSIM -------------------------------------------
def gray_bgr_order(img: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(img, cv2.COLOR_BGR2GRAY) for uint8 -- what the reference applies to its RGB arrays (channel 0 gets the
    blue weight): (c0*3735 + c1*19235 + c2*9798 + 16384) >> 15"""
    if img.ndim == 2:
        return img
    i = img.astype(np.int64)
    return ((i[..., 0] * 3735 + i[..., 1] * 19235 + i[..., 2] * 9798 + 16384) >> 15).astype(np.uint8)


def patch_ssim(img1: np.ndarray, img2: np.ndarray, k1=0.01, k2=0.03, L=255.0) -> float:
    """_compute_ssim (blending_module.py:853-903), literally: float64 np.mean / np.var / covariance of the gray patches"""
    a, b = gray_bgr_order(img1).astype(np.float64), gray_bgr_order(img2).astype(np.float64)
    mu1, mu2 = np.mean(a), np.mean(b)
    s1, s2 = np.var(a), np.var(b)
    s12 = np.mean((a - mu1) * (b - mu2))
    c1, c2 = (k1 * L) ** 2, (k2 * L) ** 2
    return float(((2 * mu1 * mu2 + c1) * (2 * s12 + c2)) / ((mu1 ** 2 + mu2 ** 2 + c1) * (s1 + s2 + c2)))


def window_ssim_map(result: np.ndarray, tile: np.ndarray, x: int, y: int, window=16, stride=8) -> np.ndarray:
    """scores of detect_seams' sliding windows for one tile (blending_module.py:806-835), tile first"""
    y2, x2 = min(y + tile.shape[0], result.shape[0]), min(x + tile.shape[1], result.shape[1])
    if x2 <= x or y2 <= y:
        return np.zeros((0, 0))
    roi_r, roi_t = result[y:y2, x:x2], tile[:y2 - y, :x2 - x]
    ys = range(0, roi_r.shape[0] - window + 1, stride)
    xs = range(0, roi_r.shape[1] - window + 1, stride)
    out = np.zeros((len(ys), len(xs)))
    for j, wy in enumerate(ys):
        for i, wx in enumerate(xs):
            out[j, i] = patch_ssim(roi_t[wy:wy + window, wx:wx + window], roi_r[wy:wy + window, wx:wx + window])
    return out
