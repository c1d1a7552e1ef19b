"""oracle/tiling_oracle.py -- TEST INFRASTRUCTURE.

NumPy restatement of the reference tiler's geometry and padding:
  tile grid        /root/reference/tiling_module.py:572-608  (_calculate_tile_positions)
  overlap metadata /root/reference/tiling_module.py:610-646  (_calculate_overlap_for_tile)
  bottom/right pad /root/reference/tiling_module.py:522-570, 718-724 (cv2.copyMakeBorder with
                   BORDER_REFLECT_101 / REPLICATE / REFLECT / CONSTANT(0); multi-bounce when pad > size-1)
  merge_tiles      /root/reference/tiling_module.py:1074-1135 and _create_blend_weight 1137-1175: feathered weighted
                   average of the processed tiles (linear ramps over the scaled overlap metadata)
"""
from __future__ import annotations

import math

import numpy as np

MODES = ("mirror", "replicate", "reflect", "constant")


def overlap_pixels(block: int, ratio: float) -> int:
    return int(block * ratio)  # tiling_module.py:482


def tile_positions(width: int, height: int, block: int, overlap: int):
    step = block - overlap
    nx = max(1, int(math.ceil((width - overlap) / step)))
    ny = max(1, int(math.ceil((height - overlap) / step)))
    out = []
    for j in range(ny):
        for i in range(nx):
            x, y = i * step, j * step
            out.append((x, y, min(block, width - x), min(block, height - y)))
    return out


def tile_overlaps(x, y, w, h, width, height, block, overlap):
    top = overlap if y > 0 else 0
    left = overlap if x > 0 else 0
    bottom = overlap if y + h < height else 0
    right = overlap if x + w < width else 0
    if y + block >= height:
        bottom = max(0, block - (height - y) - top)
    if x + block >= width:
        right = max(0, block - (width - x) - left)
    return top, bottom, left, right


def border_index(i: int, n: int, mode: str) -> int:
    """source index for out-of-range coordinate i on an axis of length n; -1 = constant zero."""
    if 0 <= i < n:
        return i
    if mode == "constant":
        return -1
    if mode == "replicate":
        return 0 if i < 0 else n - 1
    if n == 1:
        return 0
    delta = 1 if mode == "mirror" else 0  # mirror = BORDER_REFLECT_101, reflect = BORDER_REFLECT
    while i < 0 or i >= n:
        i = (-i - 1 + delta) if i < 0 else (2 * n - 1 - i - delta)
    return i


def extract_tile(image: np.ndarray, x, y, w, h, block, mode):
    """crop + bottom/right pad to (block, block, C) exactly like split_image (tiling_module.py:713-724)."""
    crop = image[y:y + h, x:x + w]
    ys = np.array([border_index(i, h, mode) for i in range(block)])
    xs = np.array([border_index(i, w, mode) for i in range(block)])
    out = crop[np.clip(ys, 0, None)][:, np.clip(xs, 0, None)].copy()
    if mode == "constant":
        out[ys < 0] = 0
        out[:, xs < 0] = 0
    return out


def blend_weight(out_h: int, out_w: int, ov_top: int, ov_bottom: int, ov_left: int, ov_right: int) -> np.ndarray:
    """_create_blend_weight (tiling_module.py:1137-1175) for already scaled overlaps: float32 ones multiplied in place,
    in this order, by float64 linspace ramps over the top rows, bottom rows, left columns, right columns"""
    wgt = np.ones((out_h, out_w), np.float32)
    if ov_top > 0:
        wgt[:ov_top, :] *= np.linspace(0, 1, ov_top).reshape(-1, 1)
    if ov_bottom > 0:
        wgt[-ov_bottom:, :] *= np.linspace(1, 0, ov_bottom).reshape(-1, 1)
    if ov_left > 0:
        wgt[:, :ov_left] *= np.linspace(0, 1, ov_left).reshape(1, -1)
    if ov_right > 0:
        wgt[:, -ov_right:] *= np.linspace(1, 0, ov_right).reshape(1, -1)
    return wgt


def merge_tiles(tiles, output_width: int, output_height: int, scale: float = 2.0, blending: bool = True) -> np.ndarray:
    """merge_tiles (tiling_module.py:1074-1135).  tiles: (data uint8 (h, w, 3), global_x, global_y, output_w, output_h,
    (overlap_top, overlap_bottom, overlap_left, overlap_right)) in list order; data whose size differs from
    (output_h, output_w) is squashed with cv2.resize's default INTER_LINEAR first, as the reference does"""
    from .resize_oracle import resize_linear_u8

    out = np.zeros((output_height, output_width, 3), np.float32)
    wmap = np.zeros((output_height, output_width), np.float32)
    for data, gx, gy, ow, oh, ov in tiles:
        x, y = int(gx * scale), int(gy * scale)
        if data.shape[0] != oh or data.shape[1] != ow:
            data = resize_linear_u8(data, ow, oh)
        td = data.astype(np.float32)
        if blending:
            wgt = blend_weight(oh, ow, *[int(v * scale) for v in ov])
        else:
            wgt = np.ones((oh, ow), np.float32)
        y1, x1 = min(y + oh, output_height), min(x + ow, output_width)
        th, tw = y1 - y, x1 - x
        out[y:y1, x:x1] += td[:th, :tw] * wgt[:th, :tw, None]
        wmap[y:y1, x:x1] += wgt[:th, :tw]
    out = out / np.maximum(wmap, 1e-6)[:, :, None]
    return out.astype(np.uint8)
