"""oracle/tiling_oracle.py -- TEST INFRASTRUCTURE.

NumPy restatement of the reference tiler's geometry and padding:
  tile grid        /root/reference/tiling_module.py:572-608  (_calculate_tile_positions)
  overlap metadata /root/reference/tiling_module.py:610-646  (_calculate_overlap_for_tile)
  bottom/right pad /root/reference/tiling_module.py:522-570, 718-724 (cv2.copyMakeBorder with
                   BORDER_REFLECT_101 / REPLICATE / REFLECT / CONSTANT(0); multi-bounce when pad > size-1)
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
