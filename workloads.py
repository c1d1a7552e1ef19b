"""Synthetic workloads of BASELINE.json's configs (harness code shared by bench.py and the tests; not product code).

A workload is what the reference pipeline would hand to its fusion + assessment stages:
    source (1280x720 textured RGB, seeded) --bicubic stub--> half-resolution image --tiler--> 2048^2 tiles
    --bicubic stub SR x2--> 4096^2 tiles at canvas positions (2*gx, 2*gy); reference canvas = bicubic stub of the
    source at canvas size (what PSNR / SSIM compare against).
The tile grid is the one the reference's tiler really produces: 12 / 20 / 24 tiles for the
100 / 150 / 200 MP configs -- BASELINE.json's "~16 / ~25 / ~36" are the README's nominal figures.
The geometry helpers below restate the tiler's grid arithmetic so the harness does not depend on the product or on
the oracle; the pixel work uses cv2 (bicubic stub + mirror padding), exactly what the reference side would call.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Tuple

import numpy as np

CONFIGS = {
    # name: (target WxH, overlap ratio, pyramid levels)
    "cfg1_100mp_720p": ((12245, 6887), 0.2, 5),   # configs[0]: 1280x720 source keeps 16:9 -> 4x2 = 8 tiles
    "cfg2_100mp": ((12244, 8162), 0.2, 5),        # configs[1]: 4x3 = 12 tiles
    "cfg3_150mp": ((15000, 10000), 0.3, 6),       # configs[2]: 5x4 = 20 tiles
    "cfg4_200mp": ((17320, 11546), 0.2, 7),       # configs[3]: 6x4 = 24 tiles  <- the metric's config
}
BLOCK = 2048


@dataclass
class Workload:
    name: str
    tiles: List[np.ndarray]            # uint8 (4096, 4096, 3) each
    positions_yx: List[Tuple[int, int]]
    canvas_hw: Tuple[int, int]
    levels: int
    reference: np.ndarray              # uint8 canvas-sized comparison image
    grid: Tuple[int, int]

    @property
    def megapixels(self) -> float:
        return self.canvas_hw[0] * self.canvas_hw[1] / 1e6

    @property
    def tile_bytes(self) -> int:
        return sum(t.nbytes for t in self.tiles)


def synthetic_source(h: int = 720, w: int = 1280, seed: int = 0) -> np.ndarray:
    import cv2
    rng = np.random.default_rng(seed)
    f = cv2.GaussianBlur(rng.random((h, w, 3), dtype=np.float32) * 255, (0, 0), 3)
    f = (f - f.mean()) * 4 + 128 + rng.normal(0, 4, f.shape)
    return np.clip(f, 0, 255).astype(np.uint8)


def grid_positions(width: int, height: int, block: int, overlap: int):
    step = block - overlap
    nx = max(1, math.ceil((width - overlap) / step))
    ny = max(1, math.ceil((height - overlap) / step))
    pos = [(i * step, j * step, min(block, width - i * step), min(block, height - j * step)) for j in range(ny) for i in range(nx)]
    return pos, (nx, ny)


def stub_sr(img: np.ndarray, size_hw) -> np.ndarray:
    import cv2
    return cv2.resize(img, (int(size_hw[1]), int(size_hw[0])), interpolation=cv2.INTER_CUBIC)


@dataclass
class IqaWorkload:
    """BASELINE.json configs[4]: full-reference IQA sweep"""
    name: str
    big_a: np.ndarray                  # uint8 (H, W, 3): the 200 MP comparison canvas of cfg4
    big_b: np.ndarray                  # the same + N(0, 5), re-clipped (seed 1)
    pairs: List[Tuple[np.ndarray, np.ndarray]]  # 16 x (2048, 2048, 3) pairs, seeds 2 .. 17

    @property
    def megapixels(self) -> float:
        return (self.big_a.shape[0] * self.big_a.shape[1] + sum(a.shape[0] * a.shape[1] for a, _ in self.pairs)) / 1e6


def perturb(img: np.ndarray, seed: int, sigma: float = 5.0) -> np.ndarray:
    """img + N(0, sigma) re-clipped to uint8, in row strips (a 200 MP image would need 5 GB of float64 at once)"""
    rng = np.random.default_rng(seed)
    out = np.empty_like(img)
    for r in range(0, img.shape[0], 512):
        blk = img[r:r + 512].astype(np.float32)
        blk += rng.normal(0, sigma, blk.shape).astype(np.float32)
        out[r:r + 512] = np.clip(blk, 0, 255).astype(np.uint8)
    return out


def build_iqa(name: str = "cfg5_iqa", big_rows: int = 0, n_pairs: int = 16, pair_size: int = 2048) -> IqaWorkload:
    """big_rows > 0 keeps only the first rows of the 200 MP pair (bounded CPU samples / small tests)"""
    (tw, th), _, _ = CONFIGS["cfg4_200mp"]
    src = synthetic_source(seed=0)
    half_h, half_w = th // 2, tw // 2
    big_a = np.ascontiguousarray(stub_sr(src, (2 * half_h, 2 * half_w)))
    if big_rows:
        big_a = np.ascontiguousarray(big_a[:big_rows])
    big_b = perturb(big_a, seed=1)
    pairs = []
    for k in range(n_pairs):
        a = synthetic_source(pair_size, pair_size, seed=2 + k)
        pairs.append((a, perturb(a, seed=100 + k)))
    return IqaWorkload(name, big_a, big_b, pairs)


def build(name: str = "cfg4_200mp", seed: int = 0, max_tiles: int = 0) -> Workload:
    import cv2
    (tw, th), ratio, levels = CONFIGS[name]
    src = synthetic_source(seed=seed)
    half = stub_sr(src, (th // 2, tw // 2))
    overlap = int(BLOCK * ratio)
    pos, grid = grid_positions(half.shape[1], half.shape[0], BLOCK, overlap)
    if max_tiles:
        pos = pos[:max_tiles]
    tiles, positions = [], []
    for (x, y, w, h) in pos:
        crop = half[y:y + h, x:x + w]
        if h < BLOCK or w < BLOCK:
            crop = cv2.copyMakeBorder(crop, 0, BLOCK - h, 0, BLOCK - w, cv2.BORDER_REFLECT_101)
        tiles.append(np.ascontiguousarray(stub_sr(crop, (2 * BLOCK, 2 * BLOCK))))
        positions.append((2 * y, 2 * x))
    canvas_hw = (2 * half.shape[0], 2 * half.shape[1])
    reference = np.ascontiguousarray(stub_sr(src, canvas_hw))
    return Workload(name, tiles, positions, canvas_hw, levels, reference, grid)
