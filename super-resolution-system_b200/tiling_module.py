"""GPU mirror of the reference's tiler: overlap grid, crop, bottom/right padding to block_size.

Mirrors /root/reference/tiling_module.py: PaddingMode 40-45, TileStatus 48-54, TileMetadata 64-125, Tile 128-171,
TilingModule.__init__ 446-504, _apply_padding 522-570, _calculate_tile_positions 572-608,
_calculate_overlap_for_tile 610-646, split_image 671-784, neighbour bookkeeping 786-850, merge_tiles 1074-1135,
_create_blend_weight 1137-1175.
Grid and overlap arithmetic run in libsrfuse's exact host functions; the crop + pad gather runs on the GPU
(csrc/srf_api.cu: k_tile_extract_pad), and so does the feathered merge (csrc/srf_merge.cu).  Out of scope: content analysis (Haar / MSER /
saliency), the L1/L2/L3 caches, JSON checkpoints and streaming -- metadata decoration and host bookkeeping.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import logging
import threading
import time
import uuid
from dataclasses import asdict, dataclass, field
from enum import Enum, auto
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _native
from ._arrays import ImageView, current_stream_ptr, torch

logger = logging.getLogger(__name__)


class PaddingMode(Enum):
    MIRROR = "mirror"        # cv2.BORDER_REFLECT_101
    REPLICATE = "replicate"  # cv2.BORDER_REPLICATE
    REFLECT = "reflect"      # cv2.BORDER_REFLECT
    CONSTANT = "constant"    # zeros


_BORDER_CODE = {PaddingMode.MIRROR: 0, PaddingMode.REPLICATE: 1, PaddingMode.REFLECT: 2, PaddingMode.CONSTANT: 3}


class TileStatus(Enum):
    PENDING = auto()
    PROCESSING = auto()
    COMPLETED = auto()
    FAILED = auto()
    CACHED = auto()


class CacheLevel(Enum):
    L1_MEMORY = "L1"
    L2_DISK = "L2"
    L3_CLOUD = "L3"


@dataclass
class TileMetadata:
    block_id: str = field(default_factory=lambda: str(uuid.uuid4()))
    global_x: int = 0
    global_y: int = 0
    input_w: int = 2048
    input_h: int = 2048
    output_w: int = 4096
    output_h: int = 4096
    overlap_top: int = 0
    overlap_bottom: int = 0
    overlap_left: int = 0
    overlap_right: int = 0
    roi_flags: Dict[str, bool] = field(default_factory=dict)
    status: TileStatus = TileStatus.PENDING
    neighbor_ids: Dict[str, Optional[str]] = field(
        default_factory=lambda: {"top": None, "bottom": None, "left": None, "right": None})
    image_hash: str = ""
    complexity_score: float = 0.0
    priority: int = 0
    created_at: float = field(default_factory=time.time)
    updated_at: float = field(default_factory=time.time)

    def to_dict(self) -> Dict:
        d = asdict(self)
        d["status"] = self.status.name
        return d

    @classmethod
    def from_dict(cls, data: Dict) -> "TileMetadata":
        d = dict(data)
        d["status"] = TileStatus[d["status"]]
        return cls(**d)


@dataclass
class Tile:
    metadata: TileMetadata
    data: Optional[np.ndarray] = None
    mask: Optional[np.ndarray] = None
    cache_path: Optional[str] = None

    def get_overlap_region(self) -> Tuple[int, int, int, int]:
        m = self.metadata
        return (m.overlap_top, m.overlap_bottom, m.overlap_left, m.overlap_right)

    def get_effective_region(self) -> Tuple[int, int, int, int]:
        m = self.metadata
        x1, y1 = m.global_x + m.overlap_left, m.global_y + m.overlap_top
        return (x1, y1, x1 + m.input_w - m.overlap_left - m.overlap_right,
                y1 + m.input_h - m.overlap_top - m.overlap_bottom)


def _read_rgb(path: str) -> np.ndarray:
    """decode to HWC uint8 RGB (I/O, not part of the hot path); unreadable -> ValueError like the reference"""
    img = None
    try:
        import cv2
        bgr = cv2.imread(path)
        if bgr is not None:
            img = np.ascontiguousarray(bgr[:, :, ::-1])
    except ImportError:
        from PIL import Image
        try:
            img = np.asarray(Image.open(path).convert("RGB"))
        except Exception:
            img = None
    if img is None:
        raise ValueError(f"cannot load image: {path}")
    return img


class TilingModule:
    def __init__(self, block_size: int = 2048, overlap_ratio: float = 0.2, padding_mode: str = "mirror",
                 output_scale: float = 2.0, l1_cache_size: int = 50, l2_cache_dir: Optional[str] = None,
                 enable_content_aware: bool = True, device: int = 0):
        if not (0.1 <= overlap_ratio <= 0.3):
            raise ValueError(f"overlap_ratio must be within 0.1-0.3, got {overlap_ratio}")
        self.block_size = block_size
        self.overlap_ratio = overlap_ratio
        self.padding_mode = PaddingMode(padding_mode)  # ValueError on an unknown mode
        self.output_scale = output_scale
        self.output_size = int(block_size * output_scale)
        self.overlap_pixels = int(block_size * overlap_ratio)
        # content analysis only decorates metadata (roi_flags, complexity_score); it is outside the GPU path
        self.enable_content_aware = False
        self.content_analyzer = None
        if enable_content_aware:
            logger.debug("content-aware analysis is out of scope for the B200 path; tiles carry no roi_flags")
        self.l1_cache_size = l1_cache_size
        self.l2_cache_dir = l2_cache_dir
        self.device = int(device)
        self.tile_registry: Dict[str, Tile] = {}
        self.registry_lock = threading.Lock()
        self.processing_state: Dict[str, dict] = {}

    # ------------------------------------------------------------------ exact host geometry
    def _calculate_tile_positions(self, image_width: int, image_height: int) -> List[Tuple[int, int, int, int]]:
        lib = _native.load()
        n = _native.check(lib.srf_tile_positions(image_width, image_height, self.block_size, self.overlap_pixels, None, 0),
                          exc=ValueError)
        buf = (C.c_int * (4 * n))()
        _native.check(lib.srf_tile_positions(image_width, image_height, self.block_size, self.overlap_pixels, buf, n))
        return [tuple(buf[4 * i: 4 * i + 4]) for i in range(n)]

    def _calculate_overlap_for_tile(self, x, y, w, h, image_width, image_height) -> Tuple[int, int, int, int]:
        out = (C.c_int * 4)()
        _native.check(_native.load().srf_tile_overlaps(x, y, w, h, image_width, image_height, self.block_size,
                                                      self.overlap_pixels, out))
        return tuple(out)

    def _compute_image_hash(self, image_path: str) -> str:
        md5 = hashlib.md5()
        with open(image_path, "rb") as f:
            for chunk in iter(lambda: f.read(1 << 16), b""):
                md5.update(chunk)
        return md5.hexdigest()

    # ------------------------------------------------------------------ GPU crop + pad
    def _extract(self, image, positions) -> "np.ndarray":
        """(n, block, block, C) uint8 stack of padded tiles; NumPy in -> NumPy out, CUDA tensor in -> CUDA tensor out"""
        v = ImageView(image, "image")
        n, b = len(positions), self.block_size
        xywh = (C.c_int * (4 * n))(*[q for p in positions for q in p])
        lib, ctx = _native.load(), _native.context(self.device)
        mode = _BORDER_CODE[self.padding_mode]
        shape = (n, b, b) if v.keep.ndim == 2 else (n, b, b, v.c)
        if v.on_device:
            out = torch.empty(shape, dtype=torch.uint8, device=v.keep.device)
            rc = lib.srf_tile_extract_pad(ctx, v.ptr, v.h, v.w, v.c, v.pitch, xywh, n, b, mode, out.data_ptr(),
                                          current_stream_ptr(v.device))
        else:
            out = np.empty(shape, np.uint8)
            rc = lib.srf_tile_extract_pad_host(ctx, v.ptr, v.h, v.w, v.c, v.pitch, xywh, n, b, mode, out.ctypes.data)
        _native.check(rc, ctx)
        return out

    def _apply_padding(self, image, pad_top: int, pad_bottom: int, pad_left: int, pad_right: int):
        """cv2.copyMakeBorder(image, top, bottom, left, right, <padding_mode>) on the GPU (tiling_module.py:522-570; any
        combination of pads, padding larger than the image included).  NumPy in -> NumPy out, CUDA tensor in -> CUDA tensor out."""
        pads = [int(pad_top), int(pad_bottom), int(pad_left), int(pad_right)]
        if min(pads) < 0:
            raise ValueError(f"negative padding {pads}")  # cv2.copyMakeBorder asserts on it
        v = ImageView(image, "image")
        oh, ow = v.h + pads[0] + pads[1], v.w + pads[2] + pads[3]
        shape = (oh, ow) if v.keep.ndim == 2 else (oh, ow, v.c)
        lib, ctx = _native.load(), _native.context(v.device if v.on_device else self.device)
        mode = _BORDER_CODE[self.padding_mode]
        if v.on_device:
            out = torch.empty(shape, dtype=torch.uint8, device=v.keep.device)
            rc = lib.srf_copy_make_border(ctx, v.ptr, v.h, v.w, v.c, v.pitch, *pads, mode, out.data_ptr(), current_stream_ptr(v.device))
        else:
            out = np.empty(shape, np.uint8)
            rc = lib.srf_copy_make_border_host(ctx, v.ptr, v.h, v.w, v.c, v.pitch, *pads, mode, out.ctypes.data)
        _native.check(rc, ctx)
        return out

    # ------------------------------------------------------------------ public API
    def split_array(self, image, image_hash: str = "", save_metadata: bool = True) -> List[Tile]:
        """split an in-memory HWC uint8 RGB image (NumPy or CUDA tensor); split_image = decode + this"""
        height, width = int(image.shape[0]), int(image.shape[1])
        positions = self._calculate_tile_positions(width, height)
        stack = self._extract(image, positions)
        tiles = []
        for i, (x, y, w, h) in enumerate(positions):
            top, bottom, left, right = self._calculate_overlap_for_tile(x, y, w, h, width, height)
            meta = TileMetadata(global_x=x, global_y=y, input_w=w, input_h=h, output_w=int(w * self.output_scale),
                                output_h=int(h * self.output_scale), overlap_top=top, overlap_bottom=bottom,
                                overlap_left=left, overlap_right=right, image_hash=image_hash, status=TileStatus.PENDING)
            tile = Tile(metadata=meta, data=stack[i])
            tiles.append(tile)
            if save_metadata:
                with self.registry_lock:
                    self.tile_registry[meta.block_id] = tile
        self._build_neighbor_relationships(tiles)
        self.processing_state[image_hash] = {"image_width": width, "image_height": height, "num_tiles": len(tiles),
                                             "tile_ids": [t.metadata.block_id for t in tiles], "timestamp": time.time()}
        return tiles

    def split_image(self, image_path: str, save_metadata: bool = True) -> List[Tile]:
        image = _read_rgb(image_path)
        image_hash = self._compute_image_hash(image_path)
        tiles = self.split_array(image, image_hash, save_metadata)
        self.processing_state[image_hash]["image_path"] = image_path
        return tiles

    def create_tile_metadata(self, tile: Tile, global_x: int, global_y: int) -> TileMetadata:
        """re-anchor a tile's metadata (tiling_module.py:648-669)"""
        m = tile.metadata
        m.global_x, m.global_y, m.updated_at = global_x, global_y, time.time()
        return m

    def _build_neighbor_relationships(self, tiles: List[Tile]) -> None:
        at = {(t.metadata.global_x, t.metadata.global_y): t.metadata.block_id for t in tiles}
        step = self.block_size - self.overlap_pixels
        for t in tiles:
            x, y = t.metadata.global_x, t.metadata.global_y
            for name, key in (("top", (x, y - step)), ("bottom", (x, y + step)), ("left", (x - step, y)),
                              ("right", (x + step, y))):
                if key in at:
                    t.metadata.neighbor_ids[name] = at[key]

    def get_neighbor_tiles(self, tile_id: str) -> List[Tile]:
        with self.registry_lock:
            tile = self.tile_registry.get(tile_id)
            if tile is None:
                return []
            return [self.tile_registry[n] for n in tile.metadata.neighbor_ids.values() if n and n in self.tile_registry]

    # ------------------------------------------------------------------ merge (feathered weighted average)
    def _blend_tables(self, tile: Tile):
        """the three 1-D factor tables of _create_blend_weight (tiling_module.py:1137-1175).  The reference multiplies a
        float32 map of ones IN PLACE by float64 linspace ramps over the top rows, bottom rows, left columns and right
        columns (scaled overlap metadata): the first two only depend on the row, the last two on the column, so
        map[y, x] = f32(f32(wy[y] * wl[x]) * wr[x]) with float64 products -- wy float32, wl / wr float64 (1 outside)."""
        m = tile.metadata
        h, w = int(m.output_h), int(m.output_w)
        top, bottom = int(m.overlap_top * self.output_scale), int(m.overlap_bottom * self.output_scale)
        left, right = int(m.overlap_left * self.output_scale), int(m.overlap_right * self.output_scale)
        if max(top, bottom) > h or max(left, right) > w:
            raise ValueError("scaled overlap exceeds the tile (the reference's ramp would not broadcast)")
        wy = np.ones(h, np.float32)
        if top > 0:
            wy[:top] *= np.linspace(0, 1, top)
        if bottom > 0:
            wy[h - bottom:] *= np.linspace(1, 0, bottom)
        wl, wr = np.ones(w, np.float64), np.ones(w, np.float64)
        if left > 0:
            wl[:left] = np.linspace(0, 1, left)
        if right > 0:
            wr[w - right:] = np.linspace(1, 0, right)
        return wy, wl, wr

    def _create_blend_weight(self, tile: Tile) -> np.ndarray:
        """the reference's (output_h, output_w) float32 feather map, built from the same tables the GPU merge uses"""
        wy, wl, wr = self._blend_tables(tile)
        m = (wy[:, None].astype(np.float64) * wl[None, :]).astype(np.float32)
        return (m.astype(np.float64) * wr[None, :]).astype(np.float32)

    def merge_tiles(self, tiles: List[Tile], output_width: int, output_height: int, blending: bool = True):
        """Feathered weighted average of the processed tiles (tiling_module.py:1074-1135) on the GPU: tile data at
        int(global_x * scale), int(global_y * scale); data whose size is not (output_h, output_w) is squashed with
        cv2.resize's default INTER_LINEAR first (the reference's treatment of the padded edge tiles); float32
        accumulation in list order, / max(sum of weights, 1e-6), astype(uint8).  NumPy tiles -> NumPy canvas, CUDA
        tensors -> CUDA tensor."""
        live = [t for t in tiles if t.data is not None]
        if not live:
            return np.zeros((output_height, output_width, 3), np.uint8)
        views = [ImageView(t.data, "tile.data") for t in live]
        on_dev = views[0].on_device
        if any(v.on_device != on_dev for v in views) or any(v.c != views[0].c for v in views):
            raise ValueError("tiles must be all host or all device arrays with the same channel count")
        ch = views[0].c
        arr = (_native.srf_merge_tile * len(live))()
        keep = []
        for k, (t, v) in enumerate(zip(live, views)):
            m = t.metadata
            a = arr[k]
            a.data, a.pitch, a.h, a.w = v.ptr, v.pitch, v.h, v.w
            a.out_h, a.out_w = int(m.output_h), int(m.output_w)
            a.y, a.x = int(m.global_y * self.output_scale), int(m.global_x * self.output_scale)
            if blending:
                wy, wl, wr = self._blend_tables(t)
                keep.append((wy, wl, wr))
                a.wy, a.wl, a.wr = wy.ctypes.data, wl.ctypes.data, wr.ctypes.data
        lib = _native.load()
        shape = (output_height, output_width) if (ch == 1 and len(live[0].data.shape) == 2) else (output_height, output_width, ch)
        if on_dev:
            dev = views[0].device
            ctx = _native.context(dev)
            canvas = torch.empty(shape, dtype=torch.uint8, device=views[0].keep.device)
            rc = lib.srf_merge_tiles(ctx, arr, len(live), ch, output_height, output_width, canvas.data_ptr(), output_width * ch,
                                     current_stream_ptr(dev))
        else:
            ctx = _native.context(self.device)
            canvas = np.empty(shape, np.uint8)
            rc = lib.srf_merge_tiles_host(ctx, arr, len(live), ch, output_height, output_width, canvas.ctypes.data, output_width * ch)
        _native.check(rc, ctx)
        return canvas

    def restore_from_cache(self, image_hash: str):
        """the reference's checkpoint restore (tiling_module.py:1040-1072) is host bookkeeping, out of scope"""
        return None
