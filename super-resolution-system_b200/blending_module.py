"""GPU mirror of the reference's blending module -- the Laplacian-pyramid tile fusion and its pyramid helpers.

Mirrors (names, arguments, error behaviour) /root/reference/blending_module.py:
  FusionMethod 36-40, WeightType 52-56, TileInfo 96-112, BlendingModule.__init__ 164-190,
  build_gaussian_pyramid 217-269, build_laplacian_pyramid 271-318, collapse_laplacian_pyramid 320-363,
  laplacian_fusion 369-506, _create_distance_weight_map 508-561.
Every pixel is computed by libsrfuse.so (csrc/srf_fuse.cu) through the C ABI; nothing here falls back to the CPU.
`weighted_average_fusion` (row N3, "next") maps onto the same gather kernel with one pyramid level.
Out of scope: Poisson / feather / gradient-domain fusion, seam detection and repair, colour
correction -- those entry points raise NotImplementedError.
"""
from __future__ import annotations

import ctypes as C
import logging
from dataclasses import dataclass
from enum import Enum
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np

from . import _native
from ._arrays import ImageView, current_stream_ptr, is_cuda_tensor, torch

logger = logging.getLogger(__name__)


class FusionMethod(Enum):
    LAPLACIAN = "laplacian"
    POISSON = "poisson"
    WEIGHTED_AVERAGE = "weighted"


class WeightType(Enum):
    LINEAR = "linear"
    COSINE = "cosine"
    SIGMOID = "sigmoid"


@dataclass
class TileInfo:
    """an upscaled tile and where its top-left pixel sits on the canvas"""
    image: np.ndarray
    x: int
    y: int
    row: int
    col: int


@dataclass
class Seam:
    """a window (or a merged group of windows) where the fused result departs from a tile (blending_module.py:59-92)"""
    x: int
    y: int
    width: int
    height: int
    ssim_score: float
    severity: str = "low"
    suggested_fix: str = ""

    def __post_init__(self):
        if self.ssim_score < 0.85:
            self.severity, self.suggested_fix = "high", "poisson_refinement"
        elif self.ssim_score < 0.92:
            self.severity, self.suggested_fix = "medium", "increase_blend_width"
        else:
            self.severity, self.suggested_fix = "low", "none"


class FusionPlan:
    """a marshalled tile list (see BlendingModule.plan_fusion)"""

    def __init__(self, *parts):
        self.parts = parts


def weight_profile(feather_width: int, weight_type: WeightType) -> np.ndarray:
    """The distance-weight map only depends on d = min distance to the tile border, and only through
    min(d, feather_width): this is that 1-D profile, evaluated with the reference's float64 NumPy expressions
    (blending_module.py:543-561) so the float32 values are the reference's bit for bit."""
    d = np.arange(feather_width + 1)
    nd = np.clip(d / feather_width, 0, 1)
    if weight_type == WeightType.LINEAR:
        w = nd
    elif weight_type == WeightType.COSINE:
        w = 0.5 * (1 - np.cos(np.pi * nd))
    elif weight_type == WeightType.SIGMOID:
        w = 1 / (1 + np.exp(-10 * (nd - 0.5)))
    else:
        w = nd
    return np.ascontiguousarray(w.astype(np.float32))


class BlendingModule:
    def __init__(self, method: str = "laplacian", num_levels: int = 6, ssim_threshold: float = 0.95,
                 use_cuda: bool = False, device: int = 0, context_slot: int = 0):
        self.method = FusionMethod(method)  # ValueError on an unknown method, like the reference
        self.num_levels = num_levels
        self.ssim_threshold = ssim_threshold
        self.use_cuda = True  # the only path there is; the reference's flag selected cv2.cuda and is accepted for compatibility
        self.device = int(device)
        self.context_slot = int(context_slot)  # extension: an independent library context on the same device (see _native.context)
        self._profiles = {}
        logger.info("BlendingModule (libsrfuse, sm_100a): method=%s levels=%d device=%d", method, num_levels, self.device)

    # ------------------------------------------------------------------ pyramids (API-compat helpers)
    def _ctx(self):
        return _native.context(self.device, self.context_slot)

    def _pyr_down(self, img: np.ndarray) -> np.ndarray:
        lib = _native.load()
        h, w = img.shape[:2]
        cn = 1 if img.ndim == 2 else img.shape[2]
        out = np.empty(((h + 1) // 2, (w + 1) // 2) + img.shape[2:], np.float32)
        _native.check(lib.srf_pyr_down_f32_host(self._ctx(), img.ctypes.data, h, w, cn, out.ctypes.data), self._ctx())
        return out

    def _pyr_up(self, img: np.ndarray, dst_hw: Tuple[int, int]) -> np.ndarray:
        lib = _native.load()
        h, w = img.shape[:2]
        cn = 1 if img.ndim == 2 else img.shape[2]
        out = np.empty((int(dst_hw[0]), int(dst_hw[1])) + img.shape[2:], np.float32)
        _native.check(lib.srf_pyr_up_f32_host(self._ctx(), img.ctypes.data, h, w, cn, out.ctypes.data, out.shape[0],
                                              out.shape[1]), self._ctx())
        return out

    def build_gaussian_pyramid(self, image: np.ndarray, levels: Optional[int] = None) -> List[np.ndarray]:
        if levels is None:
            levels = self.num_levels
        cur = np.ascontiguousarray(image, dtype=np.float32)
        pyramid = [cur.copy()]
        for i in range(levels - 1):
            if cur.shape[0] < 2 or cur.shape[1] < 2:
                logger.warning("Stopping pyramid at level %d: image too small", i + 1)
                break
            cur = self._pyr_down(cur)
            pyramid.append(cur)
        return pyramid

    def build_laplacian_pyramid(self, gaussian_pyramid: Sequence[np.ndarray]) -> List[np.ndarray]:
        out = []
        for cur, nxt in zip(gaussian_pyramid[:-1], gaussian_pyramid[1:]):
            up = self._pyr_up(np.ascontiguousarray(nxt, dtype=np.float32), cur.shape[:2])
            out.append(cur - up)
        out.append(gaussian_pyramid[-1])
        return out

    def collapse_laplacian_pyramid(self, laplacian_pyramid: Sequence[np.ndarray]) -> np.ndarray:
        cur = np.array(laplacian_pyramid[-1], dtype=np.float32, copy=True)
        for layer in reversed(laplacian_pyramid[:-1]):
            cur = self._pyr_up(np.ascontiguousarray(cur), layer.shape[:2]) + layer
        return cur

    # ------------------------------------------------------------------ weights
    def _profile(self, h: int, w: int, weight_type: WeightType, feather_width: Optional[int] = None) -> np.ndarray:
        if feather_width is None:
            feather_width = min(h, w) // 8
        if feather_width < 1:
            raise ValueError(f"tile {h}x{w} is too small for a distance weight map (feather width {feather_width})")
        key = (feather_width, weight_type)
        if key not in self._profiles:
            self._profiles[key] = weight_profile(feather_width, weight_type)
        return self._profiles[key]

    def _create_distance_weight_map(self, height: int, width: int, weight_type: WeightType,
                                    feather_width: Optional[int] = None) -> np.ndarray:
        """(H, W) float32 map in [0, 1]: the profile indexed by border distance (host-side table lookup; the fusion
        kernels index the same profile on the fly and never materialise this map)."""
        prof = self._profile(height, width, weight_type, feather_width)
        y = np.arange(height).reshape(-1, 1)
        x = np.arange(width).reshape(1, -1)
        d = np.minimum(np.minimum(y, height - 1 - y), np.minimum(x, width - 1 - x))
        return prof[np.minimum(d, len(prof) - 1)]

    # ------------------------------------------------------------------ fusion
    def _marshal(self, tiles, output_shape, weight_type):
        """TileInfo / array list -> (views, srf_tile[], srf_weight_lut[], keepalive, out_h, out_w, channels)"""
        images, positions = [], []
        for i, t in enumerate(tiles):
            if isinstance(t, TileInfo):
                images.append(t.image)
                positions.append((int(t.y), int(t.x)))
            else:
                images.append(t)
                if output_shape:
                    grid = int(np.ceil(np.sqrt(len(tiles))))
                    r, c = i // grid, i % grid
                    positions.append((r * t.shape[0], c * t.shape[1]))
        if not images:
            raise ValueError("laplacian_fusion: no tiles")
        if len(positions) != len(images):
            # the reference dies in max() over an empty sequence here (ValueError); same exception type
            raise ValueError("laplacian_fusion: bare-array tiles need output_shape to derive a grid layout")
        views = [ImageView(im, f"tile {i}") for i, im in enumerate(images)]
        on_dev = views[0].on_device
        if any(v.on_device != on_dev for v in views):
            raise ValueError("laplacian_fusion: mix of host and device tiles")
        chans = views[0].c
        if any(v.c != chans for v in views):
            raise ValueError("laplacian_fusion: tiles disagree on the channel count")
        if output_shape is None:
            out_h = max(y + v.h for v, (y, _) in zip(views, positions))
            out_w = max(x + v.w for v, (_, x) in zip(views, positions))
        else:
            out_h, out_w = int(output_shape[0]), int(output_shape[1])
        lut_index, luts = {}, []
        c_tiles = (_native.srf_tile * len(views))()
        for i, (v, (y, x)) in enumerate(zip(views, positions)):
            prof = self._profile(v.h, v.w, weight_type)
            if id(prof) not in lut_index:
                lut_index[id(prof)] = len(luts)
                luts.append(prof)
            c_tiles[i] = _native.srf_tile(v.ptr, v.pitch, v.h, v.w, y, x, lut_index[id(prof)])
        c_luts = (_native.srf_weight_lut * len(luts))(*[_native.srf_weight_lut(p.ctypes.data, len(p)) for p in luts])
        keep = list(luts)
        if on_dev:  # the kernels want 16-byte aligned rows: re-stage a misaligned tensor (the copy lives as long as `keep`)
            for i, v in enumerate(views):
                if v.ptr % 16 or (v.pitch % 16 and v.h > 1):
                    pitch = (v.w * chans + 15) // 16 * 16
                    buf = torch.empty((v.h, pitch), dtype=torch.uint8, device=v.keep.device)
                    buf[:, : v.w * chans] = v.keep.reshape(v.h, v.w * chans)
                    keep.append(buf)
                    c_tiles[i].data, c_tiles[i].pitch = buf.data_ptr(), pitch
        return views, c_tiles, c_luts, keep, out_h, out_w, chans

    def plan_fusion(self, tiles, output_shape=None, weight_type: WeightType = WeightType.COSINE):
        """Marshal a tile list once (extension): the returned plan can be passed to laplacian_fusion in place of the
        tile list any number of times -- the per-call Python work drops to one ctypes call.  The plan keeps the tile
        arrays alive; their contents may change between calls, their addresses may not (a CUDA tile that had to be
        re-staged for alignment is snapshotted in the plan)."""
        return FusionPlan(*self._marshal(tiles, output_shape, weight_type))

    def laplacian_fusion(self, tiles: List[Union[np.ndarray, TileInfo]], overlap_map=None,
                         output_shape: Optional[Tuple[int, int]] = None,
                         weight_type: WeightType = WeightType.COSINE, row_range: Optional[Tuple[int, int]] = None,
                         out=None, _levels: Optional[int] = None):
        """Fuse upscaled tiles into one uint8 canvas (see include/srfuse.h: srf_laplacian_fuse).

        tiles: TileInfo objects (image + x, y) or bare arrays (grid layout, needs output_shape).  Images are uint8
        NumPy arrays (host path: H2D, kernels, D2H inside the call) or uint8 torch CUDA tensors (device path: the
        result is a CUDA tensor and stays on the device).  `overlap_map` is accepted and ignored, as in the
        reference (blending_module.py:369-506 never reads it).  Extensions: `row_range` restricts the produced
        canvas rows (band sharding across GPUs; other rows are zero / left as they are in `out`), `out` reuses a
        caller-owned canvas.
        """
        if isinstance(tiles, FusionPlan):
            views, c_tiles, c_luts, keep, out_h, out_w, chans = tiles.parts
        else:
            views, c_tiles, c_luts, keep, out_h, out_w, chans = self._marshal(tiles, output_shape, weight_type)
        on_dev = views[0].on_device
        levels = self.num_levels if _levels is None else int(_levels)
        r0, r1 = (0, out_h) if row_range is None else (int(row_range[0]), int(row_range[1]))
        lib, ctx = _native.load(), self._ctx()
        shape = (out_h, out_w) if views[0].keep.ndim == 2 else (out_h, out_w, chans)
        if on_dev:
            dev = views[0].device
            if dev != self.device:
                raise ValueError(f"tiles live on cuda:{dev} but this BlendingModule is bound to cuda:{self.device}")
            if out is not None:
                if not is_cuda_tensor(out) or tuple(out.shape) != shape or out.dtype != torch.uint8 or not out.is_contiguous():
                    raise ValueError("out must be a contiguous uint8 CUDA tensor of the canvas shape")
                canvas = out
            elif (r0, r1) != (0, out_h):
                canvas = torch.zeros(shape, dtype=torch.uint8, device=views[0].keep.device)
            else:
                canvas = torch.empty(shape, dtype=torch.uint8, device=views[0].keep.device)
            rc = lib.srf_laplacian_fuse(ctx, c_tiles, len(views), chans, levels, c_luts, len(c_luts), out_h,
                                        out_w, canvas.data_ptr(), out_w * chans, r0, r1, current_stream_ptr(dev))
            _native.check(rc, ctx)
            return canvas
        if out is not None:
            if not isinstance(out, np.ndarray) or out.shape != shape or out.dtype != np.uint8 or not out.flags.c_contiguous:
                raise ValueError("out must be a C-contiguous uint8 array of the canvas shape")
            canvas = out
        else:
            canvas = np.zeros(shape, np.uint8) if (r0, r1) != (0, out_h) else np.empty(shape, np.uint8)
        rc = lib.srf_laplacian_fuse_host(ctx, c_tiles, len(views), chans, levels, c_luts, len(c_luts), out_h,
                                         out_w, canvas.ctypes.data, out_w * chans, r0, r1)
        _native.check(rc, ctx)
        return canvas

    def fuse_and_assess(self, tiles, reference: np.ndarray, output_shape=None,
                        weight_type: WeightType = WeightType.COSINE, data_range: float = 255.0, out=None, row_range=None,
                        _levels: Optional[int] = None):
        """Host tiles -> (fused canvas, srf_iqa_sums of canvas vs `reference`) in one library call: stage 3 + stage 4
        of the reference pipeline with the canvas kept on the device in between (include/srfuse.h:
        srf_fuse_assess_host)."""
        if isinstance(tiles, FusionPlan):
            views, c_tiles, c_luts, keep, out_h, out_w, chans = tiles.parts
        else:
            views, c_tiles, c_luts, keep, out_h, out_w, chans = self._marshal(tiles, output_shape, weight_type)
        r0, r1 = (0, out_h) if row_range is None else (int(row_range[0]), int(row_range[1]))
        if views[0].on_device:
            raise ValueError("fuse_and_assess takes host arrays; use laplacian_fusion + psnr_ssim_sums for CUDA tensors")
        shape = (out_h, out_w) if views[0].keep.ndim == 2 else (out_h, out_w, chans)
        ref = ImageView(reference, "reference")
        if (ref.h, ref.w, ref.c) != (out_h, out_w, chans) or ref.on_device:
            raise ValueError(f"reference must be a host image of the canvas shape {shape}")
        if out is not None:
            if not isinstance(out, np.ndarray) or out.shape != shape or out.dtype != np.uint8 or not out.flags.c_contiguous:
                raise ValueError("out must be a C-contiguous uint8 array of the canvas shape")
            canvas = out
        else:
            canvas = np.empty(shape, np.uint8) if (r0, r1) == (0, out_h) else np.zeros(shape, np.uint8)
        sums = _native.srf_iqa_sums()
        lib, ctx = _native.load(), self._ctx()
        rc = lib.srf_fuse_assess_host(ctx, c_tiles, len(views), chans, self.num_levels if _levels is None else int(_levels),
                                      c_luts, len(c_luts), out_h, out_w,
                                      canvas.ctypes.data, out_w * chans, ref.ptr, ref.pitch, float(data_range), r0, r1, C.byref(sums))
        _native.check(rc, ctx)
        return canvas, sums

    def gather_plan(self, sizes_hw, positions_yx, output_shape, weight_type: WeightType = WeightType.COSINE, row_range=None,
                    channels: int = 3, levels: Optional[int] = None):
        """Diagnostic (host arithmetic only, no GPU): the cells the canvas gather of laplacian_fusion would use for tiles of
        these sizes at these positions -- a list of dicts {cls, bx0, by0, nbx, nby, tiles, first} in 128 x 16-pixel block
        units (rows counted from the band's first row), or None when the call would take the general kernel
        (include/srfuse.h: srf_gather_plan)."""
        n = len(sizes_hw)
        lut_index, luts = {}, []
        c_tiles = (_native.srf_tile * n)()
        for i, ((h, w), (y, x)) in enumerate(zip(sizes_hw, positions_yx)):
            prof = self._profile(int(h), int(w), weight_type)
            if id(prof) not in lut_index:
                lut_index[id(prof)] = len(luts)
                luts.append(prof)
            c_tiles[i] = _native.srf_tile(None, int(w) * channels, int(h), int(w), int(y), int(x), lut_index[id(prof)])
        c_luts = (_native.srf_weight_lut * len(luts))(*[_native.srf_weight_lut(p.ctypes.data, len(p)) for p in luts])
        out_h, out_w = int(output_shape[0]), int(output_shape[1])
        r0, r1 = (0, out_h) if row_range is None else (int(row_range[0]), int(row_range[1]))
        lib = _native.load()
        count = C.c_int(0)
        lv = self.num_levels if levels is None else int(levels)
        rc = lib.srf_gather_plan(c_tiles, n, channels, lv, c_luts, len(luts), out_h, out_w, r0, r1, None, 0, C.byref(count))
        if rc == -5:  # SRF_ERR_UNSUPPORTED: the general kernel takes this geometry
            return None
        _native.check(rc, None)
        buf = (C.c_int * (12 * max(count.value, 1)))()
        _native.check(lib.srf_gather_plan(c_tiles, n, channels, lv, c_luts, len(luts), out_h, out_w, r0, r1, buf, count.value, C.byref(count)), None)
        rec = np.frombuffer(buf, dtype=np.int32).reshape(-1, 12)[: count.value]
        return [{"cls": int(r[0]), "bx0": int(r[1]), "by0": int(r[2]), "nbx": int(r[3]), "nby": int(r[4]),
                 "tiles": [int(t) for t in r[6:6 + min(int(r[5]), 4)]], "first": int(r[10])} for r in rec]

    # ------------------------------------------------------------------ explicitly out of scope
    def _out_of_scope(self, name):
        raise NotImplementedError(f"BlendingModule.{name} is outside the B200 hot path (SURVEY.md section 8); "
                                  "only laplacian_fusion and the pyramid helpers are provided")

    def poisson_fusion(self, *a, **k):
        self._out_of_scope("poisson_fusion")

    def weighted_average_fusion(self, tiles, weights=None, weight_type: WeightType = WeightType.COSINE,
                                output_shape: Optional[Tuple[int, int]] = None, row_range=None, out=None):
        """Mirror of BlendingModule.weighted_average_fusion (blending_module.py:661-760; row N3 of SURVEY.md 8f):
        sum_t tile_t * W_0^t / max(sum_t W_0^t, 1e-6), clip, truncate -- which is exactly the Laplacian fusion with a
        single pyramid level (no pyramid at all), so it runs the same gather kernel.  Bare-array tiles always get the
        sqrt-grid layout here (the reference does not require output_shape for them).  Custom per-tile weight maps
        (`weights=`, float32 or float64 planes at least as large as the tile's part of the canvas; tiles past the end of
        the list get the generated map) run `srf_weighted_average`, which reproduces NumPy's mixed-precision sums."""
        if weights is not None:
            if row_range is not None or out is not None:
                raise ValueError("weighted_average_fusion(weights=...) produces the whole canvas: row_range / out are not supported")
            return self._weighted_average_custom(list(tiles), list(weights), weight_type, output_shape)
        tiles = list(tiles)
        if tiles and not any(isinstance(t, TileInfo) for t in tiles):
            grid = int(np.ceil(np.sqrt(len(tiles))))
            tiles = [TileInfo(t, (i % grid) * t.shape[1], (i // grid) * t.shape[0], i // grid, i % grid) for i, t in enumerate(tiles)]
        return self.laplacian_fusion(tiles, output_shape=output_shape, weight_type=weight_type, row_range=row_range, out=out,
                                     _levels=1)

    def _weighted_average_custom(self, tiles, weights, weight_type, output_shape):
        """blending_module.py:688-757 with `weights` given: positions as in the reference (TileInfo or sqrt-grid), the
        canvas the maximum extent unless output_shape says otherwise, every tile cropped to it"""
        if not tiles:
            raise ValueError("no tiles")
        images, pos = [], []
        grid = int(np.ceil(np.sqrt(len(tiles))))
        for i, t in enumerate(tiles):
            if isinstance(t, TileInfo):
                images.append(t.image)
                pos.append((int(t.y), int(t.x)))
            else:
                images.append(t)
                pos.append(((i // grid) * int(t.shape[0]), (i % grid) * int(t.shape[1])))
        views = [ImageView(im, f"tile {i}") for i, im in enumerate(images)]
        on_dev, chans = views[0].on_device, views[0].c
        if any(v.on_device != on_dev or v.c != chans for v in views):
            raise ValueError("tiles must be all host or all device images with the same channel count")
        if output_shape is None:
            out_h = max(y + v.h for v, (y, x) in zip(views, pos))
            out_w = max(x + v.w for v, (y, x) in zip(views, pos))
        else:
            out_h, out_w = int(output_shape[0]), int(output_shape[1])
        arr = (_native.srf_wavg_tile * len(views))()
        keep = []
        for i, (v, (y, x)) in enumerate(zip(views, pos)):
            wm = weights[i] if i < len(weights) else self._create_distance_weight_map(v.h, v.w, weight_type)
            if is_cuda_tensor(wm):
                if not on_dev:
                    wm = wm.cpu().numpy()
            if not is_cuda_tensor(wm):
                wm = np.asarray(wm)
            if wm.ndim == 3 and wm.shape[2] == 1 and chans > 1:
                wm = wm[:, :, 0]  # the reference squeezes such a map for its weight accumulator
            if wm.ndim != 2:
                raise ValueError(f"weight map {i}: expected an (H, W) plane, got {tuple(wm.shape)}")
            ch, cw = max(min(y + v.h, out_h) - y, 0), max(min(x + v.w, out_w) - x, 0)
            if wm.shape[0] < ch or wm.shape[1] < cw:  # NumPy would refuse to broadcast tile[:ch, :cw] * weight[:ch, :cw]
                raise ValueError(f"weight map {i} is {tuple(wm.shape)}, smaller than the tile's part of the canvas ({ch}, {cw})")
            if is_cuda_tensor(wm):
                f64 = wm.dtype == torch.float64
                if not f64 and wm.dtype != torch.float32:  # NumPy's promotion against float32, not torch's: 32 / 64-bit integers widen
                    f64 = wm.dtype in (torch.int32, torch.int64)
                    wm = wm.to(torch.float64 if f64 else torch.float32)
                if wm.stride(1) != 1:
                    wm = wm.contiguous()
                wptr, wpitch = wm.data_ptr(), wm.stride(0) * wm.element_size()
            else:
                f64 = np.result_type(np.float32, wm.dtype) == np.float64  # what `float32 tile * map` computes in
                wm = wm.astype(np.float64 if f64 else np.float32, copy=False)
                if wm.strides[1] != wm.itemsize or wm.strides[0] < wm.shape[1] * wm.itemsize:
                    wm = np.ascontiguousarray(wm)
                if on_dev:
                    wm = torch.from_numpy(np.ascontiguousarray(wm)).to(views[0].keep.device)
                    wptr, wpitch = wm.data_ptr(), wm.stride(0) * wm.element_size()
                else:
                    wptr, wpitch = wm.ctypes.data, wm.strides[0] if wm.shape[0] > 1 else wm.shape[1] * wm.itemsize
            keep.append(wm)
            a = arr[i]
            a.data, a.pitch, a.h, a.w, a.y, a.x = v.ptr, v.pitch, v.h, v.w, y, x
            a.weight, a.weight_pitch, a.weight_h, a.weight_w, a.weight_f64 = wptr, wpitch, int(wm.shape[0]), int(wm.shape[1]), int(f64)
        lib = _native.load()
        shape = (out_h, out_w) if views[0].keep.ndim == 2 else (out_h, out_w, chans)
        if on_dev:
            dev = views[0].device
            ctx = _native.context(dev)
            canvas = torch.empty(shape, dtype=torch.uint8, device=views[0].keep.device)
            rc = lib.srf_weighted_average(ctx, arr, len(views), chans, out_h, out_w, canvas.data_ptr(), out_w * chans, current_stream_ptr(dev))
            _native.check(rc, ctx)
            torch.cuda.current_stream(dev).synchronize()  # the uploaded maps in `keep` must outlive the kernel
            return canvas
        ctx = self._ctx()
        canvas = np.empty(shape, np.uint8)
        rc = lib.srf_weighted_average_host(ctx, arr, len(views), chans, out_h, out_w, canvas.ctypes.data, out_w * chans)
        _native.check(rc, ctx)
        return canvas

    # ------------------------------------------------------------------ seam detection (blending_module.py:765-968)
    def window_ssim_map(self, result, tile, x: int, y: int, window_size: int = 16, stride: int = 8,
                        k1: float = 0.01, k2: float = 0.03, L: float = 255.0) -> np.ndarray:
        """(ny, nx) float64 scores of every window of the region `tile` (placed at x, y) shares with `result`:
        _compute_ssim(tile_window, result_window) of the reference for all windows at once (csrc/srf_seam.cu)"""
        vr, vt = ImageView(result, "result"), ImageView(tile, "tile")
        if vr.on_device != vt.on_device or vr.c != vt.c:
            raise ValueError("result and tile must be both host or both device images with the same channel count")
        lib = _native.load()
        ny, nx = C.c_int(), C.c_int()
        _native.check(lib.srf_window_ssim_shape(vr.h, vr.w, vt.h, vt.w, int(y), int(x), int(window_size), int(stride),
                                                C.byref(ny), C.byref(nx)), None)
        c1, c2 = (k1 * L) ** 2, (k2 * L) ** 2
        if ny.value == 0 or nx.value == 0:
            return np.zeros((ny.value, nx.value), np.float64)
        if vr.on_device:
            ctx = _native.context(vr.device)
            out = torch.empty((ny.value, nx.value), dtype=torch.float64, device=vr.keep.device)
            rc = lib.srf_window_ssim(ctx, vr.ptr, vr.pitch, vr.h, vr.w, vt.ptr, vt.pitch, vt.h, vt.w, vr.c, int(y), int(x),
                                     int(window_size), int(stride), c1, c2, out.data_ptr(), current_stream_ptr(vr.device))
            _native.check(rc, ctx)
            return out.cpu().numpy()
        ctx = self._ctx()
        out = np.empty((ny.value, nx.value), np.float64)
        rc = lib.srf_window_ssim_host(ctx, vr.ptr, vr.pitch, vr.h, vr.w, vt.ptr, vt.pitch, vt.h, vt.w, vr.c, int(y), int(x),
                                      int(window_size), int(stride), c1, c2, out.ctypes.data)
        _native.check(rc, ctx)
        return out

    def _compute_ssim(self, img1, img2, k1: float = 0.01, k2: float = 0.03, L: float = 255.0) -> float:
        """whole-patch SSIM of two equally sized images (blending_module.py:853-903), gray by the reference's
        COLOR_BGR2GRAY on the stored channel order"""
        a, b = np.asarray(img1), np.asarray(img2)
        if a.shape != b.shape or a.shape[0] != a.shape[1]:
            raise ValueError("_compute_ssim on the GPU takes two equally sized square patches")
        return float(self.window_ssim_map(b, a, 0, 0, window_size=a.shape[0], stride=a.shape[0], k1=k1, k2=k2, L=L)[0, 0])

    def detect_seams(self, result, tiles, window_size: int = 16, stride: int = 8) -> List[Seam]:
        """Windows of every tile whose SSIM against the fused result is below ssim_threshold, neighbours merged
        (blending_module.py:765-851).  The result is uploaded once; every tile is one kernel launch."""
        infos = [t if isinstance(t, TileInfo) else TileInfo(t, 0, 0, 0, 0) for t in tiles]
        if not is_cuda_tensor(result) and torch is not None and torch.cuda.is_available():
            result = torch.from_numpy(np.ascontiguousarray(result)).to(torch.device("cuda", self.device))
        seams: List[Seam] = []
        for info in infos:
            img = info.image
            if is_cuda_tensor(result) and not is_cuda_tensor(img):
                img = torch.from_numpy(np.ascontiguousarray(img)).to(result.device)
            if info.x >= result.shape[1] or info.y >= result.shape[0]:
                continue
            scores = self.window_ssim_map(result, img, info.x, info.y, window_size, stride)
            ys, xs = np.nonzero(scores < self.ssim_threshold)  # row-major: y outer, x inner, as the reference's loops
            for wy, wx in zip(ys.tolist(), xs.tolist()):
                seams.append(Seam(x=info.x + wx * stride, y=info.y + wy * stride, width=window_size, height=window_size,
                                  ssim_score=float(scores[wy, wx])))
        seams = self._merge_adjacent_seams(seams, distance_threshold=window_size)
        logger.info("Detected %d seams", len(seams))
        return seams

    def _merge_adjacent_seams(self, seams: List[Seam], distance_threshold: int = 16) -> List[Seam]:
        """sort by (y, x), chain a seam to the current group while it lies closer than the threshold to the group's LAST
        member, replace every group by its bounding box and mean score (blending_module.py:905-968)"""
        if not seams:
            return []
        ordered = sorted(seams, key=lambda s: (s.y, s.x))
        merged, group = [], [ordered[0]]
        for s in ordered[1:]:
            last = group[-1]
            if np.sqrt((s.x - last.x) ** 2 + (s.y - last.y) ** 2) < distance_threshold:
                group.append(s)
            else:
                merged.append(self._merge_seam_group(group))
                group = [s]
        merged.append(self._merge_seam_group(group))
        return merged

    def _merge_seam_group(self, group: List[Seam]) -> Seam:
        if len(group) == 1:
            return group[0]
        x0, y0 = min(s.x for s in group), min(s.y for s in group)
        x1, y1 = max(s.x + s.width for s in group), max(s.y + s.height for s in group)
        return Seam(x=x0, y=y0, width=x1 - x0, height=y1 - y0, ssim_score=np.mean([s.ssim_score for s in group]))
