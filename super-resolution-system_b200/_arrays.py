"""Small helpers shared by the host-side mirrors: classify an image argument (NumPy host array or torch CUDA
tensor) and hand back pointer / pitch / shape without touching the pixels."""
from __future__ import annotations

import numpy as np

try:  # torch is plumbing only (device memory + streams); the host path works without it
    import torch
except Exception:  # pragma: no cover
    torch = None


def is_cuda_tensor(x) -> bool:
    return torch is not None and isinstance(x, torch.Tensor) and x.is_cuda


def current_stream_ptr(device_index: int) -> int:
    return int(torch.cuda.current_stream(device_index).cuda_stream)


_CAST_HELP = ("the GPU path computes on uint8 pixels (the tiles the pipeline produces); other dtypes are accepted when every "
              "value is an integer in [0, 255] -- the reference's `astype(float32)` of such a tile (blending_module.py:440-441) "
              "is then the same numbers")


def _exact_u8_array(a: np.ndarray, what: str) -> np.ndarray:
    """float32 / float64 / wider-integer pixels that ARE uint8 values -> uint8; anything else is a TypeError"""
    if a.dtype == np.bool_ or not (np.issubdtype(a.dtype, np.floating) or np.issubdtype(a.dtype, np.integer)):
        raise TypeError(f"{what}: expected uint8 pixels, got {a.dtype}; {_CAST_HELP}")
    u = a.astype(np.uint8) if a.size == 0 else np.clip(a, 0, 255).astype(np.uint8)
    if not np.array_equal(u, a):
        raise TypeError(f"{what}: {a.dtype} pixels that are not exact uint8 values; {_CAST_HELP}")
    return u


def _exact_u8_tensor(t, what: str):
    if not (t.dtype.is_floating_point or t.dtype in (torch.int16, torch.int32, torch.int64, torch.int8)):
        raise TypeError(f"{what}: CUDA tensors must hold uint8 pixels (got {t.dtype}); {_CAST_HELP}")
    u = t.clamp(0, 255).to(torch.uint8)
    if not bool((u.to(t.dtype) == t).all()):
        raise TypeError(f"{what}: {t.dtype} pixels that are not exact uint8 values; {_CAST_HELP}")
    return u


class ImageView:
    """uint8 (h, w) or (h, w, C) image with unit element stride inside a row."""

    __slots__ = ("ptr", "pitch", "h", "w", "c", "on_device", "device", "keep")

    def __init__(self, img, what="image"):
        if is_cuda_tensor(img):
            if img.dtype != torch.uint8:
                img = _exact_u8_tensor(img, what)
            if img.dim() not in (2, 3):
                raise ValueError(f"{what}: expected (H, W) or (H, W, C), got {tuple(img.shape)}")
            c = 1 if img.dim() == 2 else img.shape[2]
            ok = (img.dim() == 2 and img.stride(1) == 1) or (img.dim() == 3 and img.stride(2) == 1 and img.stride(1) == c)
            if not ok:
                img = img.contiguous()
            self.keep, self.ptr, self.pitch = img, img.data_ptr(), img.stride(0)
            self.on_device, self.device = True, img.device.index
        else:
            a = np.asarray(img)
            if a.dtype != np.uint8:
                a = _exact_u8_array(a, what)
            img = a
            if a.ndim not in (2, 3):
                raise ValueError(f"{what}: expected (H, W) or (H, W, C), got {a.shape}")
            c = 1 if a.ndim == 2 else a.shape[2]
            ok = (a.ndim == 2 and a.strides[1] == 1) or (a.ndim == 3 and a.strides[2] == 1 and a.strides[1] == c)
            if not ok or a.strides[0] < a.shape[1] * c:  # e.g. channel-reversed or transposed views
                a = np.ascontiguousarray(a)
            self.keep, self.ptr, self.pitch = a, a.ctypes.data, a.strides[0] if a.shape[0] > 1 else a.shape[1] * c
            self.on_device, self.device = False, None
        self.h, self.w = int(img.shape[0]), int(img.shape[1])
        self.c = int(c)
        if self.c not in (1, 3):
            raise ValueError(f"{what}: 1 or 3 channels supported, got {self.c}")
