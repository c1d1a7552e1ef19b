"""oracle/pyr.py -- TEST INFRASTRUCTURE.  pyrDown / pyrUp for the oracle.

Two interchangeable backends with identical bits (tests/test_oracle_pyr.py asserts it):
  "c"   : oracle/pyr_oracle.c (plain C restatement of OpenCV's float32 pyramids, compiled by oracle/Makefile)
  "cv2" : the reference's own third-party dependency, called exactly as the reference calls it
          (/root/reference/blending_module.py:263, 306, 358)
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    """Compile the C restatement in place (gcc); returns the .so path."""
    so = os.path.join(_HERE, "libsrf_oracle.so")
    src = os.path.join(_HERE, "pyr_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libsrf_oracle.so"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build())
        vp, ci = ctypes.c_void_p, ctypes.c_int
        lib.orc_pyr_down_f32.argtypes = [vp, ci, ci, ci, vp]
        lib.orc_pyr_down_f32.restype = ci
        lib.orc_pyr_up_f32.argtypes = [vp, ci, ci, ci, vp, ci, ci]
        lib.orc_pyr_up_f32.restype = ci
        lib.orc_down_vec_end.argtypes = [ci, ci, ci]
        lib.orc_down_vec_end.restype = ci
        _LIB = lib
    return _LIB


def have_cv2() -> bool:
    try:
        import cv2  # noqa: F401
        return True
    except Exception:
        return False


def _geom(img: np.ndarray):
    if img.dtype != np.float32 or img.ndim not in (2, 3):
        raise TypeError("oracle pyramids take float32 HxW or HxWxC arrays")
    h, w = img.shape[:2]
    cn = 1 if img.ndim == 2 else img.shape[2]
    return h, w, cn


def pyr_down(img: np.ndarray, backend: str = "c") -> np.ndarray:
    h, w, cn = _geom(img)
    if backend == "cv2":
        import cv2
        return cv2.pyrDown(img)
    src = np.ascontiguousarray(img)
    out = np.empty(((h + 1) // 2, (w + 1) // 2) + img.shape[2:], np.float32)
    rc = _lib().orc_pyr_down_f32(src.ctypes.data, h, w, cn, out.ctypes.data)
    if rc:
        raise RuntimeError(f"orc_pyr_down_f32 rc={rc}")
    return out


def pyr_up(img: np.ndarray, dst_hw, backend: str = "c") -> np.ndarray:
    h, w, cn = _geom(img)
    dh, dw = int(dst_hw[0]), int(dst_hw[1])
    if backend == "cv2":
        import cv2
        return cv2.pyrUp(img, dstsize=(dw, dh))
    src = np.ascontiguousarray(img)
    out = np.empty((dh, dw) + img.shape[2:], np.float32)
    rc = _lib().orc_pyr_up_f32(src.ctypes.data, h, w, cn, out.ctypes.data, dh, dw)
    if rc:
        raise ValueError(f"orc_pyr_up_f32 rc={rc} (dst {dh}x{dw} from {h}x{w})")
    return out


def down_vec_end(w: int, ow: int, cn: int) -> int:
    return int(_lib().orc_down_vec_end(w, ow, cn))
