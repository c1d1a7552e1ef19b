"""ctypes binding of libsrfuse.so (include/srfuse.h).  No CPU fallback: if the library or a usable device is
missing, every entry point raises."""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

from . import build as _build

c_void_p, c_int, c_size_t, c_int64 = C.c_void_p, C.c_int, C.c_size_t, C.c_int64


class SrfError(RuntimeError):
    pass


class srf_tile(C.Structure):
    _fields_ = [("data", c_void_p), ("pitch", c_size_t), ("h", c_int), ("w", c_int), ("y", c_int), ("x", c_int),
                ("lut", c_int)]


class srf_weight_lut(C.Structure):
    _fields_ = [("values", c_void_p), ("len", c_int)]


class srf_merge_tile(C.Structure):
    _fields_ = [("data", c_void_p), ("pitch", c_size_t), ("h", c_int), ("w", c_int), ("out_h", c_int), ("out_w", c_int),
                ("y", c_int), ("x", c_int), ("wy", c_void_p), ("wl", c_void_p), ("wr", c_void_p)]


class srf_wavg_tile(C.Structure):
    _fields_ = [("data", c_void_p), ("pitch", c_size_t), ("h", c_int), ("w", c_int), ("y", c_int), ("x", c_int),
                ("weight", c_void_p), ("weight_pitch", c_size_t), ("weight_h", c_int), ("weight_w", c_int),
                ("weight_f64", c_int), ("reserved", c_int)]


class srf_image_pair(C.Structure):
    _fields_ = [("a", c_void_p), ("b", c_void_p), ("pitch_a", c_size_t), ("pitch_b", c_size_t)]


class srf_profile_entry(C.Structure):
    _fields_ = [("name", C.c_char * 40), ("ms", C.c_float), ("bytes", C.c_double)]


class srf_iqa_sums(C.Structure):
    _fields_ = [("sq_err", c_int64), ("n_elems", c_int64), ("ssim_sum", C.c_double), ("ssim_count", c_int64),
                ("bits_a", C.c_int32), ("bits_b", C.c_int32)]


# every symbol include/srfuse.h declares: name -> (restype, argtypes)
P = C.POINTER
SIGNATURES = {
    "srf_version": (c_int, []),
    "srf_status_string": (C.c_char_p, [c_int]),
    "srf_last_error": (C.c_char_p, [c_void_p]),
    "srf_create": (c_int, [c_int, P(c_void_p)]),
    "srf_destroy": (None, [c_void_p]),
    "srf_sync": (c_int, [c_void_p, c_void_p]),
    "srf_launch_count": (c_int64, [c_void_p]),
    "srf_graph_replays": (c_int64, [c_void_p]),
    "srf_workspace_bytes": (c_size_t, [c_void_p]),
    "srf_last_transfer_bytes": (c_int, [c_void_p, P(c_size_t), P(c_size_t)]),
    "srf_last_device_ms": (C.c_float, [c_void_p]),
    "srf_profile_enable": (c_int, [c_void_p, c_int]),
    "srf_profile_read": (c_int, [c_void_p, P(srf_profile_entry), c_int]),
    "srf_tile_positions": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_int]),
    "srf_tile_overlaps": (c_int, [c_int] * 8 + [c_void_p]),
    "srf_tile_extract_pad": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_size_t, c_void_p, c_int, c_int, c_int,
                                     c_void_p, c_void_p]),
    "srf_tile_extract_pad_host": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_size_t, c_void_p, c_int, c_int,
                                          c_int, c_void_p]),
    "srf_copy_make_border": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_size_t, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                     c_void_p]),
    "srf_copy_make_border_host": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_size_t, c_int, c_int, c_int, c_int, c_int,
                                          c_void_p]),
    "srf_pyr_down_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "srf_pyr_up_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p]),
    "srf_pyr_down_f32_host": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "srf_pyr_up_f32_host": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int]),
    "srf_weight_profile": (c_int, [c_int, c_int, c_void_p]),
    "srf_laplacian_fuse": (c_int, [c_void_p, P(srf_tile), c_int, c_int, c_int, P(srf_weight_lut), c_int, c_int, c_int,
                                   c_void_p, c_size_t, c_int, c_int, c_void_p]),
    "srf_laplacian_fuse_host": (c_int, [c_void_p, P(srf_tile), c_int, c_int, c_int, P(srf_weight_lut), c_int, c_int,
                                        c_int, c_void_p, c_size_t, c_int, c_int]),
    "srf_fuse_input_rows": (c_int, [c_int] * 9 + [P(c_int), P(c_int), P(c_int)]),
    "srf_gather_plan": (c_int, [P(srf_tile), c_int, c_int, c_int, P(srf_weight_lut), c_int, c_int, c_int, c_int, c_int, P(c_int), c_int, P(c_int)]),
    "srf_psnr_ssim": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_int, c_int, c_int, c_int, c_int,
                              C.c_double, c_int, c_int, c_int, P(srf_iqa_sums), c_void_p]),
    "srf_psnr_ssim_enqueue": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_int, c_int, c_int, c_int, c_int,
                                      C.c_double, c_int, c_int, c_int, c_void_p, c_void_p]),
    "srf_psnr_ssim_host": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_int, c_int, c_int, c_int, c_int,
                                   C.c_double, c_int, c_int, c_int, P(srf_iqa_sums)]),
    "srf_psnr_ssim_batch": (c_int, [c_void_p, P(srf_image_pair), c_int, c_int, c_int, c_int, C.c_double, c_int, P(srf_iqa_sums),
                                    c_void_p]),
    "srf_psnr_ssim_batch_enqueue": (c_int, [c_void_p, P(srf_image_pair), c_int, c_int, c_int, c_int, C.c_double, c_int, c_void_p,
                                            c_void_p]),
    "srf_psnr_ssim_batch_host": (c_int, [c_void_p, P(srf_image_pair), c_int, c_int, c_int, c_int, C.c_double, c_int,
                                         P(srf_iqa_sums)]),
    "srf_fuse_assess_host": (c_int, [c_void_p, P(srf_tile), c_int, c_int, c_int, P(srf_weight_lut), c_int, c_int, c_int,
                                     c_void_p, c_size_t, c_void_p, c_size_t, C.c_double, c_int, c_int, P(srf_iqa_sums)]),
    "srf_comm_handle": (c_int, [c_void_p, c_void_p]),
    "srf_comm_init": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "srf_allreduce_metrics": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "srf_comm_result": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "srf_iqa_finish": (c_int, [P(srf_iqa_sums), C.c_double, P(C.c_double), P(C.c_double)]),
    "srf_window_ssim_shape": (c_int, [c_int] * 8 + [P(c_int), P(c_int)]),
    "srf_window_ssim": (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_int, c_void_p, c_size_t, c_int, c_int, c_int, c_int,
                                c_int, c_int, c_int, C.c_double, C.c_double, c_void_p, c_void_p]),
    "srf_window_ssim_host": (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_int, c_void_p, c_size_t, c_int, c_int, c_int,
                                     c_int, c_int, c_int, c_int, C.c_double, C.c_double, c_void_p]),
    "srf_merge_tiles": (c_int, [c_void_p, P(srf_merge_tile), c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "srf_merge_tiles_host": (c_int, [c_void_p, P(srf_merge_tile), c_int, c_int, c_int, c_int, c_void_p, c_size_t]),
    "srf_weighted_average": (c_int, [c_void_p, P(srf_wavg_tile), c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "srf_weighted_average_host": (c_int, [c_void_p, P(srf_wavg_tile), c_int, c_int, c_int, c_int, c_void_p, c_size_t]),
    "srf_resize_linear_taps": (c_int, [c_int, c_int, c_void_p, c_void_p]),
    "srf_resize_linear_taps_axis": (c_int, [c_int, c_int, c_int, c_void_p, c_void_p]),
    "srf_resize_cubic_taps": (c_int, [c_int, c_int, c_void_p, c_void_p]),
    "srf_resize_cubic_u8": (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_int, c_int, c_void_p, c_size_t, c_int, c_int,
                                    c_void_p]),
    "srf_resize_cubic_u8_host": (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_int, c_int, c_void_p, c_size_t, c_int,
                                         c_int]),
}

_lib = None
_lock = threading.RLock()
_ctx = {}
_ctx_locks = {}  # srf_ctx address -> RLock: calls on one ctx must not overlap (include/srfuse.h), and ctypes drops the GIL

# entry points that do NOT take an srf_ctx* first (nothing to serialise)
_NO_CTX = {"srf_fuse_input_rows", "srf_gather_plan", "srf_version", "srf_status_string", "srf_last_error", "srf_create", "srf_tile_positions", "srf_tile_overlaps",
           "srf_weight_profile", "srf_iqa_finish", "srf_window_ssim_shape", "srf_resize_linear_taps", "srf_resize_linear_taps_axis", "srf_resize_cubic_taps"}


def lib_path() -> str:
    return os.environ.get("SRF_LIB") or _build.LIB_PATH


def _guarded(fn):
    """serialise every call on one srf_ctx: SuperResolutionPipeline.process runs run() on executor threads, and two
    threads inside the library on the same workspaces would corrupt each other's staging buffers and descriptors"""
    def call(ctx, *args):
        lk = _ctx_locks.get(getattr(ctx, "value", ctx))
        if lk is None:
            return fn(ctx, *args)
        with lk:
            return fn(ctx, *args)
    call.__name__ = fn.__name__
    return call


class _Library:
    """the bound entry points of libsrfuse.so; those that take an srf_ctx hold that context's lock for the call"""

    def __init__(self, cdll):
        self._cdll = cdll
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(cdll, name)  # AttributeError if the export is missing
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, fn if name in _NO_CTX else _guarded(fn))


def load():
    """dlopen libsrfuse.so and bind every declared symbol; raises if the library has not been built."""
    global _lib
    with _lock:
        if _lib is None:
            path = lib_path()
            if not os.path.exists(path):
                raise SrfError(f"{path} is missing -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(libsrfuse has no CPU fallback)")
            _lib = _Library(C.CDLL(path))
    return _lib


def last_error(ctx=None) -> str:
    msg = load().srf_last_error(ctx)
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, ctx=None, exc=None):
    if rc >= 0:
        return rc
    text = last_error(ctx) or load().srf_status_string(rc).decode()
    if rc == -1:
        raise (exc or ValueError)(text)
    if rc == -4:
        raise MemoryError(text)
    if rc == -5:
        raise NotImplementedError(text)
    raise SrfError(f"libsrfuse status {rc}: {text}")


def context(device: int = 0, slot: int = 0):
    """one srf_ctx per (process, device, slot); raises SrfError when no sm_100 device is usable.  Slot 0 is the context
    every mirror class uses; further slots are independent contexts on the same device (own workspaces, own captured
    graph) for callers that keep two fusions in flight (distributed.PipelinedFuseAssess)."""
    lib = load()
    key = device if slot == 0 else (device, int(slot))
    with _lock:
        if key not in _ctx:
            h = c_void_p()
            rc = lib.srf_create(device, C.byref(h))
            if rc != 0:
                raise SrfError(f"srf_create(device={device}) failed ({rc}): {last_error(None)}")
            _ctx[key] = h
            _ctx_locks[h.value] = threading.RLock()
    return _ctx[key]


def ctx_lock(ctx):
    """the lock that serialises library calls on `ctx`: hold it across a call + the read of its side results
    (last_transfer_bytes, last_device_ms) when several threads share the context"""
    return _ctx_locks[getattr(ctx, "value", ctx)]


def destroy_contexts():
    lib = load()
    with _lock:
        for h in _ctx.values():
            lib.srf_destroy(h)
        _ctx.clear()
        _ctx_locks.clear()


def last_transfer_bytes(ctx):
    a, b = c_size_t(), c_size_t()
    check(load().srf_last_transfer_bytes(ctx, C.byref(a), C.byref(b)), ctx)
    return int(a.value), int(b.value)


def profile_enable(ctx, on: bool) -> None:
    check(load().srf_profile_enable(ctx, int(on)), ctx)


def profile_read(ctx, cap: int = 4096):
    """[(kernel name, ms, algorithmic bytes)] in launch order since the last read"""
    buf = (srf_profile_entry * cap)()
    n = check(load().srf_profile_read(ctx, buf, cap), ctx)
    return [(buf[i].name.decode(), float(buf[i].ms), float(buf[i].bytes)) for i in range(min(n, cap))]
