"""oracle/ref_arm.py -- TEST / BASELINE INFRASTRUCTURE: run the REFERENCE's own modules (oracle/_ref, see make_ref.py).

bench.py's `--impl reference` arm and `cpu_baseline` leg time these calls; nothing in the product package imports
this file.  The modules are the unmodified reference sources:
    BlendingModule.laplacian_fusion          /root/reference/blending_module.py:369-506
    QualityAssessmentModule.calculate_psnr   /root/reference/quality_assessment_module.py:277-320
    QualityAssessmentModule.calculate_ssim   /root/reference/quality_assessment_module.py:322-389
quality_assessment_module imports `lpips` and `skimage` at module level; neither is in this image (requirements.txt
of the reference leaves them to pip), so -- exactly as tests/golden/make_golden.py does -- they are stubbed with
objects that RAISE when called.  The reference then takes the fallbacks it ships for that case: the manual PSNR formula
(313-318) and `_calculate_ssim_simple` (391-417, cv2.GaussianBlur in float64).  That is what the reference computes on
this box; if scikit-image is importable the real one is used and no stub is installed.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import logging
import os
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_mods = None


def available() -> bool:
    from . import make_ref
    return make_ref.available()


def _stub_missing():
    stubs = []
    try:
        import lpips  # noqa: F401
    except Exception:
        lp = types.ModuleType("lpips")

        class _NoLpips:
            def __init__(self, *a, **k):
                raise RuntimeError("lpips absent")

        lp.LPIPS = _NoLpips
        sys.modules["lpips"] = lp
        stubs.append("lpips")
    try:
        import skimage.metrics  # noqa: F401
    except Exception:
        sk, skm = types.ModuleType("skimage"), types.ModuleType("skimage.metrics")

        def _absent(*a, **k):
            raise RuntimeError("skimage absent")

        skm.peak_signal_noise_ratio = skm.structural_similarity = _absent
        sk.metrics = skm
        sys.modules["skimage"], sys.modules["skimage.metrics"] = sk, skm
        stubs.append("skimage")
    return stubs


def load():
    """import the reference modules from oracle/_ref -> (blending_module, quality_assessment_module, tiling_module, stubs)"""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError("oracle/_ref is missing: run `python oracle/make_ref.py` where /root/reference exists")
    stubs = _stub_missing()
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp(prefix="srf_ref_"))  # the reference modules call logging.basicConfig / write logs into the cwd
    sys.path.insert(0, REF_DIR)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            rb = importlib.import_module("blending_module")
            rq = importlib.import_module("quality_assessment_module")
            rt = importlib.import_module("tiling_module")
    finally:
        sys.path.remove(REF_DIR)
        os.chdir(cwd)
    for m in (rb, rq, rt):
        if os.path.dirname(os.path.abspath(m.__file__)) != REF_DIR:
            raise RuntimeError(f"{m.__name__} resolved to {m.__file__}, not to oracle/_ref (a same-named module was imported first)")
    logging.getLogger().setLevel(logging.WARNING)
    _mods = (rb, rq, rt, stubs)
    return _mods


def laplacian_fuse(tiles, positions_yx, canvas_hw, levels: int):
    """the reference's BlendingModule.laplacian_fusion on TileInfo objects (cosine weights, the default)"""
    rb = load()[0]
    bm = rb.BlendingModule(method="laplacian", num_levels=int(levels))
    infos = [rb.TileInfo(image=t, x=int(x), y=int(y), row=0, col=0) for t, (y, x) in zip(tiles, positions_yx)]
    return bm.laplacian_fusion(infos, output_shape=(int(canvas_hw[0]), int(canvas_hw[1])))


_qa = None


def _quality_module():
    global _qa
    if _qa is None:
        rq = load()[1]
        with contextlib.redirect_stdout(io.StringIO()):
            _qa = rq.QualityAssessmentModule(device="cpu")
    return _qa


def psnr(a, b) -> float:
    return _quality_module().calculate_psnr(a, b)


def ssim(a, b) -> float:
    return _quality_module().calculate_ssim(a, b, multiscale=True)


def describe() -> str:
    stubs = load()[3]
    return ("reference modules from oracle/_ref (unmodified): BlendingModule.laplacian_fusion + "
            "QualityAssessmentModule.calculate_psnr / calculate_ssim(multiscale=True)"
            + (f"; absent third-party packages stubbed to raise ({', '.join(stubs)}) -> the reference's own fallbacks run" if stubs else ""))
