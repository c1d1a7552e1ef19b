"""GPU mirror of the reference's full-reference metrics: PSNR and 11x11-Gaussian SSIM in one pass.

Mirrors /root/reference/quality_assessment_module.py: AssessmentLevel 35-41, QualityThresholds 44-75,
ScaleConfig 78-86, QualityAssessmentModule.__init__ 107-133, _preprocess_image 169-195, calculate_psnr 277-320,
downsample_bicubic 226-253, upsample_bicubic 255-275, calculate_ssim 322-389, _calculate_ssim_simple 391-417,
evaluate_full_reference 467-516,
_evaluate_downsample_comparison 518-555, _assess_* 557-588, _calculate_overall_score 590-609, batch_evaluate 1350-1371.
The arithmetic (scikit-image's peak_signal_noise_ratio / structural_similarity) runs in csrc/srf_iqa.cu (both SSIM
windows: 11x11 Gaussian and scikit-image's default 7x7 uniform); cv2.resize(INTER_CUBIC) runs in csrc/srf_resize.cu.
Out of scope: LPIPS, NIQE, BRISQUE, commercial metrics, reports.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from enum import Enum
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _native
from ._arrays import ImageView, current_stream_ptr, is_cuda_tensor, torch


class AssessmentLevel(Enum):
    EXCELLENT = "excellent"
    GOOD = "good"
    FAIR = "fair"
    POOR = "poor"
    BAD = "bad"


@dataclass
class QualityThresholds:
    PSNR_EXCELLENT: float = 40.0
    PSNR_GOOD: float = 35.0
    PSNR_FAIR: float = 30.0
    SSIM_EXCELLENT: float = 0.98
    SSIM_GOOD: float = 0.95
    SSIM_FAIR: float = 0.90
    LPIPS_EXCELLENT: float = 0.02
    LPIPS_GOOD: float = 0.05
    LPIPS_FAIR: float = 0.10


@dataclass
class ScaleConfig:
    scale_factors: List[float] = field(default_factory=lambda: [0.1, 0.2, 0.4])
    scale_names: Dict[float, str] = field(default_factory=lambda: {0.1: "structure_color", 0.2: "mid_frequency",
                                                                   0.4: "high_frequency"})


class QualityAssessmentModule:
    def __init__(self, device: str = "cpu", thresholds: Optional[QualityThresholds] = None,
                 scale_config: Optional[ScaleConfig] = None):
        # the reference's `device` picks where LPIPS runs; PSNR/SSIM here always run on a B200.  'cuda:N' selects N.
        self.device = device
        self.device_index = int(device.split(":")[1]) if isinstance(device, str) and ":" in device else 0
        self.thresholds = thresholds or QualityThresholds()
        self.scale_config = scale_config or ScaleConfig()
        self.lpips_model_vgg = None  # LPIPS is out of scope; evaluate_full_reference omits its keys exactly as the
        self.lpips_model_alex = None  # reference does when the models fail to load (q.a.m:508-511)

    # ------------------------------------------------------------------ preprocessing
    def _preprocess_image(self, image, to_tensor: bool = False):
        """tensor (C,H,W)/(1,C,H,W) -> HWC; float images in [0, 1] -> uint8 by *255 and truncation.  uint8 images are
        returned untouched: their `max() <= 1` case is resolved inside the kernel pass (srf_iqa_sums.bits_a/bits_b)."""
        if is_cuda_tensor(image) and image.dtype == torch.uint8:
            return image
        if torch is not None and isinstance(image, torch.Tensor):
            if image.dim() == 4:
                image = image.squeeze(0)
            if image.dim() == 3:
                image = image.permute(1, 2, 0)
            image = image.cpu().numpy()
        image = np.asarray(image)
        if image.dtype != np.uint8:
            if image.max() <= 1.0:
                image = (image * 255).astype(np.uint8)
            else:
                raise TypeError("calculate_psnr/calculate_ssim: images must be uint8 or float in [0, 1]")
        return image

    @staticmethod
    def _crop_common(a, b):
        if tuple(a.shape) != tuple(b.shape):
            h, w = min(a.shape[0], b.shape[0]), min(a.shape[1], b.shape[1])
            a, b = a[:h, :w], b[:h, :w]
        return a, b

    # ------------------------------------------------------------------ the one-pass kernel
    def psnr_ssim_sums(self, img1, img2, data_range: float = 255.0, row_range: Optional[Tuple[int, int]] = None,
                       window: int = 0):
        """additive partial sums (srf_iqa_sums) of the pair over a row band -- what ranks all-reduce.
        window: 0 = 11x11 Gaussian (multiscale=True), 1 = 7x7 uniform with sample covariance (multiscale=False)"""
        a, b = self._crop_common(self._preprocess_image(img1), self._preprocess_image(img2))
        if len(a.shape) != len(b.shape) or (len(a.shape) == 3 and a.shape[2] != b.shape[2]):
            raise ValueError("image pair disagrees on channels")
        va, vb = ImageView(a, "img1"), ImageView(b, "img2")
        if va.on_device != vb.on_device:
            raise ValueError("image pair must be both host or both device")
        h, w = va.h, va.w
        r0, r1 = (0, h) if row_range is None else (int(row_range[0]), int(row_range[1]))
        lib, ctx = _native.load(), _native.context(va.device if va.on_device else self.device_index)
        sums = _native.srf_iqa_sums()

        def run(sa, sb):
            if va.on_device:
                rc = lib.srf_psnr_ssim(ctx, va.ptr, va.pitch, vb.ptr, vb.pitch, h, w, va.c, sa, sb, float(data_range),
                                       int(window), r0, r1, C.byref(sums), current_stream_ptr(va.device))
            else:
                rc = lib.srf_psnr_ssim_host(ctx, va.ptr, va.pitch, vb.ptr, vb.pitch, h, w, va.c, sa, sb,
                                            float(data_range), int(window), r0, r1, C.byref(sums))
            _native.check(rc, ctx)

        run(0, 0)
        if row_range is None and (sums.bits_a <= 1 or sums.bits_b <= 1):
            # the reference rescales an image whose max() <= 1.0 by 255 (q.a.m:192-193) -- also an all-0/1 uint8 one
            run(int(sums.bits_a <= 1), int(sums.bits_b <= 1))
        return sums

    def psnr_ssim_sums_device(self, img1, img2, out, data_range: float = 255.0, row_range=None):
        """Enqueue-only form for CUDA tensors (extension): the band's additive sums {sq_err, ssim_sum, ssim_count,
        any byte of img1 > 1, any byte of img2 > 1} are written into `out` (float64 CUDA tensor, 5 elements) on the
        current stream, with no host synchronisation -- all-reduce `out` and read it back once
        (see distributed.allreduce_device_sums)."""
        va, vb = ImageView(img1, "img1"), ImageView(img2, "img2")
        if not (va.on_device and vb.on_device) or (va.h, va.w, va.c) != (vb.h, vb.w, vb.c):
            raise ValueError("psnr_ssim_sums_device takes two CUDA uint8 images of equal shape")
        if out.dtype != torch.float64 or out.numel() < 5 or not out.is_cuda:
            raise ValueError("out must be a float64 CUDA tensor with 5 elements")
        h, w = va.h, va.w
        r0, r1 = (0, h) if row_range is None else (int(row_range[0]), int(row_range[1]))
        lib, ctx = _native.load(), _native.context(va.device)
        rc = lib.srf_psnr_ssim_enqueue(ctx, va.ptr, va.pitch, vb.ptr, vb.pitch, h, w, va.c, 0, 0, float(data_range), 0, r0, r1,
                                       out.data_ptr(), current_stream_ptr(va.device))
        _native.check(rc, ctx)
        return out

    def psnr_ssim_sums_device_rows(self, rows1, rows2, out, image_rows: int, first_row: int, data_range: float = 255.0,
                                   row_range=None):
        """psnr_ssim_sums_device for a rank that holds only a ROW SLICE of a tall pair: `rows1` / `rows2` are rows
        [first_row, first_row + k) of two images with `image_rows` rows; `row_range` (image coordinates) must lie inside
        the slice minus the 5-row window halo (except at the image's own top / bottom edge)."""
        va, vb = ImageView(rows1, "rows1"), ImageView(rows2, "rows2")
        if not (va.on_device and vb.on_device) or (va.h, va.w, va.c) != (vb.h, vb.w, vb.c):
            raise ValueError("psnr_ssim_sums_device_rows takes two CUDA uint8 slices of equal shape")
        r0, r1 = (first_row, first_row + va.h) if row_range is None else (int(row_range[0]), int(row_range[1]))
        lo, hi = max(r0 - 5, 0), min(r1 + 5, int(image_rows))
        if r1 > r0 and (lo < first_row or hi > first_row + va.h):
            raise ValueError(f"rows [{r0}, {r1}) +- 5 are not inside the slice [{first_row}, {first_row + va.h})")
        lib, ctx = _native.load(), _native.context(va.device)
        # the kernel addresses rows of the whole image: shift the base pointers up by the rows the slice does not hold
        rc = lib.srf_psnr_ssim_enqueue(ctx, va.ptr - first_row * va.pitch, va.pitch, vb.ptr - first_row * vb.pitch, vb.pitch, int(image_rows),
                                       va.w, va.c, 0, 0, float(data_range), 0, r0, r1, out.data_ptr(), current_stream_ptr(va.device))
        _native.check(rc, ctx)
        return out

    @staticmethod
    def finish(sums, data_range: float = 255.0) -> Tuple[float, float]:
        p, s = C.c_double(), C.c_double()
        _native.check(_native.load().srf_iqa_finish(C.byref(sums), float(data_range), C.byref(p), C.byref(s)))
        return p.value, s.value

    # ------------------------------------------------------------------ reference API
    def calculate_psnr(self, img1, img2, data_range: float = 255.0) -> float:
        return float(self.finish(self.psnr_ssim_sums(img1, img2, data_range), data_range)[0])

    def calculate_ssim(self, img1, img2, multiscale: bool = True, data_range: float = 255.0) -> float:
        # multiscale=True: 11x11 Gaussian window (the reference's "ms_ssim" is single-scale, q.a.m:366-376);
        # multiscale=False: scikit-image's default 7x7 uniform window with sample covariance (q.a.m:378-384)
        sums = self.psnr_ssim_sums(img1, img2, data_range, window=0 if multiscale else 1)
        if sums.ssim_count <= 0:
            raise ValueError("image smaller than the SSIM window")
        return float(self.finish(sums, data_range)[1])

    def _calculate_ssim_simple(self, img1, img2) -> float:
        """The reference's fallback SSIM (q.a.m:391-417; what calculate_ssim returns wherever scikit-image is missing):
        11x11 Gaussian filters with BORDER_REFLECT_101 and the mean over the FULL map.  Filtering with that border equals
        filtering the image reflect-padded by the window radius, so this is the interior mean (the kernel's own) of the
        padded pair -- the padding is a host / torch copy, the arithmetic stays in k_psnr_ssim."""
        a, b = self._crop_common(self._preprocess_image(img1), self._preprocess_image(img2))
        h, w = int(a.shape[0]), int(a.shape[1])
        if min(h, w) <= 5:
            raise ValueError("image smaller than the window radius")
        if is_cuda_tensor(a) != is_cuda_tensor(b):
            raise ValueError("image pair must be both host or both device")
        if is_cuda_tensor(a):
            iy = torch.tensor([abs(i) if i < h else 2 * (h - 1) - i for i in range(-5, h + 5)], device=a.device)
            ix = torch.tensor([abs(i) if i < w else 2 * (w - 1) - i for i in range(-5, w + 5)], device=a.device)
            a, b = (t.index_select(0, iy).index_select(1, ix).contiguous() for t in (a, b))
        else:
            pad = ((5, 5), (5, 5)) + (((0, 0),) if a.ndim == 3 else ())
            a, b = np.pad(a, pad, mode="reflect"), np.pad(b, pad, mode="reflect")
        return self.calculate_ssim(a, b, multiscale=True)

    def calculate_psnr_ssim(self, img1, img2, data_range: float = 255.0) -> Tuple[float, float]:
        """both metrics from ONE pass over the pair (extension; the reference makes two)"""
        sums = self.psnr_ssim_sums(img1, img2, data_range)
        if sums.ssim_count <= 0:
            raise ValueError("image smaller than the 11x11 SSIM window")
        return self.finish(sums, data_range)

    def _level(self, value, exc, good, fair) -> str:
        if value >= exc:
            return AssessmentLevel.EXCELLENT.value
        if value >= good:
            return AssessmentLevel.GOOD.value
        if value >= fair:
            return AssessmentLevel.FAIR.value
        return AssessmentLevel.POOR.value

    def _assess_psnr(self, v: float) -> str:
        t = self.thresholds
        return self._level(v, t.PSNR_EXCELLENT, t.PSNR_GOOD, t.PSNR_FAIR)

    def _assess_ssim(self, v: float) -> str:
        t = self.thresholds
        return self._level(v, t.SSIM_EXCELLENT, t.SSIM_GOOD, t.SSIM_FAIR)

    def _calculate_overall_score(self, metrics: Dict[str, float]) -> float:
        scores = []
        if "psnr" in metrics:
            scores.append(min(100, max(0, metrics["psnr"])))
        if "ms_ssim" in metrics:
            scores.append(metrics["ms_ssim"] * 100)
        if "lpips_vgg" in metrics:
            scores.append(max(0, (1 - metrics["lpips_vgg"]) * 100))
        return float(np.mean(scores)) if scores else 0.0

    # ------------------------------------------------------------------ bicubic resize (cv2.resize INTER_CUBIC)
    def _resize_cubic(self, image, dst_w: int, dst_h: int):
        v = ImageView(image, "image")
        if dst_w < 1 or dst_h < 1:
            raise ValueError(f"resize to an empty size ({dst_w}x{dst_h})")  # cv2.resize raises on an empty dsize
        two_d = len(image.shape) == 2
        shape = (dst_h, dst_w) if two_d else (dst_h, dst_w, v.c)
        lib, ctx = _native.load(), _native.context(v.device if v.on_device else self.device_index)
        if v.on_device:
            out = torch.empty(shape, dtype=torch.uint8, device=v.keep.device)
            rc = lib.srf_resize_cubic_u8(ctx, v.ptr, v.pitch, v.h, v.w, v.c, out.data_ptr(), dst_w * v.c, dst_h, dst_w,
                                         current_stream_ptr(v.device))
        else:
            out = np.empty(shape, np.uint8)
            rc = lib.srf_resize_cubic_u8_host(ctx, v.ptr, v.pitch, v.h, v.w, v.c, out.ctypes.data, dst_w * v.c, dst_h, dst_w)
        _native.check(rc, ctx)
        return out

    def downsample_bicubic(self, image, scale_factor: float):
        """cv2.resize(image, (int(w*s), int(h*s)), INTER_CUBIC) (q.a.m:226-253); uint8 NumPy array or CUDA tensor"""
        if scale_factor >= 1.0 or scale_factor <= 0:
            raise ValueError(f"scale_factor must lie in (0, 1), got {scale_factor}")
        h, w = int(image.shape[0]), int(image.shape[1])
        return self._resize_cubic(image, int(w * scale_factor), int(h * scale_factor))

    def upsample_bicubic(self, image, target_size: Tuple[int, int]):
        """cv2.resize(image, (target_size[1], target_size[0]), INTER_CUBIC) (q.a.m:255-275)"""
        return self._resize_cubic(image, int(target_size[1]), int(target_size[0]))

    def _evaluate_downsample_comparison(self, original, upscaled, scale_factor: int) -> Dict[str, float]:
        """both images down-sampled to each scale of scale_config, then PSNR and the 7x7-window SSIM (q.a.m:518-555)"""
        metrics = {}
        for scale in self.scale_config.scale_factors:
            name = self.scale_config.scale_names.get(scale, f"scale_{scale}")
            sr, hr = self.downsample_bicubic(upscaled, scale), self.downsample_bicubic(original, scale)
            mh, mw = min(sr.shape[0], hr.shape[0]), min(sr.shape[1], hr.shape[1])
            sr, hr = sr[:mh, :mw], hr[:mh, :mw]
            metrics[f"psnr_{name}"] = self.calculate_psnr(hr, sr)
            metrics[f"ssim_{name}"] = self.calculate_ssim(hr, sr, multiscale=False)
        return metrics

    def evaluate_full_reference(self, original, upscaled, scale_factor: int = 4) -> Dict[str, float]:
        """the reference's dict (q.a.m:467-516): psnr_/ssim_ of the three down-sampled pairs, psnr, psnr_level, ssim (7x7
        uniform window), ms_ssim (= single-scale 11x11 Gaussian SSIM, the reference's naming), ssim_level, overall_score.
        LPIPS keys are omitted as in the reference when its models are unavailable.  Host images are uploaded ONCE and
        every kernel then runs on the device copies."""
        a, b = self._preprocess_image(original), self._preprocess_image(upscaled)
        if not is_cuda_tensor(a) and not is_cuda_tensor(b):
            if torch is None or not torch.cuda.is_available():
                raise _native.SrfError("evaluate_full_reference needs a CUDA device (there is no CPU fallback)")
            dev = torch.device("cuda", self.device_index)
            a = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
            b = torch.from_numpy(np.ascontiguousarray(b)).to(dev)
        metrics = dict(self._evaluate_downsample_comparison(a, b, scale_factor))
        psnr, ssim = self.calculate_psnr_ssim(a, b)
        metrics["psnr"] = float(psnr)
        metrics["psnr_level"] = self._assess_psnr(metrics["psnr"])
        metrics["ssim"] = self.calculate_ssim(a, b, multiscale=False)
        metrics["ms_ssim"] = float(ssim)
        metrics["ssim_level"] = self._assess_ssim(metrics["ms_ssim"])
        metrics["overall_score"] = self._calculate_overall_score(metrics)
        return metrics

    # ------------------------------------------------------------------ batched pairs
    def batch_psnr_ssim_sums(self, image_pairs, data_range: float = 255.0, window: int = 0, out=None):
        """srf_iqa_sums of MANY equally shaped pairs from ONE kernel launch (grid.z = pair): include/srfuse.h
        srf_psnr_ssim_batch[_host | _enqueue].  Pairs are (original, upscaled) uint8 images, all NumPy (host path: the
        library stages them) or all CUDA tensors.  With `out` (float64 CUDA tensor, 5 * n elements, device pairs only)
        nothing is synchronised and `out` receives 5 additive doubles per pair -- the form ranks all-reduce."""
        views = [(ImageView(a, f"pair {i} original"), ImageView(b, f"pair {i} upscaled")) for i, (a, b) in enumerate(image_pairs)]
        if not views:
            return []
        va0 = views[0][0]
        for va, vb in views:
            if (va.h, va.w, va.c) != (va0.h, va0.w, va0.c) or (vb.h, vb.w, vb.c) != (va0.h, va0.w, va0.c):
                raise ValueError("batch_psnr_ssim_sums: every image of the batch must have the same shape")
            if va.on_device != va0.on_device or vb.on_device != va0.on_device:
                raise ValueError("batch_psnr_ssim_sums: mix of host and device images")
        n = len(views)
        c_pairs = (_native.srf_image_pair * n)(*[_native.srf_image_pair(va.ptr, vb.ptr, va.pitch, vb.pitch) for va, vb in views])
        lib, ctx = _native.load(), _native.context(va0.device if va0.on_device else self.device_index)
        if out is not None:
            if not va0.on_device or out.dtype != torch.float64 or out.numel() < 5 * n or not out.is_cuda:
                raise ValueError("out needs device pairs and a float64 CUDA tensor with 5 elements per pair")
            rc = lib.srf_psnr_ssim_batch_enqueue(ctx, c_pairs, n, va0.h, va0.w, va0.c, float(data_range), int(window), out.data_ptr(),
                                                 current_stream_ptr(va0.device))
            _native.check(rc, ctx)
            return out
        sums = (_native.srf_iqa_sums * n)()
        if va0.on_device:
            rc = lib.srf_psnr_ssim_batch(ctx, c_pairs, n, va0.h, va0.w, va0.c, float(data_range), int(window), sums,
                                         current_stream_ptr(va0.device))
        else:
            rc = lib.srf_psnr_ssim_batch_host(ctx, c_pairs, n, va0.h, va0.w, va0.c, float(data_range), int(window), sums)
        _native.check(rc, ctx)
        return list(sums)

    def batch_evaluate(self, image_pairs, scale_factor: int = 4) -> List[Dict[str, float]]:
        """one evaluate_full_reference dict per pair, in order (q.a.m:1350-1371; exceptions propagate as in the
        reference's loop).  Pairs of equal shape are uploaded once and scored together: the 11x11-Gaussian pass (psnr,
        ms_ssim) and the 7x7 pass (ssim) are ONE batched launch each over all of them (srf_psnr_ssim_batch, grid.z =
        pair), and so is every scale of the down-sample comparison; odd-shaped pairs take the per-pair route."""
        pairs = list(image_pairs)
        results: List[Optional[Dict[str, float]]] = [None] * len(pairs)
        groups: Dict[tuple, List[int]] = {}
        prepared = {}
        for i, (o, u) in enumerate(pairs):
            a, b = self._crop_common(self._preprocess_image(o), self._preprocess_image(u))
            if is_cuda_tensor(a) != is_cuda_tensor(b):
                raise ValueError("image pair must be both host or both device")
            if not is_cuda_tensor(a):
                if torch is None or not torch.cuda.is_available():
                    raise _native.SrfError("batch_evaluate needs a CUDA device (there is no CPU fallback)")
                dev = torch.device("cuda", self.device_index)
                a = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
                b = torch.from_numpy(np.ascontiguousarray(b)).to(dev)
            prepared[i] = (a.contiguous(), b.contiguous())
            groups.setdefault((tuple(a.shape), a.device.index), []).append(i)

        def degenerate(sums):  # the max() <= 1 rescale rule or an image below the window size: per-pair route
            return any(s.bits_a <= 1 or s.bits_b <= 1 or s.ssim_count <= 0 for s in sums)

        for idx in groups.values():
            batch = [prepared[i] for i in idx]
            mets = None
            if len(idx) > 1:
                full = self.batch_psnr_ssim_sums(batch, window=0)
                uni = self.batch_psnr_ssim_sums(batch, window=1)
                if not degenerate(list(full) + list(uni)):
                    mets = [dict() for _ in idx]
                    for scale in self.scale_config.scale_factors:
                        name = self.scale_config.scale_names.get(scale, f"scale_{scale}")
                        small = [(self.downsample_bicubic(b, scale), self.downsample_bicubic(a, scale)) for a, b in batch]
                        small = [(hr, sr) for sr, hr in small]
                        sg, su = self.batch_psnr_ssim_sums(small, window=0), self.batch_psnr_ssim_sums(small, window=1)
                        if degenerate(list(sg) + list(su)):
                            mets = None
                            break
                        for m, g1, u1 in zip(mets, sg, su):
                            m[f"psnr_{name}"] = float(self.finish(g1)[0])
                            m[f"ssim_{name}"] = float(self.finish(u1)[1])
            if mets is None:
                for k in idx:
                    results[k] = self.evaluate_full_reference(*prepared[k], scale_factor)
                continue
            for k, m, g1, u1 in zip(idx, mets, full, uni):
                m["psnr"] = float(self.finish(g1)[0])
                m["psnr_level"] = self._assess_psnr(m["psnr"])
                m["ssim"] = float(self.finish(u1)[1])
                m["ms_ssim"] = float(self.finish(g1)[1])
                m["ssim_level"] = self._assess_ssim(m["ms_ssim"])
                m["overall_score"] = self._calculate_overall_score(m)
                results[k] = m
        return results
