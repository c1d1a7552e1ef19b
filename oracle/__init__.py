"""oracle/ -- TEST INFRASTRUCTURE, not product code.

CPU restatement of the reference's tile -> Laplacian-fuse -> PSNR/SSIM path
(/root/reference/tiling_module.py, blending_module.py, quality_assessment_module.py).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it;
the product package never does (tests/test_boundary.py greps for that).

Parity status (see DESIGN.md "Oracle"):
  * tiling + fusion: PINNED -- checked against outputs of the reference modules themselves, imported in the
    build container; the vectors are committed under tests/golden/ with the script that made them
    (tests/golden/make_golden.py).
  * PSNR: PINNED against the reference's own manual formula (quality_assessment_module.py:313-318).
  * Gaussian SSIM: the arithmetic lives in scikit-image (requirements.txt:13, ">=0.21.0", not vendored, not
    installed in this image), so in this image the reference itself returns its fallback `_calculate_ssim_simple`
    (quality_assessment_module.py:391-417: same 11 taps, BORDER_REFLECT_101, mean over the full map).
    iqa_oracle.ssim_simple restates that path (interior SSIM of the reflect-padded pair) and is PINNED to the
    reference's own outputs to 1e-16 (tests/golden: iqa_ssim_fallback, iqa_demo_ssim_fallback); the GPU kernel matches
    them to 1e-7 through QualityAssessmentModule._calculate_ssim_simple.  iqa_oracle.ssim (the scikit-image variant:
    interior mean, no padding -- the oracle of record for calculate_ssim) is the same filter arithmetic without the
    padding, restated from scikit-image's published algorithm and equal to scipy.ndimage.gaussian_filter (what
    scikit-image calls) to 1e-12; its values differ from the fallback's by <= 3e-4 (border handling only).
  * bicubic resize (cv2.resize INTER_CUBIC, uint8): the opencv-python wheel calls Intel IPP (closed source); restated
    as the Keys cubic at float64 sample positions with half-to-even rounding.  PINNED against outputs of the
    reference's own downsample_bicubic / upsample_bicubic (tests/golden) and against cv2 directly: equal on all
    but ~4e-6 of the pixels, never more than 1 LSB apart, bit-exact for the 2x up-scale.
"""
from . import pyr, tiling_oracle, fusion_oracle, iqa_oracle, reference_port, resize_oracle  # noqa: F401
