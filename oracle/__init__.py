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
    installed in this image).  Restated from its published algorithm; pinned to <=2e-5 relative against the
    reference's own fallback `_calculate_ssim_simple` (quality_assessment_module.py:391-417) and exactly
    against scipy.ndimage.gaussian_filter (what scikit-image calls) -- "parity pinned to the fallback, not to
    scikit-image itself".
  * bicubic resize (cv2.resize INTER_CUBIC, uint8): the opencv-python wheel calls Intel IPP (closed source); restated
    as the Keys cubic at float64 sample positions with half-to-even rounding.  PINNED against outputs of the
    reference's own downsample_bicubic / upsample_bicubic (tests/golden) and against cv2 directly: equal on all
    but ~4e-6 of the pixels, never more than 1 LSB apart, bit-exact for the 2x up-scale.
"""
from . import pyr, tiling_oracle, fusion_oracle, iqa_oracle, reference_port, resize_oracle  # noqa: F401
