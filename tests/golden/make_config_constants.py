"""Pin the FULL BASELINE configs to outputs of the reference itself (run in the build container, where
/root/reference exists; `oracle/_ref` holds verbatim copies of its modules, see oracle/make_ref.py).

    python tests/golden/make_config_constants.py [cfg4_200mp cfg2_100mp ...]

For every named workload (workloads.py; synthetic, seeded) this runs the REFERENCE's
BlendingModule.laplacian_fusion on all tiles and writes to tests/golden/config_constants.json:
    canvas_sha256      of the fused uint8 canvas (bit-exactness check at full size: `bench.py --verify`, tests -m gpu)
    block_sums         int64 sums of the canvas over 512 x 512 blocks and channels (diagnostic when the hash differs)
    sq_err, n_elems    sum (reference_canvas - fused)^2 over all elements (exact integers) -> psnr
    psnr               10 log10(255^2 n / sq_err)
    ssim               oracle/iqa_oracle.ssim (scikit-image's algorithm: 11x11 Gaussian, interior mean) in float64
    ref_psnr, ref_ssim what the reference's QualityAssessmentModule.calculate_psnr / calculate_ssim return here
                       (scikit-image absent -> its own fallbacks; _calculate_ssim_simple differs from `ssim` only in
                       the 5-pixel border)
    seconds            wall time of the reference calls on this container's cores (context for the CPU baseline)
The oracle restatement (oracle/fusion_oracle.py) is asserted to produce the SAME canvas bytes.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
OUT = os.path.join(HERE, "config_constants.json")
BLOCK = 512


def block_sums(img: np.ndarray) -> list:
    h, w = img.shape[:2]
    ny, nx = -(-h // BLOCK), -(-w // BLOCK)
    out = np.zeros((ny, nx, img.shape[2]), np.int64)
    for j in range(ny):
        rows = img[j * BLOCK:(j + 1) * BLOCK].astype(np.int64)
        for i in range(nx):
            out[j, i] = rows[:, i * BLOCK:(i + 1) * BLOCK].sum(axis=(0, 1))
    return out.tolist()


def main():
    import workloads
    from oracle import fusion_oracle, iqa_oracle, make_ref, ref_arm
    make_ref.make()
    names = sys.argv[1:] or ["cfg4_200mp", "cfg2_100mp"]
    doc = {}
    if os.path.exists(OUT):
        with open(OUT) as f:
            doc = json.load(f)
    for name in names:
        wl = workloads.build(name)
        H, W = wl.canvas_hw
        t0 = time.time()
        canvas = ref_arm.laplacian_fuse(wl.tiles, wl.positions_yx, (H, W), wl.levels)
        t_fuse = time.time() - t0
        t0 = time.time()
        rp = ref_arm.psnr(wl.reference, canvas)
        t_psnr = time.time() - t0
        t0 = time.time()
        rs = ref_arm.ssim(wl.reference, canvas)
        t_ssim = time.time() - t0
        want = fusion_oracle.laplacian_fuse(wl.tiles, wl.positions_yx, (H, W), wl.levels, "cosine", "c")
        assert np.array_equal(want, canvas), f"{name}: the oracle restatement and the reference disagree on the canvas"
        del want
        sq = 0
        for r in range(0, H, 1024):  # exact integer sum in row strips
            d = wl.reference[r:r + 1024].astype(np.int64) - canvas[r:r + 1024].astype(np.int64)
            sq += int((d * d).sum())
        n = H * W * 3
        doc[name] = {
            "canvas_hw": [H, W], "tiles": len(wl.tiles), "levels": wl.levels,
            "canvas_sha256": hashlib.sha256(np.ascontiguousarray(canvas).tobytes()).hexdigest(),
            "block": BLOCK, "block_sums": block_sums(canvas),
            "sq_err": sq, "n_elems": n, "psnr": float(10 * np.log10(255.0 ** 2 * n / sq)) if sq else float("inf"),
            "ssim": float(iqa_oracle.ssim(wl.reference, canvas)),
            "ref_psnr": float(rp), "ref_ssim": float(rs),
            "seconds": {"laplacian_fusion": t_fuse, "calculate_psnr": t_psnr, "calculate_ssim": t_ssim, "cores": os.cpu_count()},
            "generated_by": "tests/golden/make_config_constants.py (reference modules via oracle/_ref)",
        }
        print(name, {k: v for k, v in doc[name].items() if k != "block_sums"}, flush=True)
        with open(OUT, "w") as f:
            json.dump(doc, f)
    return 0


if __name__ == "__main__":
    sys.exit(main())
