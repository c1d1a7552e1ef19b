"""Generate the golden vectors in this directory by running the REFERENCE modules themselves
(/root/reference, importable only in the build container -- the vectors travel, the reference does not).

    python tests/golden/make_golden.py

tiling_module and blending_module import as shipped (cv2 + numpy).  quality_assessment_module imports lpips and
scikit-image at module level; neither is installed, so both are stubbed with functions that RAISE, which makes the
reference take its own manual fallbacks: PSNR formula (quality_assessment_module.py:313-318) and
_calculate_ssim_simple (391-417).  Inputs are stored next to the outputs so nothing depends on RNG stability.
"""
import logging
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def synth(rng, h, w, c=3):
    """textured, non-saturated, non-flat uint8 image (blurred noise, contrast-stretched, + N(0,4))"""
    import cv2
    f = cv2.GaussianBlur(rng.random((h, w, c), dtype=np.float32) * 255, (0, 0), 3)
    f = (f - f.mean()) * 4 + 128 + rng.normal(0, 4, f.shape)
    return np.clip(f, 0, 255).astype(np.uint8).reshape((h, w, c) if c > 1 else (h, w))


def main():
    os.chdir(tempfile.mkdtemp())
    logging.disable(logging.CRITICAL)
    lp = types.ModuleType("lpips")

    class _NoLpips:
        def __init__(self, *a, **k):
            raise RuntimeError("lpips absent")

    lp.LPIPS = _NoLpips
    sys.modules["lpips"] = lp
    sk, skm = types.ModuleType("skimage"), types.ModuleType("skimage.metrics")

    def _absent(*a, **k):
        raise RuntimeError("skimage absent")

    skm.peak_signal_noise_ratio = skm.structural_similarity = _absent
    sk.metrics = skm
    sys.modules["skimage"], sys.modules["skimage.metrics"] = sk, skm
    sys.path.insert(0, REF)
    import blending_module as rb
    import quality_assessment_module as rq
    import tiling_module as rt
    import cv2

    rng = np.random.default_rng(20261018)
    out = {}

    # ---- tiling: grid, overlaps, padded tiles for every padding mode -------------------------------------
    geo = []
    for (W, H, block, ratio) in [(8660, 5773, 2048, 0.2), (7500, 5000, 2048, 0.3), (6122, 4081, 2048, 0.1),
                                 (6122, 3443, 2048, 0.2), (1280, 720, 2048, 0.2), (300, 200, 128, 0.25),
                                 (129, 513, 128, 0.1), (100, 50, 64, 0.3)]:
        tm = rt.TilingModule(block, ratio, "mirror", enable_content_aware=False, l2_cache_dir=tempfile.mkdtemp())
        pos = tm._calculate_tile_positions(W, H)
        ov = [tm._calculate_overlap_for_tile(x, y, w, h, W, H) for (x, y, w, h) in pos]
        geo.append((W, H, block, ratio, len(pos)))
        out[f"tiling_pos_{W}x{H}_{block}_{ratio}"] = np.array(pos, np.int32)
        out[f"tiling_ov_{W}x{H}_{block}_{ratio}"] = np.array(ov, np.int32)
    out["tiling_cases"] = np.array(geo, np.float64)
    src = synth(rng, 150, 217)
    out["tiling_src"] = src
    png = os.path.join(tempfile.mkdtemp(), "src.png")
    cv2.imwrite(png, src[:, :, ::-1])
    for mode in ("mirror", "replicate", "reflect", "constant"):
        tm = rt.TilingModule(96, 0.2, mode, enable_content_aware=False, l2_cache_dir=tempfile.mkdtemp())
        tiles = tm.split_image(png)
        out[f"tiling_tiles_{mode}"] = np.stack([t.data for t in tiles])
        out[f"tiling_meta_{mode}"] = np.array([[t.metadata.global_x, t.metadata.global_y, t.metadata.input_w,
                                                t.metadata.input_h, t.metadata.output_w, t.metadata.output_h,
                                                t.metadata.overlap_top, t.metadata.overlap_bottom,
                                                t.metadata.overlap_left, t.metadata.overlap_right] for t in tiles], np.int32)
    # a crop much smaller than the block: multi-bounce reflection (pad > size - 1)
    tiny = synth(rng, 5, 7)
    out["tiling_tiny_src"] = tiny
    for mode in ("mirror", "replicate", "reflect", "constant"):
        tm = rt.TilingModule(32, 0.2, mode, enable_content_aware=False, l2_cache_dir=tempfile.mkdtemp())
        out[f"tiling_tiny_{mode}"] = tm._apply_padding(tiny, 0, 32 - 5, 0, 32 - 7)

    # ---- pyramids: the reference's own helpers ---------------------------------------------------------------
    bm = rb.BlendingModule(num_levels=5)
    img = synth(rng, 77, 101)
    out["pyr_src"] = img
    gp = bm.build_gaussian_pyramid(img, 5)
    lp_ = bm.build_laplacian_pyramid(gp)
    for i, (g, l) in enumerate(zip(gp, lp_)):
        out[f"pyr_g{i}"] = g
        out[f"pyr_l{i}"] = l
    out["pyr_collapse"] = bm.collapse_laplacian_pyramid(lp_)
    for kind in ("linear", "cosine", "sigmoid"):
        out[f"weight_{kind}_64x80"] = bm._create_distance_weight_map(64, 80, rb.WeightType(kind))

    # ---- fusion: 2x3 ragged grid with overhang, three weight types, odd sizes ------------------------------
    def fuse_case(name, shapes_pos, out_hw, levels, kind="cosine", c=3):
        tiles = [synth(rng, h, w, c) for (h, w, _, _) in shapes_pos]
        infos = [rb.TileInfo(t, x, y, 0, 0) for t, (_, _, y, x) in zip(tiles, shapes_pos)]
        res = rb.BlendingModule(num_levels=levels).laplacian_fusion(infos, output_shape=out_hw, weight_type=rb.WeightType(kind))
        out[f"fuse_{name}_n"] = np.array([len(tiles), levels, out_hw[0], out_hw[1]], np.int32)
        out[f"fuse_{name}_pos"] = np.array([(y, x) for (_, _, y, x) in shapes_pos], np.int32)
        for i, t in enumerate(tiles):
            out[f"fuse_{name}_t{i}"] = t
        out[f"fuse_{name}_kind"] = np.array(kind)
        out[f"fuse_{name}_out"] = res

    grid = [(128, 128, y, x) for y in (0, 96) for x in (0, 96, 192)]
    fuse_case("grid", grid, (210, 300), 5)                       # bottom/right tiles overhang and are cropped
    fuse_case("ragged", [(100, 141, 0, 0), (77, 90, 11, 120), (130, 64, 60, 40)], (160, 215), 4, "sigmoid")
    fuse_case("linear6", [(64, 64, 0, 0), (64, 64, 0, 40), (64, 64, 40, 0), (64, 64, 40, 40)], (104, 104), 6, "linear")
    fuse_case("gray", [(96, 96, 0, 0), (96, 96, 0, 64)], (96, 160), 3, "cosine", c=1)
    fuse_case("single1", [(72, 88, 3, 5)], (80, 100), 1)         # levels=1: no pyramid, hole around the tile
    fuse_case("deep", [(48, 40, 0, 0), (48, 40, 0, 24)], (48, 64), 9)  # asks for more levels than the tile supports

    # ---- IQA: the reference's manual fallbacks ------------------------------------------------------------
    qa = rq.QualityAssessmentModule()
    a = synth(rng, 211, 307)
    b = np.clip(a.astype(np.float64) + rng.normal(0, 5, a.shape), 0, 255).astype(np.uint8)
    out["iqa_a"], out["iqa_b"] = a, b
    out["iqa_psnr"] = np.array(qa.calculate_psnr(a, b))
    out["iqa_ssim_fallback"] = np.array(qa.calculate_ssim(a, b, multiscale=True))
    out["iqa_gray_a"] = cv2.cvtColor(a, cv2.COLOR_RGB2GRAY)
    out["iqa_psnr_crop"] = np.array(qa.calculate_psnr(a, b[:200, :300]))
    out["iqa_psnr_same"] = np.array(qa.calculate_psnr(a, a))
    # the seeded demo of the reference's own __main__ (quality_assessment_module.py:1394-1400)
    np.random.seed(42)
    hr = np.random.randint(0, 255, (512, 512, 3), dtype=np.uint8)
    noise = np.random.normal(0, 5, hr.shape).astype(np.int16)
    sr = np.clip(hr.astype(np.int16) + noise, 0, 255).astype(np.uint8)
    out["iqa_demo_psnr"] = np.array(qa.calculate_psnr(hr, sr))
    out["iqa_demo_ssim_fallback"] = np.array(qa.calculate_ssim(hr, sr, multiscale=True))

    # ---- weighted-average fusion (the "next" row N3): same tiles as the ragged Laplacian case ----------------
    wt = [out[f"fuse_ragged_t{i}"] for i in range(3)]
    wp = out["fuse_ragged_pos"]
    winfos = [rb.TileInfo(t, int(x), int(y), 0, 0) for t, (y, x) in zip(wt, wp)]
    out["wavg_ragged_out"] = rb.BlendingModule().weighted_average_fusion(winfos, weight_type=rb.WeightType.LINEAR, output_shape=(160, 215))
    out["wavg_grid_out"] = rb.BlendingModule().weighted_average_fusion([out[f"fuse_linear6_t{i}"] for i in range(4)])

    # ---- bicubic down-/up-sampling and the down-sample comparison (row N1; appended last so that the draws above
    # keep their values).  downsample_bicubic / upsample_bicubic are cv2.resize(INTER_CUBIC), which this wheel routes to
    # Intel IPP; the PSNR keys come from the reference's manual formula.  Its SSIM keys would come from the 11x11
    # fallback (scikit-image is stubbed out), which is NOT the 7x7 window of multiscale=False, so they are not stored.
    ra = synth(rng, 157, 203)
    out["resize_src"] = ra
    for s_ in (0.1, 0.2, 0.4):
        out[f"resize_down_{s_}"] = qa.downsample_bicubic(ra, s_)
    out["resize_up_2x"] = qa.upsample_bicubic(ra[:60, :70], (120, 140))
    out["resize_up_odd"] = qa.upsample_bicubic(ra[:60, :70], (101, 93))
    out["resize_gray_down_0.4"] = qa.downsample_bicubic(np.ascontiguousarray(ra[:, :, 1]), 0.4)
    da = synth(rng, 300, 420)
    db = np.clip(da.astype(np.float64) + rng.normal(0, 5, da.shape), 0, 255).astype(np.uint8)
    out["dsc_a"], out["dsc_b"] = da, db
    dsc = qa._evaluate_downsample_comparison(da, db, 4)
    for k_, v_ in dsc.items():
        if k_.startswith("psnr_"):
            out[f"dsc_{k_}"] = np.array(v_)

    # ---- merge_tiles (row N3): the reference's own split_image -> (bicubic 2x stand-in) -> merge_tiles.  The right and
    # bottom tiles are padded, so their processed data is larger than output_w x output_h and the reference squashes
    # it with cv2.resize's default INTER_LINEAR before the feathered average.
    ms = synth(rng, 100, 150)
    cv2.imwrite("merge_src.png", ms[:, :, ::-1])
    tmm = rt.TilingModule(block_size=64, overlap_ratio=0.25, padding_mode="mirror", output_scale=2.0,
                          l2_cache_dir=tempfile.mkdtemp(), enable_content_aware=False)
    mt = tmm.split_image("merge_src.png", save_metadata=False)
    for t_ in mt:
        t_.data = cv2.resize(t_.data, (128, 128), interpolation=cv2.INTER_CUBIC)
    out["merge_n"] = np.array([len(mt), 64, 150, 100], np.int32)
    out["merge_meta"] = np.array([[t_.metadata.global_x, t_.metadata.global_y, t_.metadata.input_w, t_.metadata.input_h,
                                   t_.metadata.output_w, t_.metadata.output_h, t_.metadata.overlap_top,
                                   t_.metadata.overlap_bottom, t_.metadata.overlap_left, t_.metadata.overlap_right] for t_ in mt], np.int32)
    out["merge_tiles"] = np.stack([t_.data for t_ in mt])
    out["merge_out_blend"] = tmm.merge_tiles(mt, 300, 200, blending=True)
    out["merge_out_plain"] = tmm.merge_tiles(mt, 300, 200, blending=False)
    out["merge_weight_4"] = tmm._create_blend_weight(mt[4])
    out["merge_squash_2"] = cv2.resize(mt[2].data, (mt[2].metadata.output_w, mt[2].metadata.output_h))

    # ---- detect_seams (row N4): the reference's window scores and merged seam list on the fused "grid" case; the result
    # is disturbed in one corner so that some windows fall below the threshold
    sg_tiles = [out[f"fuse_grid_t{i}"] for i in range(6)]
    sg_pos = out["fuse_grid_pos"]
    sg_res = out["fuse_grid_out"].copy()
    sg_res[40:90, 100:170] = np.clip(sg_res[40:90, 100:170].astype(np.int16) + rng.integers(-60, 61, (50, 70, 3)), 0, 255).astype(np.uint8)
    sbm = rb.BlendingModule(num_levels=5, ssim_threshold=0.9)
    infos = [rb.TileInfo(t, int(x), int(y), 0, 0) for t, (y, x) in zip(sg_tiles, sg_pos)]
    seams = sbm.detect_seams(sg_res, infos, window_size=16, stride=8)
    out["seam_result"] = sg_res
    out["seam_list"] = np.array([[s_.x, s_.y, s_.width, s_.height, s_.ssim_score] for s_ in seams], np.float64).reshape(-1, 5)
    t1, (y1_, x1_) = sg_tiles[1], sg_pos[1]
    roi = sg_res[int(y1_):int(y1_) + 128, int(x1_):int(x1_) + 128]
    out["seam_scores_t1"] = np.array([[sbm._compute_ssim(t1[wy:wy + 16, wx:wx + 16], roi[wy:wy + 16, wx:wx + 16])
                                       for wx in range(0, 113, 8)] for wy in range(0, 113, 8)])

    # ---- appended in round 2 (after every earlier draw, so the vectors above do not move) ----
    # merge_tiles on tiles that come back SMALLER than output_w x output_h (a 1.5x stand-in under output_scale 2): the
    # reference enlarges them with cv2.resize's default INTER_LINEAR (tiling_module.py:1107-1109)
    for t_ in mt:
        t_.data = cv2.resize(out["merge_tiles"][mt.index(t_)], (96, 96), interpolation=cv2.INTER_AREA)
    out["merge_up_tiles"] = np.stack([t_.data for t_ in mt])
    out["merge_up_out_blend"] = tmm.merge_tiles(mt, 300, 200, blending=True)
    out["merge_up_out_plain"] = tmm.merge_tiles(mt, 300, 200, blending=False)

    # weighted_average_fusion with CUSTOM weight maps (blending_module.py:661-760, `weights=`): float32 maps, float64
    # maps, a list shorter than the tile list (the last tile gets the generated cosine map), maps larger than the tile,
    # a tile that overhangs the canvas, RGB and gray
    wq_tiles = [synth(rng, 64, 80), synth(rng, 64, 80), synth(rng, 48, 96), synth(rng, 64, 64)]
    wq_pos = [(0, 0), (0, 60), (40, 10), (50, 90)]  # (y, x)
    wq_w32 = [rng.random((64, 80), dtype=np.float32), (rng.random((70, 90), dtype=np.float32) * 3).astype(np.float32),
              rng.random((48, 96, 1), dtype=np.float32)]
    wq_w64 = [rng.random((64, 80)), rng.random((64, 80)) * 1e-3, rng.random((48, 96)) + 0.5, rng.random((64, 64))]
    wbm = rb.BlendingModule(num_levels=5)
    infos = [rb.TileInfo(t, int(x), int(y), 0, 0) for t, (y, x) in zip(wq_tiles, wq_pos)]
    for i, t in enumerate(wq_tiles):
        out[f"wcus_t{i}"] = t
    out["wcus_pos"] = np.array(wq_pos, np.int32)
    for i, m in enumerate(wq_w32):
        out[f"wcus_w32_{i}"] = m
    for i, m in enumerate(wq_w64):
        out[f"wcus_w64_{i}"] = m
    out["wcus_out32"] = wbm.weighted_average_fusion(infos, weights=wq_w32, output_shape=(100, 150))
    out["wcus_out64"] = wbm.weighted_average_fusion(infos, weights=wq_w64)
    out["wcus_mixed"] = wbm.weighted_average_fusion(infos, weights=[wq_w32[0], wq_w64[1]], weight_type=rb.WeightType.LINEAR)
    ginfos = [rb.TileInfo(t[:, :, 1].copy(), int(x), int(y), 0, 0) for t, (y, x) in zip(wq_tiles, wq_pos)]
    out["wcus_gray64"] = wbm.weighted_average_fusion(ginfos, weights=[wq_w64[0], wq_w32[1]], output_shape=(110, 140))

    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_vectors.npz"), os.path.getsize(os.path.join(HERE, "reference_vectors.npz")) // 1024, "KiB")
    print("cv2", cv2.__version__, "numpy", np.__version__)


if __name__ == "__main__":
    main()
