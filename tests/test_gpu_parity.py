"""GPU suite: every kernel of the hot path, called through the C ABI, against the oracle and the reference's golden
vectors.  Bars (BASELINE.json north_star): tiles bit-exact; canvas within +-1 LSB with >= 99.9 % exact; PSNR / SSIM
within 1e-4 relative."""
import ctypes as C

import numpy as np
import pytest

from helpers import FUSE_CASES, canvas_agreement, fuse_case, merge_case, synth

pytestmark = pytest.mark.gpu
REL = 1e-4


@pytest.fixture(scope="module")
def bm5(srs):
    return srs.BlendingModule(num_levels=5)


# ------------------------------------------------------------------ tiling
@pytest.mark.parametrize("mode", ["mirror", "replicate", "reflect", "constant"])
def test_tiles_bit_exact_vs_reference(srs, golden, mode):
    tm = srs.TilingModule(96, 0.2, mode, enable_content_aware=False)
    tiles = tm.split_array(golden["tiling_src"])
    want, meta = golden[f"tiling_tiles_{mode}"], golden[f"tiling_meta_{mode}"]
    assert len(tiles) == len(want)
    for t, w, m in zip(tiles, want, meta):
        md = t.metadata
        assert [md.global_x, md.global_y, md.input_w, md.input_h, md.output_w, md.output_h, md.overlap_top,
                md.overlap_bottom, md.overlap_left, md.overlap_right] == [int(v) for v in m]
        assert t.data.dtype == np.uint8 and np.array_equal(t.data, w)
    tm32 = srs.TilingModule(32, 0.2, mode, enable_content_aware=False)
    got = tm32._apply_padding(golden["tiling_tiny_src"], 0, 27, 0, 25)  # multi-bounce: pad >> size
    assert np.array_equal(got, golden[f"tiling_tiny_{mode}"])


def test_tiles_vs_oracle_large_and_gray(srs, oracle):
    rng = np.random.default_rng(5)
    for shape, block, ratio in [((1500, 2300, 3), 1024, 0.2), ((700, 333), 256, 0.3), ((64, 64, 3), 64, 0.1)]:
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        tm = srs.TilingModule(block, ratio, "mirror", enable_content_aware=False)
        for t in tm.split_array(img):
            m = t.metadata
            assert np.array_equal(t.data, oracle.tiling_oracle.extract_tile(img, m.global_x, m.global_y, m.input_w,
                                                                          m.input_h, block, "mirror"))


def test_tiles_device_path(srs):
    import torch
    rng = np.random.default_rng(6)
    img = rng.integers(0, 256, (500, 700, 3), dtype=np.uint8)
    tm = srs.TilingModule(256, 0.2, "reflect", enable_content_aware=False)
    host = tm.split_array(img)
    dev = tm.split_array(torch.from_numpy(img).cuda())
    for a, b in zip(host, dev):
        assert b.data.is_cuda and np.array_equal(a.data, b.data.cpu().numpy())


# ------------------------------------------------------------------ pyramids (API-compat entry points)
def test_pyramid_helpers_bit_exact_vs_reference(bm5, golden):
    gp = bm5.build_gaussian_pyramid(golden["pyr_src"], 5)
    assert len(gp) == 5
    for i, g in enumerate(gp):
        assert np.array_equal(g, golden[f"pyr_g{i}"]), i
    lp = bm5.build_laplacian_pyramid(gp)
    for i, l in enumerate(lp):
        assert np.array_equal(l, golden[f"pyr_l{i}"]), i
    assert np.array_equal(bm5.collapse_laplacian_pyramid(lp), golden["pyr_collapse"])


def test_pyr_down_up_vs_oracle_shapes(bm5, oracle):
    rng = np.random.default_rng(8)
    for h, w in [(1, 1), (1, 5), (2, 2), (3, 7), (16, 16), (37, 53), (128, 130), (257, 64)]:
        for cn in (1, 3):
            x = (rng.random((h, w, cn)) * 255 + rng.random((h, w, cn))).astype(np.float32)
            x = np.ascontiguousarray(x[:, :, 0] if cn == 1 else x)
            assert np.array_equal(bm5._pyr_down(x), oracle.pyr.pyr_down(x, "c")), (h, w, cn)
            for dh, dw in ((2 * h, 2 * w), (2 * h - 1, 2 * w - 1)):
                if dh >= 1 and dw >= 1:
                    assert np.array_equal(bm5._pyr_up(x, (dh, dw)), oracle.pyr.pyr_up(x, (dh, dw), "c")), (h, w, cn, dh, dw)
    with pytest.raises(ValueError):
        bm5._pyr_up(np.zeros((4, 4), np.float32), (9, 9))


# ------------------------------------------------------------------ fusion
@pytest.mark.parametrize("name", FUSE_CASES)
def test_fusion_vs_reference_golden(srs, golden, name):
    tiles, pos, out_hw, levels, kind, want = fuse_case(golden, name)
    infos = [srs.TileInfo(np.ascontiguousarray(t), x, y, 0, 0) for t, (y, x) in zip(tiles, pos)]
    got = srs.BlendingModule(num_levels=levels).laplacian_fusion(infos, output_shape=out_hw, weight_type=srs.WeightType(kind))
    assert got.shape == want.shape and got.dtype == np.uint8
    mx, exact = canvas_agreement(got, want)
    assert mx <= 1 and exact >= 0.999, (name, mx, exact)


def _grid_case(rng, th, tw, ny, nx, sy, sx, c=3):
    tiles, pos = [], []
    for j in range(ny):
        for i in range(nx):
            tiles.append(synth(rng, th, tw, c))
            pos.append((j * sy, i * sx))
    return tiles, pos


@pytest.mark.parametrize("levels", [1, 2, 3, 5, 7])
def test_fusion_vs_oracle_levels(srs, oracle, levels):
    rng = np.random.default_rng(100 + levels)
    tiles, pos = _grid_case(rng, 256, 320, 2, 3, 200, 250)
    out_hw = (440, 800)  # bottom/right tiles overhang the canvas
    infos = [srs.TileInfo(t, x, y, 0, 0) for t, (y, x) in zip(tiles, pos)]
    got = srs.BlendingModule(num_levels=levels).laplacian_fusion(infos, output_shape=out_hw)
    want = oracle.fusion_oracle.laplacian_fuse(tiles, pos, out_hw, levels, "cosine", "c")
    mx, exact = canvas_agreement(got, want)
    assert mx <= 1 and exact >= 0.999, (levels, mx, exact)


@pytest.mark.parametrize("c", [1, 3])
def test_fusion_every_gather_class(srs, oracle, c):
    """tiles wide enough for whole interior blocks (128 x 16 pixels) on an even grid: unit-weight cores, 1 / 2 / 4-tile interior
    cells, and the edge cells (tile edges inside a block, the canvas edge) all go through the lean gather kernels, gray and RGB"""
    rng = np.random.default_rng(40 + c)
    tiles, pos = _grid_case(rng, 640, 704, 2, 2, 500, 520, c=c)
    out_hw = (1100, 1190)  # bottom / right tiles overhang the canvas
    infos = [srs.TileInfo(t, x, y, 0, 0) for t, (y, x) in zip(tiles, pos)]
    for levels, kind in ((3, "cosine"), (5, "sigmoid")):  # sigmoid: the weight at a tile's border is not 0
        got = srs.BlendingModule(num_levels=levels).laplacian_fusion(infos, output_shape=out_hw, weight_type=srs.WeightType(kind))
        want = oracle.fusion_oracle.laplacian_fuse(tiles, pos, out_hw, levels, kind, "c")
        mx, exact = canvas_agreement(got, want)
        assert mx <= 1 and exact >= 0.999, (c, levels, kind, mx, exact)
    import os
    os.environ["SRF_GATHER_ONE_CLASS"] = "1"  # the same bytes through the generic kernel alone
    try:
        generic = srs.BlendingModule(num_levels=5).laplacian_fusion(infos, output_shape=out_hw, weight_type=srs.WeightType("sigmoid"))
    finally:
        del os.environ["SRF_GATHER_ONE_CLASS"]
    assert np.array_equal(generic, got)


def test_fusion_odd_positions_and_sizes(srs, oracle):
    """odd tile origins flip the pyrUp phase relative to the canvas blocks; odd sizes hit the 2n-1 pyrUp case"""
    rng = np.random.default_rng(21)
    shapes = [(201, 333), (201, 333), (157, 129), (96, 415)]
    pos = [(3, 5), (101, 207), (150, 1), (7, 300)]
    tiles = [synth(rng, h, w) for h, w in shapes]
    out_hw = (330, 560)
    infos = [srs.TileInfo(t, x, y, 0, 0) for t, (y, x) in zip(tiles, pos)]
    for kind in ("cosine", "linear", "sigmoid"):
        got = srs.BlendingModule(num_levels=6).laplacian_fusion(infos, output_shape=out_hw, weight_type=srs.WeightType(kind))
        want = oracle.fusion_oracle.laplacian_fuse(tiles, pos, out_hw, 6, kind, "c")
        mx, exact = canvas_agreement(got, want)
        assert mx <= 1 and exact >= 0.999, (kind, mx, exact)


def test_fusion_default_output_shape_and_grid_layout(srs, oracle):
    rng = np.random.default_rng(22)
    tiles = [synth(rng, 128, 128) for _ in range(4)]
    bm = srs.BlendingModule(num_levels=4)
    got = bm.laplacian_fusion(tiles, output_shape=(256, 256))  # bare arrays: sqrt-grid layout
    pos = [(0, 0), (0, 128), (128, 0), (128, 128)]
    want = oracle.fusion_oracle.laplacian_fuse(tiles, pos, (256, 256), 4, "cosine", "c")
    assert canvas_agreement(got, want)[0] <= 1
    infos = [srs.TileInfo(t, x, y, 0, 0) for t, (y, x) in zip(tiles, [(0, 0), (0, 100), (90, 0), (90, 100)])]
    got = bm.laplacian_fusion(infos)  # output_shape=None -> max extents
    assert got.shape == (218, 228, 3)


def test_fusion_device_path_equals_host_path(srs):
    import torch
    rng = np.random.default_rng(23)
    tiles, pos = _grid_case(rng, 192, 192, 2, 2, 150, 150)
    bm = srs.BlendingModule(num_levels=5)
    host = bm.laplacian_fusion([srs.TileInfo(t, x, y, 0, 0) for t, (y, x) in zip(tiles, pos)], output_shape=(342, 342))
    dev = bm.laplacian_fusion([srs.TileInfo(torch.from_numpy(t).cuda(), x, y, 0, 0) for t, (y, x) in zip(tiles, pos)],
                              output_shape=(342, 342))
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), host)


def test_fusion_row_bands_reassemble(srs):
    """band sharding: any split of the canvas rows gives the same bytes as the single pass"""
    rng = np.random.default_rng(24)
    tiles, pos = _grid_case(rng, 256, 256, 3, 2, 190, 200)
    infos = [srs.TileInfo(t, x, y, 0, 0) for t, (y, x) in zip(tiles, pos)]
    bm = srs.BlendingModule(num_levels=6)
    out_hw = (636, 456)
    full = bm.laplacian_fusion(infos, output_shape=out_hw)
    for cuts in ([0, 318, 636], [0, 101, 333, 500, 636], [0, 1, 635, 636]):
        acc = np.zeros_like(full)
        for a, b in zip(cuts[:-1], cuts[1:]):
            part = bm.laplacian_fusion(infos, output_shape=out_hw, row_range=(a, b))
            assert not part[:a].any() and not part[b:].any()
            acc[a:b] = part[a:b]
        assert np.array_equal(acc, full), cuts


def test_fusion_full_size_tile_properties(srs, oracle):
    """one 4096^2 tile at BASELINE size: compared with the oracle on a band, plus idempotence of the call"""
    rng = np.random.default_rng(25)
    small = synth(rng, 512, 512)
    tile = np.ascontiguousarray(np.kron(small, np.ones((8, 8, 1), np.uint8)))
    tile = np.clip(tile.astype(np.int16) + rng.integers(-5, 6, tile.shape), 0, 255).astype(np.uint8)
    bm = srs.BlendingModule(num_levels=7)
    info = [srs.TileInfo(tile, 0, 0, 0, 0)]
    a = bm.laplacian_fusion(info, output_shape=(4096, 4096))
    b = bm.laplacian_fusion(info, output_shape=(4096, 4096))
    assert np.array_equal(a, b)
    want = oracle.fusion_oracle.laplacian_fuse([tile], [(0, 0)], (4096, 4096), 7, "cosine", "cv2" if oracle.pyr.have_cv2() else "c")
    mx, exact = canvas_agreement(a, want)
    assert mx <= 1 and exact >= 0.999, (mx, exact)


# ------------------------------------------------------------------ IQA
def test_iqa_vs_reference_golden(srs, golden, oracle):
    qa = srs.QualityAssessmentModule()
    a, b = golden["iqa_a"], golden["iqa_b"]
    assert qa.calculate_psnr(a, b) == pytest.approx(float(golden["iqa_psnr"]), rel=1e-12)
    assert qa.calculate_psnr(a, b[:200, :300]) == pytest.approx(float(golden["iqa_psnr_crop"]), rel=1e-12)
    assert qa.calculate_psnr(a, a) == float("inf")
    assert qa.calculate_ssim(a, b) == pytest.approx(oracle.iqa_oracle.ssim(a, b), rel=REL)
    assert qa.calculate_ssim(a, b) == pytest.approx(float(golden["iqa_ssim_fallback"]), rel=3e-4)
    assert qa.calculate_ssim(a, a) == pytest.approx(1.0, rel=1e-6)


@pytest.mark.parametrize("shape", [(11, 11, 3), (12, 300, 3), (300, 13, 3), (129, 257, 3), (640, 480), (1000, 1500, 3)])
def test_iqa_vs_oracle_shapes(srs, oracle, shape):
    rng = np.random.default_rng(sum(shape))
    a = synth(rng, shape[0], shape[1], shape[2] if len(shape) == 3 else 1)
    b = np.clip(a.astype(np.int16) + rng.integers(-12, 13, a.shape), 0, 255).astype(np.uint8)
    qa = srs.QualityAssessmentModule()
    p, s = qa.calculate_psnr_ssim(a, b)
    assert p == pytest.approx(oracle.iqa_oracle.psnr(a, b), rel=1e-12)
    assert s == pytest.approx(oracle.iqa_oracle.ssim(a, b), rel=REL)
    sums = qa.psnr_ssim_sums(a, b)
    assert sums.sq_err == oracle.iqa_oracle.sq_err_sum(a, b)
    assert sums.ssim_count == (shape[0] - 10) * (shape[1] - 10)


def test_iqa_hard_cases(srs, oracle):
    qa = srs.QualityAssessmentModule()
    rng = np.random.default_rng(31)
    flat = np.full((200, 260, 3), 201, np.uint8)
    near = flat.copy()
    near[::7, ::5] -= 1  # near-flat pair: worst case for float32 cancellation in the variances
    assert qa.calculate_ssim(flat, near) == pytest.approx(oracle.iqa_oracle.ssim(flat, near), rel=REL)
    sat = rng.choice(np.array([0, 255], np.uint8), (150, 170, 3))
    assert qa.calculate_ssim(sat, 255 - sat) == pytest.approx(oracle.iqa_oracle.ssim(sat, 255 - sat), rel=REL, abs=1e-6)
    with pytest.raises(ValueError):
        qa.calculate_ssim(np.zeros((8, 40, 3), np.uint8), np.zeros((8, 40, 3), np.uint8))
    # the reference's max() <= 1 rule: an all-0/1 uint8 image is multiplied by 255 first
    bits = rng.integers(0, 2, (64, 64, 3), dtype=np.uint8)
    other = rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)
    assert qa.calculate_psnr(bits, other) == pytest.approx(oracle.iqa_oracle.psnr(bits, other), rel=1e-12)
    f01 = rng.random((64, 64, 3))
    assert qa.calculate_psnr(f01, other) == pytest.approx(oracle.iqa_oracle.psnr(f01, other), rel=1e-12)


def test_iqa_bands_add_up(srs):
    rng = np.random.default_rng(32)
    a = synth(rng, 700, 900)
    b = np.clip(a.astype(np.int16) + rng.integers(-9, 10, a.shape), 0, 255).astype(np.uint8)
    qa = srs.QualityAssessmentModule()
    full = qa.psnr_ssim_sums(a, b)
    sq = cnt = 0
    ss = 0.0
    for r0, r1 in [(0, 3), (3, 64), (64, 350), (350, 699), (699, 700)]:
        s = qa.psnr_ssim_sums(a, b, row_range=(r0, r1))
        sq, cnt, ss = sq + s.sq_err, cnt + s.ssim_count, ss + s.ssim_sum
    assert sq == full.sq_err and cnt == full.ssim_count
    assert ss == pytest.approx(full.ssim_sum, rel=1e-7)  # float32 per-thread partial sums regroup with the band split


def test_iqa_device_path_and_evaluate(srs, oracle):
    import torch
    rng = np.random.default_rng(33)
    a = synth(rng, 400, 520)
    b = np.clip(a.astype(np.int16) + rng.integers(-9, 10, a.shape), 0, 255).astype(np.uint8)
    qa = srs.QualityAssessmentModule()
    host = qa.calculate_psnr_ssim(a, b)
    dev = qa.calculate_psnr_ssim(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())
    assert host == dev
    rep = qa.evaluate_full_reference(a, b)
    assert set(rep) >= {"psnr", "psnr_level", "ssim", "ms_ssim", "ssim_level", "overall_score"}
    assert rep["ssim"] == pytest.approx(oracle.iqa_oracle.ssim_uniform(a, b), rel=REL)
    assert rep["overall_score"] == pytest.approx(np.mean([min(100, rep["psnr"]), rep["ms_ssim"] * 100]))
    assert len(qa.batch_evaluate([(a, b), (b, a)])) == 2


def test_fallback_ssim_equals_reference_outputs(srs, golden, oracle):
    """_calculate_ssim_simple on the GPU against the numbers the reference itself produced (its fallback path)"""
    import torch
    qa = srs.QualityAssessmentModule()
    a, b = golden["iqa_a"], golden["iqa_b"]
    want = float(golden["iqa_ssim_fallback"])
    assert qa._calculate_ssim_simple(a, b) == pytest.approx(want, rel=1e-5)
    assert qa._calculate_ssim_simple(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()) == pytest.approx(want, rel=1e-5)
    np.random.seed(42)
    hr = np.random.randint(0, 255, (512, 512, 3), dtype=np.uint8)
    sr = np.clip(hr.astype(np.int16) + np.random.normal(0, 5, hr.shape).astype(np.int16), 0, 255).astype(np.uint8)
    assert qa._calculate_ssim_simple(hr, sr) == pytest.approx(float(golden["iqa_demo_ssim_fallback"]), rel=1e-5)


# ------------------------------------------------------------------ detect_seams (row N4)
def test_window_ssim_and_seams_vs_reference(srs, golden, oracle):
    import torch
    bm = srs.BlendingModule(num_levels=5, ssim_threshold=0.9)
    res = golden["seam_result"]
    y1, x1 = (int(v) for v in golden["fuse_grid_pos"][1])
    t1 = golden["fuse_grid_t1"]
    host = bm.window_ssim_map(res, t1, x1, y1)
    assert np.array_equal(host, golden["seam_scores_t1"])              # the reference's _compute_ssim loop, bit for bit
    dev = bm.window_ssim_map(torch.from_numpy(res).cuda(), torch.from_numpy(t1).cuda(), x1, y1)
    assert np.array_equal(dev, host)
    assert bm._compute_ssim(t1[:16, :16], res[y1:y1 + 16, x1:x1 + 16]) == golden["seam_scores_t1"][0, 0]
    infos = [srs.TileInfo(golden[f"fuse_grid_t{i}"], int(x), int(y), 0, 0) for i, (y, x) in enumerate(golden["fuse_grid_pos"])]
    seams = bm.detect_seams(res, infos)
    want = golden["seam_list"]
    assert len(seams) == len(want)
    for s, w in zip(seams, want):
        assert (s.x, s.y, s.width, s.height) == tuple(int(v) for v in w[:4])
        assert s.ssim_score == pytest.approx(w[4], rel=1e-13, abs=1e-15)
    # other window sizes (n not a power of two: NumPy's two-pass variance differs in the last bits), gray, overhang
    rng = np.random.default_rng(91)
    a = synth(rng, 150, 170)
    b = np.clip(a.astype(np.int16) + rng.integers(-30, 31, a.shape), 0, 255).astype(np.uint8)
    for win, st, (x, y) in [(10, 5, (0, 0)), (16, 8, (37, 21)), (7, 3, (100, 90))]:
        got = bm.window_ssim_map(b, a[:120, :110], x, y, win, st)
        want_m = oracle.fusion_oracle.window_ssim_map(b, a[:120, :110], x, y, win, st)
        assert got.shape == want_m.shape and np.allclose(got, want_m, rtol=1e-12, atol=1e-14)
    g = bm.window_ssim_map(b[:, :, 0].copy(), a[:, :, 0].copy(), 0, 0)
    assert np.array_equal(g, oracle.fusion_oracle.window_ssim_map(b[:, :, 0], a[:, :, 0], 0, 0))


# ------------------------------------------------------------------ merge_tiles (row N3)
def _merge_tiles_objects(srs, tiles, to_device=False):
    import torch
    out = []
    for data, gx, gy, ow, oh, ov in tiles:
        meta = srs.tiling_module.TileMetadata(global_x=gx, global_y=gy, output_w=ow, output_h=oh, overlap_top=ov[0],
                                              overlap_bottom=ov[1], overlap_left=ov[2], overlap_right=ov[3])
        out.append(srs.tiling_module.Tile(metadata=meta, data=torch.from_numpy(data).cuda() if to_device else data))
    return out


def test_merge_tiles_bit_exact_vs_reference(srs, golden, oracle):
    tiles, (w, h) = merge_case(golden)
    tm = srs.TilingModule(64, 0.25, "mirror", enable_content_aware=False)
    objs = _merge_tiles_objects(srs, tiles)
    assert np.array_equal(tm.merge_tiles(objs, w, h, blending=True), golden["merge_out_blend"])
    assert np.array_equal(tm.merge_tiles(objs, w, h, blending=False), golden["merge_out_plain"])
    dev = tm.merge_tiles(_merge_tiles_objects(srs, tiles, to_device=True), w, h)
    assert np.array_equal(dev.cpu().numpy(), golden["merge_out_blend"])
    # a larger, oracle-checked case through the mirror's own split (block 256, 3 x 3 tiles, squashed right / bottom tiles)
    rng = np.random.default_rng(81)
    src = synth(rng, 600, 720)
    tm2 = srs.TilingModule(256, 0.2, "mirror", enable_content_aware=False)
    qa = srs.QualityAssessmentModule()
    got_tiles, ref_tiles = [], []
    for t in tm2.split_array(src):
        t.data = qa.upsample_bicubic(t.data, (512, 512))
        got_tiles.append(t)
        m = t.metadata
        ref_tiles.append((t.data, m.global_x, m.global_y, m.output_w, m.output_h,
                          (m.overlap_top, m.overlap_bottom, m.overlap_left, m.overlap_right)))
    want = oracle.tiling_oracle.merge_tiles(ref_tiles, 1440, 1200, 2.0, True)
    assert np.array_equal(tm2.merge_tiles(got_tiles, 1440, 1200), want)
    thin = [t for t in tm2.split_array(synth(rng, 300, 700))]  # last column 85 px wide: its scaled "overlap" (240) exceeds
    with pytest.raises(ValueError):                             # the 170 px tile -- the reference's ramp fails to broadcast
        tm2.merge_tiles(thin, 1400, 600)
    # tiles the upscaler returned SMALLER than their canvas size are enlarged (cv2.resize INTER_LINEAR, tiling_module.py:1107-1109):
    # the reference's own output for 1.5x data under output_scale 2 (golden), then mixed directions against the oracle
    up_tiles = [(d, *t[1:]) for d, t in zip(golden["merge_up_tiles"], tiles)]
    assert np.array_equal(tm.merge_tiles(_merge_tiles_objects(srs, up_tiles), w, h, blending=True), golden["merge_up_out_blend"])
    dev_up = tm.merge_tiles(_merge_tiles_objects(srs, up_tiles, to_device=True), w, h, blending=False)
    assert np.array_equal(dev_up.cpu().numpy(), golden["merge_up_out_plain"])
    mixed = []
    for k, t in enumerate(ref_tiles):
        hh, ww = [(300, 512), (512, 333), (700, 640), (512, 512), (97, 1201)][k % 5]
        mixed.append((synth(rng, hh, ww), *t[1:]))
    want = oracle.tiling_oracle.merge_tiles(mixed, 1440, 1200, 2.0, True)
    assert np.array_equal(tm2.merge_tiles(_merge_tiles_objects(srs, mixed), 1440, 1200), want)


# ------------------------------------------------------------------ bicubic resize + down-sample comparison (row N1)
def test_resize_cubic_bit_equals_oracle_and_reference(srs, golden, oracle):
    import torch
    qa = srs.QualityAssessmentModule()
    ro = oracle.resize_oracle
    src = golden["resize_src"]
    for s in (0.1, 0.2, 0.4):
        got = qa.downsample_bicubic(src, s)
        assert np.array_equal(got, ro.downsample_bicubic(src, s))          # the restatement, bit for bit
        d = np.abs(got.astype(np.int16) - golden[f"resize_down_{s}"])      # the reference's own output (cv2 -> IPP)
        assert d.max() <= 1 and (d != 0).mean() <= 1e-3
    assert np.array_equal(qa.upsample_bicubic(src[:60, :70], (120, 140)), golden["resize_up_2x"])
    assert np.array_equal(qa.upsample_bicubic(src[:60, :70], (101, 93)), ro.resize_cubic_u8(src[:60, :70], 93, 101))
    g = np.ascontiguousarray(src[:, :, 1])
    assert np.array_equal(qa.downsample_bicubic(g, 0.4), ro.downsample_bicubic(g, 0.4))
    rng = np.random.default_rng(77)
    big = synth(rng, 1500, 2100)
    dev = qa.downsample_bicubic(torch.from_numpy(big).cuda(), 0.4)          # device path, strided views included
    assert np.array_equal(dev.cpu().numpy(), ro.downsample_bicubic(big, 0.4))
    view = big[3:1203, 5:1805]
    assert np.array_equal(qa.downsample_bicubic(view, 0.2), ro.downsample_bicubic(np.ascontiguousarray(view), 0.2))
    assert np.array_equal(qa.upsample_bicubic(big[:300, :400], (600, 800)), ro.resize_cubic_u8(big[:300, :400], 800, 600))
    with pytest.raises(ValueError):
        qa.downsample_bicubic(np.zeros((5, 5, 3), np.uint8), 0.1)            # int(5 * 0.1) == 0: cv2 rejects an empty dsize


def test_evaluate_full_reference_has_every_key(srs, golden, oracle):
    qa = srs.QualityAssessmentModule()
    a, b = golden["dsc_a"], golden["dsc_b"]
    rep = qa.evaluate_full_reference(a, b)
    want = oracle.resize_oracle.downsample_comparison(a, b)
    assert list(rep)[:6] == ["psnr_structure_color", "ssim_structure_color", "psnr_mid_frequency", "ssim_mid_frequency",
                             "psnr_high_frequency", "ssim_high_frequency"]          # the reference's key order
    for k, v in want.items():
        assert rep[k] == pytest.approx(v, rel=REL), k
    for name in ("structure_color", "mid_frequency", "high_frequency"):              # the reference's own numbers
        assert rep[f"psnr_{name}"] == pytest.approx(float(golden[f"dsc_psnr_{name}"]), rel=REL)
    assert rep["psnr"] == pytest.approx(oracle.iqa_oracle.psnr(a, b), rel=1e-12)
    assert rep["ms_ssim"] == pytest.approx(oracle.iqa_oracle.ssim(a, b), rel=REL)
    assert rep["ssim"] == pytest.approx(oracle.iqa_oracle.ssim_uniform(a, b), rel=REL)


# ------------------------------------------------------------------ end to end
def test_pipeline_run_small(srs, oracle):
    rng = np.random.default_rng(40)
    src = synth(rng, 90, 160)
    cfg = srs.PipelineConfig(block_size=128, overlap_ratio=0.25, target_resolution="640x360", num_pyramid_levels=4)
    pipe = srs.SuperResolutionPipeline(cfg)
    res = pipe.run(src)
    assert res["canvas"].shape == (360, 640, 3) and res["tiles"] >= 2
    assert 0 < res["report"]["ms_ssim"] <= 1
    cfg_w = srs.PipelineConfig(block_size=128, overlap_ratio=0.25, target_resolution="640x360", blend_method="weighted")
    res_w = srs.SuperResolutionPipeline(cfg_w).run(src)
    assert res_w["canvas"].shape == (360, 640, 3) and not np.array_equal(res_w["canvas"], res["canvas"])
    assert res_w["report"]["psnr"] > res["report"]["psnr"] - 20


def test_pipeline_device_resident_equals_host_upscaler(srs):
    """run() with the built-in stub keeps every stage on the device; a user upscaler (cv2 on the host) takes the host
    path -- the 2x bicubic up-scale is bit-identical, so both give the same canvas"""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(43)
    src = synth(rng, 120, 200)
    cfg = srs.PipelineConfig(block_size=128, overlap_ratio=0.2, target_resolution="800x480", num_pyramid_levels=5)
    dev = srs.SuperResolutionPipeline(cfg).run(src)
    host = srs.SuperResolutionPipeline(cfg, upscaler=lambda t, hw: cv2.resize(t, (hw[1], hw[0]), interpolation=cv2.INTER_CUBIC)).run(src)
    assert dev["tiles"] == host["tiles"] and np.array_equal(dev["canvas"], host["canvas"])
    assert {"psnr_structure_color", "ssim_high_frequency", "psnr", "ssim", "ms_ssim", "overall_score"} <= set(dev["report"])
    assert dev["report"]["psnr"] == pytest.approx(host["report"]["psnr"], rel=1e-12)
    assert dev["report"]["ms_ssim"] == pytest.approx(host["report"]["ms_ssim"], rel=1e-6)


def test_fuse_and_assess_band_pipeline_equals_two_calls(srs):
    """the pipelined host entry point (several bands, three streams) returns the bytes and sums of the plain calls"""
    rng = np.random.default_rng(41)
    tiles, pos = [], []
    for j in range(4):
        for i in range(2):
            tiles.append(synth(rng, 1024, 512))
            pos.append((j * 820, i * 400))
    out_hw = (3400, 900)  # 3 bands of ~1134 rows; the last tile row overhangs
    infos = [srs.TileInfo(t, x, y, 0, 0) for t, (y, x) in zip(tiles, pos)]
    ref = synth(rng, out_hw[0], out_hw[1])
    bm = srs.BlendingModule(num_levels=6)
    qa = srs.QualityAssessmentModule()
    canvas, sums = bm.fuse_and_assess(infos, ref, out_hw)
    want = bm.laplacian_fusion(infos, output_shape=out_hw)
    assert np.array_equal(canvas, want)
    s2 = qa.psnr_ssim_sums(ref, want)
    assert (sums.sq_err, sums.ssim_count, sums.n_elems) == (s2.sq_err, s2.ssim_count, s2.n_elems)
    assert sums.ssim_sum == pytest.approx(s2.ssim_sum, rel=1e-7)
    pipe = srs.SuperResolutionPipeline(srs.PipelineConfig(num_pyramid_levels=6))
    c2, rep = pipe.fuse_and_assess(infos, ref, out_hw)
    assert np.array_equal(c2, want) and rep["psnr"] == pytest.approx(qa.finish(s2)[0], rel=1e-12)


def test_fusion_plan_reuse_and_misaligned_device_tiles(srs):
    import torch
    rng = np.random.default_rng(42)
    tiles, pos = _grid_case(rng, 130, 131, 2, 2, 100, 100)  # 131*3 bytes per row: misaligned pitch on the device
    bm = srs.BlendingModule(num_levels=4)
    host = bm.laplacian_fusion([srs.TileInfo(t, x, y, 0, 0) for t, (y, x) in zip(tiles, pos)], output_shape=(230, 231))
    dev_infos = [srs.TileInfo(torch.from_numpy(t).cuda(), x, y, 0, 0) for t, (y, x) in zip(tiles, pos)]
    plan = bm.plan_fusion(dev_infos, output_shape=(230, 231))
    out = torch.empty((230, 231, 3), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        got = bm.laplacian_fusion(plan, out=out)
        assert got is out and np.array_equal(got.cpu().numpy(), host)
    hplan = bm.plan_fusion([srs.TileInfo(t, x, y, 0, 0) for t, (y, x) in zip(tiles, pos)], output_shape=(230, 231))
    assert np.array_equal(bm.laplacian_fusion(hplan), host)


def test_repeated_fuse_call_is_served_by_the_captured_graph(srs):
    """consecutive identical device calls on a real stream: the second is captured, the later ones are one graph launch each
    (srf_graph_replays counts them -- a signature that never repeats would silently re-plan every call), and a replay
    reads the tiles' CURRENT pixels"""
    import torch
    rng = np.random.default_rng(44)
    tiles, pos = _grid_case(rng, 320, 384, 2, 2, 250, 300)
    out_hw = (570, 684)
    bm = srs.BlendingModule(num_levels=5)
    dev_tiles = [torch.from_numpy(t).cuda() for t in tiles]
    plan = bm.plan_fusion([srs.TileInfo(d, x, y, 0, 0) for d, (y, x) in zip(dev_tiles, pos)], output_shape=out_hw)
    out = torch.empty(out_hw + (3,), dtype=torch.uint8, device="cuda")
    lib, ctx = srs._native.load(), srs._native.context(0)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        r0 = lib.srf_graph_replays(ctx)
        for _ in range(5):
            bm.laplacian_fusion(plan, out=out)
        side.synchronize()
        assert lib.srf_graph_replays(ctx) - r0 >= 3
        first = out.cpu().numpy().copy()
        for d, t in zip(dev_tiles, tiles):  # new pixels behind the same pointers
            d.copy_(torch.from_numpy(np.ascontiguousarray(t[::-1, ::-1])), non_blocking=False)
        r1 = lib.srf_graph_replays(ctx)
        bm.laplacian_fusion(plan, out=out)
        side.synchronize()
        assert lib.srf_graph_replays(ctx) == r1 + 1
        second = out.cpu().numpy().copy()
        r2 = lib.srf_graph_replays(ctx)
        bm.laplacian_fusion(plan, row_range=(100, 400), out=out)  # another band: a new plan, not a replay
        side.synchronize()
        assert lib.srf_graph_replays(ctx) == r2
    infos = [srs.TileInfo(t, x, y, 0, 0) for t, (y, x) in zip(tiles, pos)]
    assert np.array_equal(first, bm.laplacian_fusion(infos, output_shape=out_hw))
    flipped = [srs.TileInfo(np.ascontiguousarray(t[::-1, ::-1]), x, y, 0, 0) for t, (y, x) in zip(tiles, pos)]
    assert np.array_equal(second, bm.laplacian_fusion(flipped, output_shape=out_hw))


def test_pipelined_fuse_assess_equals_the_serial_loop(srs):
    """distributed.PipelinedFuseAssess (bench.py's default loop): the assess of step k on its own stream while a second
    library context fuses step k + 1 into the other canvas.  Both canvases hold the serial call's bytes, the sums of the
    last step are the serial sums, both contexts replay their captured graphs, and new pixels behind the same tile
    pointers show up in the next steps (nothing is cached across steps)"""
    import torch
    rng = np.random.default_rng(45)
    tiles, pos = _grid_case(rng, 320, 384, 2, 3, 250, 300)
    out_hw = (570, 984)
    ref = synth(rng, *out_hw)
    bm = srs.BlendingModule(num_levels=5)
    qa = srs.QualityAssessmentModule()
    want = bm.laplacian_fusion([srs.TileInfo(t, x, y, 0, 0) for t, (y, x) in zip(tiles, pos)], output_shape=out_hw)
    s_want = qa.psnr_ssim_sums(ref, want)
    dev = torch.device("cuda", 0)
    dev_tiles = [torch.from_numpy(t).to(dev) for t in tiles]
    infos = [srs.TileInfo(d, x, y, 0, 0) for d, (y, x) in zip(dev_tiles, pos)]
    D = srs.distributed
    lib = srs._native.load()
    pipe = D.PipelinedFuseAssess(srs, infos, out_hw, 5, 0, torch.from_numpy(ref).to(dev), (0, out_hw[0]), (0, out_hw[0]), D.SumsExchange(dev))
    c0, c1 = srs._native.context(0), srs._native.context(0, 1)
    assert c0.value != c1.value
    g0 = lib.srf_graph_replays(c0) + lib.srf_graph_replays(c1)
    for _ in range(9):
        pipe.submit()
    got = pipe.result()
    assert lib.srf_graph_replays(c0) + lib.srf_graph_replays(c1) - g0 >= 5
    pipe.join()
    torch.cuda.synchronize()
    for cv in pipe.canvas:
        assert np.array_equal(cv.cpu().numpy(), want)
    assert (got.sq_err, got.ssim_count, got.n_elems) == (s_want.sq_err, s_want.ssim_count, s_want.n_elems)
    assert got.ssim_sum == s_want.ssim_sum  # same kernel, same bytes, same fixed-order reduction
    for d, t in zip(dev_tiles, tiles):
        d.copy_(torch.from_numpy(np.ascontiguousarray(t[::-1, ::-1])))
    torch.cuda.synchronize()
    flipped = bm.laplacian_fusion([srs.TileInfo(np.ascontiguousarray(t[::-1, ::-1]), x, y, 0, 0) for t, (y, x) in zip(tiles, pos)], output_shape=out_hw)
    for _ in range(2):
        pipe.submit()
    got2 = pipe.result()
    pipe.join()
    torch.cuda.synchronize()
    for cv in pipe.canvas:
        assert np.array_equal(cv.cpu().numpy(), flipped)
    assert got2.sq_err == qa.psnr_ssim_sums(ref, flipped).sq_err
    # a band: only the band's rows are fused and assessed, the sums are the band's share
    pipe_b = D.PipelinedFuseAssess(srs, infos, out_hw, 5, 0, torch.from_numpy(ref).to(dev), (200, 400), (194, 405), D.SumsExchange(dev))
    for _ in range(4):
        pipe_b.submit()
    band = pipe_b.result()
    pipe_b.join()
    torch.cuda.synchronize()
    assert np.array_equal(pipe_b.last_canvas().cpu().numpy()[194:405], flipped[194:405])
    assert band.sq_err == int(((flipped[200:400].astype(np.int64) - ref[200:400].astype(np.int64)) ** 2).sum())


def test_fusion_graph_replay_and_device_sums(srs):
    """a repeated device call on a side stream is captured into a CUDA graph and replayed; it must track the tile
    CONTENTS (same addresses, new pixels) and agree with the host path; the enqueue-only IQA agrees with the sync one"""
    import torch
    rng = np.random.default_rng(43)
    tiles, pos = _grid_case(rng, 256, 256, 2, 2, 200, 200)
    bm = srs.BlendingModule(num_levels=5)
    qa = srs.QualityAssessmentModule()
    out_hw = (456, 456)
    dev_tiles = [torch.from_numpy(t).cuda() for t in tiles]
    plan = bm.plan_fusion([srs.TileInfo(d, x, y, 0, 0) for d, (y, x) in zip(dev_tiles, pos)], output_shape=out_hw)
    out = torch.empty(out_hw + (3,), dtype=torch.uint8, device="cuda")
    ctx = srs._native.context(0)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for rep in range(5):
            if rep == 3:  # new pixels behind the same pointers
                tiles = [np.ascontiguousarray(t[::-1, ::-1]) for t in tiles]
                for d, t in zip(dev_tiles, tiles):
                    d.copy_(torch.from_numpy(t), non_blocking=False)
            l0 = srs._native.load().srf_launch_count(ctx)
            bm.laplacian_fusion(plan, out=out)
            side.synchronize()
            assert srs._native.load().srf_launch_count(ctx) > l0
            want = bm.laplacian_fusion([srs.TileInfo(t, x, y, 0, 0) for t, (y, x) in zip(tiles, pos)], output_shape=out_hw)
            assert np.array_equal(out.cpu().numpy(), want), rep
        ref = torch.from_numpy(synth(rng, *out_hw)).cuda()
        d_sums = torch.zeros(5, dtype=torch.float64, device="cuda")
        qa.psnr_ssim_sums_device(ref, out, d_sums, row_range=(10, 400))
        got = srs.distributed.allreduce_device_sums(d_sums, 390 * 456 * 3)
        s = qa.psnr_ssim_sums(ref, out, row_range=(10, 400))
    assert (got.sq_err, got.ssim_count) == (s.sq_err, s.ssim_count) and got.ssim_sum == s.ssim_sum


@pytest.mark.parametrize("shape", [(7, 7, 3), (9, 300, 3), (257, 129, 3), (480, 640)])
def test_uniform_window_ssim_vs_oracle(srs, oracle, shape):
    """calculate_ssim(multiscale=False): scikit-image's default 7x7 uniform window, sample covariance"""
    rng = np.random.default_rng(1000 + sum(shape))
    a = synth(rng, shape[0], shape[1], shape[2] if len(shape) == 3 else 1)
    b = np.clip(a.astype(np.int16) + rng.integers(-12, 13, a.shape), 0, 255).astype(np.uint8)
    qa = srs.QualityAssessmentModule()
    assert qa.calculate_ssim(a, b, multiscale=False) == pytest.approx(oracle.iqa_oracle.ssim_uniform(a, b), rel=REL)
    s = qa.psnr_ssim_sums(a, b, window=1)
    assert s.ssim_count == (shape[0] - 6) * (shape[1] - 6) and s.sq_err == oracle.iqa_oracle.sq_err_sum(a, b)
    with pytest.raises(ValueError):
        qa.calculate_ssim(np.zeros((6, 40, 3), np.uint8), np.zeros((6, 40, 3), np.uint8), multiscale=False)


def test_fuse_and_assess_row_bands_add_up(srs):
    """band mode of the pipelined host call: bands reassemble the canvas and their sums add up to the full call's"""
    rng = np.random.default_rng(44)
    tiles, pos = [], []
    for j in range(3):
        for i in range(2):
            tiles.append(synth(rng, 1024, 600))
            pos.append((j * 800, i * 450))
    out_hw = (2600, 1000)
    infos = [srs.TileInfo(t, x, y, 0, 0) for t, (y, x) in zip(tiles, pos)]
    ref = synth(rng, out_hw[0], out_hw[1])
    bm = srs.BlendingModule(num_levels=5)
    full, fs = bm.fuse_and_assess(infos, ref, out_hw)
    acc = np.zeros_like(full)
    sq = cnt = 0
    ss = 0.0
    for r0, r1 in [(0, 700), (700, 701), (701, 2000), (2000, 2600)]:
        part, s = bm.fuse_and_assess(infos, ref, out_hw, row_range=(r0, r1))
        assert not part[:r0].any() and not part[r1:].any()
        acc[r0:r1] = part[r0:r1]
        sq, cnt, ss = sq + s.sq_err, cnt + s.ssim_count, ss + s.ssim_sum
    assert np.array_equal(acc, full)
    assert (sq, cnt) == (fs.sq_err, fs.ssim_count) and ss == pytest.approx(fs.ssim_sum, rel=1e-7)


def test_weighted_average_fusion_vs_reference_golden(srs, golden):
    tiles, pos, out_hw, _, _, _ = fuse_case(golden, "ragged")
    bm = srs.BlendingModule()
    infos = [srs.TileInfo(np.ascontiguousarray(t), x, y, 0, 0) for t, (y, x) in zip(tiles, pos)]
    got = bm.weighted_average_fusion(infos, weight_type=srs.WeightType.LINEAR, output_shape=out_hw)
    assert np.array_equal(got, golden["wavg_ragged_out"])
    t4 = [np.ascontiguousarray(golden[f"fuse_linear6_t{i}"]) for i in range(4)]
    assert np.array_equal(bm.weighted_average_fusion(t4), golden["wavg_grid_out"])  # bare arrays: sqrt-grid, max extents
    # unit maps given explicitly == a plain average of the tiles where they overlap (here they do not): same canvas as
    # the reference computes for weights of ones
    ones = bm.weighted_average_fusion(t4, weights=[np.ones((64, 64), np.float32)] * 4)
    assert np.array_equal(ones, np.vstack([np.hstack([t4[0], t4[1]]), np.hstack([t4[2], t4[3]])]))


def test_weighted_average_custom_maps_vs_reference_golden(srs, golden, oracle):
    """weighted_average_fusion(weights=...) through srf_weighted_average[_host]: the reference's own outputs for float32,
    float64 and mixed map lists, RGB and gray, host arrays and CUDA tensors; then a larger random case against the oracle"""
    import torch
    tiles = [np.ascontiguousarray(golden[f"wcus_t{i}"]) for i in range(4)]
    pos = [tuple(int(v) for v in p) for p in golden["wcus_pos"]]
    w32 = [golden[f"wcus_w32_{i}"] for i in range(3)]
    w64 = [golden[f"wcus_w64_{i}"] for i in range(4)]
    bm = srs.BlendingModule(num_levels=5)
    infos = [srs.TileInfo(t, x, y, 0, 0) for t, (y, x) in zip(tiles, pos)]
    assert np.array_equal(bm.weighted_average_fusion(infos, weights=w32, output_shape=(100, 150)), golden["wcus_out32"])
    assert np.array_equal(bm.weighted_average_fusion(infos, weights=w64), golden["wcus_out64"])
    assert np.array_equal(bm.weighted_average_fusion(infos, weights=[w32[0], w64[1]], weight_type=srs.WeightType.LINEAR), golden["wcus_mixed"])
    ginfos = [srs.TileInfo(np.ascontiguousarray(t[:, :, 1]), x, y, 0, 0) for t, (y, x) in zip(tiles, pos)]
    assert np.array_equal(bm.weighted_average_fusion(ginfos, weights=[w64[0], w32[1]], output_shape=(110, 140)), golden["wcus_gray64"])
    dinfos = [srs.TileInfo(torch.from_numpy(t).cuda(), x, y, 0, 0) for t, (y, x) in zip(tiles, pos)]
    dev = bm.weighted_average_fusion(dinfos, weights=[torch.from_numpy(w64[0]).cuda(), w64[1], w64[2], w64[3]])
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), golden["wcus_out64"])
    rng = np.random.default_rng(77)
    big = [synth(rng, 300, 420) for _ in range(5)]
    bpos = [(0, 0), (0, 380), (250, 100), (260, 500), (500, 0)]
    maps = [rng.random((300, 420)) if k % 2 else rng.random((300, 420), dtype=np.float32) for k in range(5)]
    maps[3] = -maps[3]  # negative weights: the denominator clamps at 1e-6, the quotient clips
    want = oracle.fusion_oracle.weighted_average_custom(big, bpos, maps, (700, 900))
    got = bm.weighted_average_fusion([srs.TileInfo(t, x, y, 0, 0) for t, (y, x) in zip(big, bpos)], weights=maps, output_shape=(700, 900))
    assert np.array_equal(got, want)
    # a tile that lies entirely off the canvas contributes nothing (host and device paths)
    off = big + [synth(rng, 64, 64)]
    opos = bpos + [(900, 50)]
    omaps = maps + [np.ones((64, 64), np.float32)]
    assert np.array_equal(bm.weighted_average_fusion([srs.TileInfo(t, x, y, 0, 0) for t, (y, x) in zip(off, opos)], weights=omaps,
                                                     output_shape=(700, 900)), want)
    dev_off = bm.weighted_average_fusion([srs.TileInfo(torch.from_numpy(t).cuda(), x, y, 0, 0) for t, (y, x) in zip(off, opos)], weights=omaps,
                                         output_shape=(700, 900))
    assert np.array_equal(dev_off.cpu().numpy(), want)
    with pytest.raises(ValueError):  # a map smaller than its tile does not broadcast in the reference either
        bm.weighted_average_fusion(infos, weights=[np.ones((10, 10), np.float32)])
    with pytest.raises(ValueError):
        bm.weighted_average_fusion(infos, weights=w32, row_range=(0, 10))


def test_config_sized_two_tile_fuse_and_iqa_vs_oracle(srs, oracle):
    """BASELINE config geometry at full tile size: two adjacent bicubic-stub tiles of the 200 MP grid (4096^2 each,
    818-pixel overlap, 7 levels) fused into their 4096 x 7374 canvas and scored on a 1024-row strip, against the oracle"""
    import workloads
    wl = workloads.build("cfg4_200mp", max_tiles=2)
    tiles, pos = wl.tiles, wl.positions_yx
    assert pos == [(0, 0), (0, 3278)] and tiles[0].shape == (4096, 4096, 3)
    out_hw = (4096, 3278 + 4096)
    bm = srs.BlendingModule(num_levels=wl.levels)
    got = bm.laplacian_fusion([srs.TileInfo(t, x, y, 0, 0) for t, (y, x) in zip(tiles, pos)], output_shape=out_hw)
    want = oracle.fusion_oracle.laplacian_fuse(tiles, pos, out_hw, wl.levels, "cosine", "cv2" if oracle.pyr.have_cv2() else "c")
    mx, exact = canvas_agreement(got, want)
    assert mx <= 1 and exact >= 0.999, (mx, exact)
    ref = wl.reference[:1024, :out_hw[1]]
    qa = srs.QualityAssessmentModule()
    p, s = qa.calculate_psnr_ssim(ref, got[:1024])
    assert p == pytest.approx(oracle.iqa_oracle.psnr(ref, want[:1024]), rel=1e-4 if exact < 1 else 1e-12)
    assert s == pytest.approx(oracle.iqa_oracle.ssim(ref, want[:1024]), rel=REL)
