"""CPU suite, part 2: the drop-in boundary -- the C-ABI library loads and exports everything include/srfuse.h
declares, the host-exact entry points agree with the reference's golden vectors, the host-side mirrors keep the
reference's argument/error behaviour, and the product never routes through the oracle or a CPU fallback."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "super-resolution-system_b200")


def declared_functions():
    text = open(os.path.join(ROOT, "include", "srfuse.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(srf_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(srs):
    lib = ctypes.CDLL(srs._native.lib_path())
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/srfuse.h but not exported"
    assert set(names) == set(srs._native.SIGNATURES), "ctypes binding and header disagree"
    assert srs._native.load().srf_version() == 100


def test_no_oracle_or_cpu_fallback_in_product():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f"{f} imports the oracle"
                assert "libsrf_oracle" not in text, f"{f} loads the oracle library"
    for f in ("blending_module.py", "quality_assessment_module.py"):
        text = open(os.path.join(PKG, f)).read()
        assert not re.search(r"^\s*(from|import)\s+(cv2|scipy|skimage)\b", text, re.M), f"{f} must not compute on the CPU"


def test_context_creation_fails_loudly_without_gpu(srs):
    if os.path.exists("/dev/nvidiactl"):
        pytest.skip("a GPU is present")
    with pytest.raises(srs._native.SrfError, match="no CPU fallback"):
        srs._native.context(0)
    bm = srs.BlendingModule(num_levels=3)
    with pytest.raises(srs._native.SrfError):
        bm.laplacian_fusion([srs.TileInfo(np.zeros((64, 64, 3), np.uint8), 0, 0, 0, 0)], output_shape=(64, 64))


def test_tile_geometry_is_exact(srs, golden):
    for W, H, block, ratio, n in golden["tiling_cases"]:
        W, H, block, n = int(W), int(H), int(block), int(n)
        tm = srs.TilingModule(block, float(ratio), "mirror", enable_content_aware=False)
        pos = tm._calculate_tile_positions(W, H)
        assert len(pos) == n
        assert np.array_equal(np.array(pos, np.int32), golden[f"tiling_pos_{W}x{H}_{block}_{ratio}"])
        ov = [tm._calculate_overlap_for_tile(x, y, w, h, W, H) for (x, y, w, h) in pos]
        assert np.array_equal(np.array(ov, np.int32), golden[f"tiling_ov_{W}x{H}_{block}_{ratio}"])


def test_tile_geometry_property(srs, oracle):
    rng = np.random.default_rng(11)
    for _ in range(300):
        block = int(rng.integers(16, 3000))
        ratio = float(rng.uniform(0.1, 0.3))
        W, H = int(rng.integers(1, 20000)), int(rng.integers(1, 20000))
        tm = srs.TilingModule(block, ratio, enable_content_aware=False)
        assert tm.overlap_pixels == oracle.tiling_oracle.overlap_pixels(block, ratio)
        assert tm._calculate_tile_positions(W, H) == oracle.tiling_oracle.tile_positions(W, H, block, tm.overlap_pixels)


def test_constructor_errors_match_reference(srs):
    with pytest.raises(ValueError):
        srs.TilingModule(overlap_ratio=0.05)
    with pytest.raises(ValueError):
        srs.TilingModule(overlap_ratio=0.31)
    with pytest.raises(ValueError):
        srs.TilingModule(padding_mode="wrap")
    with pytest.raises(ValueError):
        srs.BlendingModule(method="nope")
    with pytest.raises(ValueError):
        srs.TilingModule(enable_content_aware=False).split_image("/nonexistent/image.png")
    tm = srs.TilingModule(2048, 0.2)
    assert (tm.overlap_pixels, tm.output_size) == (409, 4096)
    assert srs.TilingModule(2048, 0.1).overlap_pixels == 204 and srs.TilingModule(2048, 0.3).overlap_pixels == 614


def test_fusion_argument_errors(srs):
    bm = srs.BlendingModule(num_levels=4)
    with pytest.raises(ValueError):
        bm.laplacian_fusion([])
    with pytest.raises(ValueError):  # bare arrays without output_shape: the reference raises ValueError too
        bm.laplacian_fusion([np.zeros((64, 64, 3), np.uint8)])
    with pytest.raises(TypeError):  # float tiles that are not uint8 values (exact ones are cast, as blending_module.py:440-441 would)
        bm.laplacian_fusion([srs.TileInfo(np.full((64, 64, 3), 0.5, np.float32), 0, 0, 0, 0)], output_shape=(64, 64))
    with pytest.raises(TypeError):
        bm.laplacian_fusion([srs.TileInfo(np.full((64, 64, 3), 300, np.int32), 0, 0, 0, 0)], output_shape=(64, 64))
    with pytest.raises(ValueError):
        bm.laplacian_fusion([srs.TileInfo(np.zeros((4, 64, 3), np.uint8), 0, 0, 0, 0)], output_shape=(64, 64))


def test_weight_profile_is_the_reference_map(srs, golden):
    bm = srs.BlendingModule()
    for kind in ("linear", "cosine", "sigmoid"):
        got = bm._create_distance_weight_map(64, 80, srs.WeightType(kind))
        assert got.dtype == np.float32 and np.array_equal(got, golden[f"weight_{kind}_64x80"])
    # C helper for non-Python callers agrees with the NumPy profile to 1 ulp
    buf = (ctypes.c_float * 513)()
    assert srs._native.load().srf_weight_profile(1, 512, buf) == 0
    prof = srs.blending_module.weight_profile(512, srs.WeightType.COSINE)
    assert np.max(np.abs(np.frombuffer(buf, np.float32) - prof)) <= 6e-8


def test_pipeline_api_surface(srs):
    cfg = srs.PipelineConfig()
    assert (cfg.block_size, cfg.overlap_ratio, cfg.padding_mode, cfg.num_pyramid_levels, cfg.blend_method) == \
        (2048, 0.2, "mirror", 6, "laplacian")
    pipe = srs.SuperResolutionPipeline(cfg)
    assert pipe._calculate_target_size((1280, 720), "100MP") == (12245, 6887)
    assert pipe._calculate_target_size((3000, 2000), "200MP") == (17320, 11546)
    assert pipe._calculate_target_size((100, 100), "640x480") == (640, 480)
    tc = srs.TilingConfig()
    assert (tc.block_size, tc.output_block_size, tc.overlap_ratio, tc.min_overlap_ratio, tc.max_overlap_ratio) == \
        (2048, 4096, 0.2, 0.1, 0.3)
    md = srs.TileMetadata(global_x=3)
    assert srs.TileMetadata.from_dict(md.to_dict()).global_x == 3
    res = srs.PipelineResult(True, "x", 0.1, 1, 1, 0, None, None, None)
    assert res.success


def test_iqa_finish_matches_oracle_formula(srs, oracle):
    s = srs._native.srf_iqa_sums(sq_err=123456, n_elems=3000, ssim_sum=87.5, ssim_count=100, bits_a=255, bits_b=255)
    p, q = srs.QualityAssessmentModule.finish(s)
    assert p == pytest.approx(10 * np.log10(255.0 ** 2 / (123456 / 3000)), rel=1e-15) and q == 0.875
    s.sq_err = 0
    assert srs.QualityAssessmentModule.finish(s)[0] == float("inf")


def test_resize_tap_tables_equal_oracle(srs, oracle):
    """host side of srf_resize_cubic_u8: the per-axis tap tables (float64 sample positions, Keys weights) bit for bit"""
    lib = srs._native.load()
    for src, dst in [(1531, 153), (1013, 405), (17320, 6928), (11546, 1154), (2048, 4096), (720, 1224), (5, 20), (7, 1), (1, 3)]:
        idx, w = np.zeros((dst, 4), np.int32), np.zeros((dst, 4), np.float32)
        assert lib.srf_resize_cubic_taps(src, dst, idx.ctypes.data, w.ctypes.data) == 0
        i2, w2 = oracle.resize_oracle.axis_table(src, dst)
        assert np.array_equal(idx, i2) and np.array_equal(w.view(np.uint32), w2.view(np.uint32))
    assert lib.srf_resize_cubic_taps(0, 4, idx.ctypes.data, w.ctypes.data) < 0


def test_downsample_argument_errors(srs):
    qa = srs.QualityAssessmentModule()
    img = np.zeros((20, 20, 3), np.uint8)
    for bad in (1.0, 0.0, -0.5, 1.5):
        with pytest.raises(ValueError):  # quality_assessment_module.py:240-241
            qa.downsample_bicubic(img, bad)


def test_linear_tap_tables_and_blend_weight(srs, golden, oracle):
    """host side of merge_tiles: OpenCV's bilinear tap tables (exact ints) and the feather map of _create_blend_weight"""
    lib = srs._native.load()
    for src, dst in [(128, 108), (128, 104), (4096, 1712), (4096, 930), (64, 64), (7, 3)]:
        idx, co = np.zeros((dst, 2), np.int32), np.zeros((dst, 2), np.int32)
        assert lib.srf_resize_linear_taps(src, dst, idx.ctypes.data, co.ctypes.data) == 0
        i0, i1, a0, a1 = oracle.resize_oracle.linear_table(src, dst)
        assert np.array_equal(idx[:, 0], i0) and np.array_equal(idx[:, 1], i1)
        assert np.array_equal(co[:, 0], a0) and np.array_equal(co[:, 1], a1)
    for src, dst, vertical in [(96, 128, 1), (96, 128, 0), (64, 128, 1), (7, 30, 1), (128, 104, 1), (1, 9, 1), (300, 301, 1)]:
        idx, co = np.zeros((dst, 2), np.int32), np.zeros((dst, 2), np.int32)
        assert lib.srf_resize_linear_taps_axis(src, dst, vertical, idx.ctypes.data, co.ctypes.data) == 0
        i0, i1, a0, a1 = oracle.resize_oracle.linear_table(src, dst, vertical=bool(vertical))
        assert np.array_equal(idx[:, 0], i0) and np.array_equal(idx[:, 1], i1)
        assert np.array_equal(co[:, 0], a0) and np.array_equal(co[:, 1], a1)
    tm = srs.TilingModule(64, 0.25, "mirror", enable_content_aware=False)
    m = golden["merge_meta"][4]
    meta = srs.tiling_module.TileMetadata(global_x=int(m[0]), global_y=int(m[1]), input_w=int(m[2]), input_h=int(m[3]),
                                          output_w=int(m[4]), output_h=int(m[5]), overlap_top=int(m[6]),
                                          overlap_bottom=int(m[7]), overlap_left=int(m[8]), overlap_right=int(m[9]))
    tile = srs.tiling_module.Tile(metadata=meta, data=None)
    assert np.array_equal(tm._create_blend_weight(tile), golden["merge_weight_4"])


def test_create_tile_metadata_reanchors(srs):
    tm = srs.TilingModule(64, 0.25, "mirror", enable_content_aware=False)
    tile = srs.tiling_module.Tile(metadata=srs.tiling_module.TileMetadata(), data=None)
    before = tile.metadata.updated_at
    m = tm.create_tile_metadata(tile, 96, 48)  # tiling_module.py:648-669
    assert m is tile.metadata and (m.global_x, m.global_y) == (96, 48) and m.updated_at >= before


def test_blend_weight_tables_equal_the_reference_map_random(srs, oracle):
    """the three 1-D tables the GPU merge uses reproduce _create_blend_weight's full map for any overlap geometry,
    including ramps that overlap each other (top + bottom > height)"""
    rng = np.random.default_rng(5)
    tm = srs.TilingModule(64, 0.25, "mirror", enable_content_aware=False)
    for _ in range(60):
        h, w = int(rng.integers(4, 90)), int(rng.integers(4, 90))
        ov = [int(rng.integers(0, h // 2 + 1)), int(rng.integers(0, h // 2 + 1)), int(rng.integers(0, w // 2 + 1)), int(rng.integers(0, w // 2 + 1))]
        if rng.random() < 0.3:
            ov[1] = int(rng.integers(0, h // 2 + 1)) + h // 2 - ov[0] // 2  # overlapping vertical ramps
            ov[1] = min(ov[1], h // 2)
        meta = srs.tiling_module.TileMetadata(output_w=w, output_h=h, overlap_top=ov[0], overlap_bottom=ov[1],
                                              overlap_left=ov[2], overlap_right=ov[3])
        tile = srs.tiling_module.Tile(metadata=meta, data=None)
        want = oracle.tiling_oracle.blend_weight(h, w, *[2 * v for v in ov])
        assert np.array_equal(tm._create_blend_weight(tile), want), (h, w, ov)
    big = srs.tiling_module.Tile(metadata=srs.tiling_module.TileMetadata(output_w=10, output_h=10, overlap_top=8), data=None)
    with pytest.raises(ValueError):  # 16 scaled ramp rows on a 10-row tile: the reference's ramp does not broadcast either
        tm._create_blend_weight(big)


def _check_gather_plan(srs, sizes, pos, out_hw, levels, kind, row_range):
    """independent restatement of what every class of the gather's plan PROMISES about each block of a cell"""
    FB_W, FB_H = 128, 16
    H, W = out_hw
    bm = srs.BlendingModule(num_levels=levels)
    plan = bm.gather_plan(sizes, pos, out_hw, weight_type=srs.WeightType(kind), row_range=row_range)
    rb, re = (0, H) if row_range is None else (max(row_range[0], 0), min(row_range[1], H))
    nbx, nby = -(-W // FB_W), -(-(re - rb) // FB_H)
    lv1 = [((h + 1) // 2, (w + 1) // 2) for h, w in sizes]
    if plan is None:
        return None
    seen = np.zeros((nby, nbx), np.int32)
    for c in plan:
        seen[c["by0"]:c["by0"] + c["nby"], c["bx0"]:c["bx0"] + c["nbx"]] += 1
        T, cls = len(c["tiles"]), c["cls"]
        assert {0: T == 1, 1: T == 1, 2: T == 2, 3: T in (3, 4), 4: T <= 2, 5: T in (3, 4), 6: 1 <= T <= 2, 7: T in (3, 4)}[cls], c
        # the corner blocks and a few inner ones decide (every condition is an interval per axis)
        bys = sorted({c["by0"], c["by0"] + c["nby"] - 1, c["by0"] + c["nby"] // 2})
        bxs = sorted({c["bx0"], c["bx0"] + c["nbx"] - 1, c["bx0"] + c["nbx"] // 2})
        for by in bys:
            for bx in bxs:
                cy0, cx0 = rb + by * FB_H, bx * FB_W
                cy1, cx1 = min(cy0 + FB_H, re), min(cx0 + FB_W, W)
                cover = [t for t, ((h, w), (y, x)) in enumerate(zip(sizes, pos)) if y < cy1 and y + h > cy0 and x < cx1 and x + w > cx0]
                assert cover == c["tiles"], (c, by, bx, cover)
                for t in cover:
                    (h, w), (y, x), (h1, w1) = sizes[t], pos[t], lv1[t]
                    tyf, txf = cy0 - y, cx0 - x
                    even = tyf % 2 == 0 and txf % 2 == 0
                    if cls <= 3:  # interior: whole block inside the tile, every level-1 window strictly inside the plane
                        assert even and cy1 == cy0 + FB_H and cx1 == cx0 + FB_W, (c, t)
                        assert tyf >= 2 and txf >= 2 and y + h >= cy0 + FB_H and x + w >= cx0 + FB_W, (c, t)
                        assert txf // 2 + FB_W // 2 + 1 <= w1 and tyf // 2 + FB_H // 2 + 1 <= h1, (c, t)
                    if cls == 0:  # unit weight over the whole block
                        F = min(h, w) // 8
                        assert tyf >= F and txf >= F and (y + h - 1) - (cy1 - 1) >= F and (x + w - 1) - (cx1 - 1) >= F, (c, t)
                        assert kind != "sigmoid"  # its profile never reaches exactly 1.0
                    if cls in (6, 7):  # edge cells on the interior arithmetic: even geometry, half-size level-1 plane
                        assert even and h == 2 * h1 and w == 2 * w1 and h1 >= 2 and w1 >= 2, (c, t)
    assert (seen == 1).all(), "every block of the band belongs to exactly one cell"
    firsts = {}
    for c in plan:  # `first` numbers the blocks of a class consecutively (a CTA finds its cell by searching it)
        firsts.setdefault(c["cls"], []).append((c["first"], c["nbx"] * c["nby"]))
    for cls, lst in firsts.items():
        nxt = 0
        for first, n in lst:
            assert first == nxt, (cls, lst)
            nxt += n
    return plan


def test_gather_plan_keeps_its_promises_on_random_geometry(srs):
    """the host-side planner of the canvas gather (include/srfuse.h: srf_gather_plan; no GPU involved): on random grids,
    bands, tile sizes and parities every block is in exactly one cell, a cell's tile list is the block's coverage in list
    order, and a cell's class never claims more than holds for its blocks"""
    rng = np.random.default_rng(5)
    classes = set()
    for case in range(60):
        th, tw = int(rng.integers(300, 1500)), int(rng.integers(300, 1500))
        if case % 3:
            th, tw = th & ~1, tw & ~1
        ny, nx = int(rng.integers(1, 4)), int(rng.integers(1, 4))
        sy, sx = int(th * rng.uniform(0.6, 1.0)), int(tw * rng.uniform(0.6, 1.0))
        if case % 2:
            sy, sx = sy & ~1, sx & ~1
        oy, ox = (0, 0) if case % 4 else (int(rng.integers(0, 40)), int(rng.integers(0, 300)))
        pos = [(oy + j * sy, ox + i * sx) for j in range(ny) for i in range(nx)]
        sizes = [(th, tw)] * len(pos)
        H = oy + (ny - 1) * sy + th - int(rng.integers(0, 60))
        W = ox + (nx - 1) * sx + tw - int(rng.integers(0, 60))
        band = None
        if case % 5 == 0:
            a = int(rng.integers(0, H - 1)) & ~1
            band = (a, int(rng.integers(a + 1, H + 1)))
        plan = _check_gather_plan(srs, sizes, pos, (H, W), int(rng.integers(3, 7)), ["cosine", "linear", "sigmoid"][case % 3], band)
        if plan:
            classes |= {c["cls"] for c in plan}
    assert classes >= {0, 1, 2, 3, 4, 6, 7}, classes  # the random cases reach every class but the rare generic 3/4-tile one


def test_gather_plan_of_the_200mp_config(srs):
    import collections
    pos = [(y, x) for y in (0, 3278, 6556, 9834) for x in range(0, 6 * 3278, 3278)]
    plan = _check_gather_plan(srs, [(4096, 4096)] * 24, pos, (11546, 17320), 7, "cosine", None)
    blocks = collections.Counter()
    for c in plan:
        blocks[c["cls"]] += c["nbx"] * c["nby"]
    # unit-weight cores, interior 1 / 2 / 4-tile cells, edge cells; nothing is left to the generic kernel
    assert dict(blocks) == {0: 50635, 1: 4637, 2: 29364, 3: 3900, 6: 7940, 7: 1716}
    assert sum(blocks.values()) == -(-17320 // 128) * -(-11546 // 16)
    band = _check_gather_plan(srs, [(4096, 4096)] * 24, pos, (11546, 17320), 7, "cosine", (4438, 6212))  # a band of an 8-rank job
    assert band and not any(c["cls"] in (4, 5) for c in band)
    assert srs.BlendingModule(num_levels=2).gather_plan([(256, 256)], [(0, 0)], (256, 256)) is None  # 2 levels: general kernel
