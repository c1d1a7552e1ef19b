"""CPU suite, part 3: the N > 1 host logic under a 2-process gloo group -- band partition, tile selection and the
all-reduce of the PSNR / SSIM partial sums.  Each rank's partial sums come from the oracle restricted to its band
(the GPU kernel produces the same sums per band: tests/test_gpu_parity.py::test_iqa_bands_add_up)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import oracle
    import super_resolution_system_b200 as srs
    from helpers import synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(77)  # same image pair on every rank
        a = synth(rng, 151, 200)
        b = np.clip(a.astype(np.int16) + rng.integers(-9, 10, a.shape), 0, 255).astype(np.uint8)
        H, W = a.shape[:2]
        r0, r1 = srs.distributed.band(rank, world, H)
        io = oracle.iqa_oracle
        smap = io.ssim_map_interior(io.rgb_to_gray(a), io.rgb_to_gray(b))  # rows 5 .. H-6
        lo, hi = max(r0, 5), min(r1, H - 5)
        part = srs._native.srf_iqa_sums()
        part.sq_err = io.sq_err_sum(a[r0:r1], b[r0:r1])
        part.n_elems = (r1 - r0) * W * 3
        part.ssim_sum = float(smap[lo - 5:hi - 5].sum()) if hi > lo else 0.0
        part.ssim_count = max(hi - lo, 0) * (W - 10)
        part.bits_a = part.bits_b = 255
        total = srs.distributed.allreduce_sums(part, H * W * 3)
        psnr, ssim = srs.QualityAssessmentModule.finish(total)
        q.put((rank, (r0, r1), psnr, ssim, io.psnr(a, b), io.ssim(a, b)))
    finally:
        dist.destroy_process_group()


def test_two_rank_allreduce_matches_single_pass():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    bands = sorted(o[1] for o in out)
    assert bands == [(0, 76), (76, 151)]
    for _, _, psnr, ssim, want_p, want_s in out:
        assert psnr == pytest.approx(want_p, rel=1e-12)
        assert ssim == pytest.approx(want_s, rel=1e-12)


def test_band_partition_and_tile_selection(srs):
    d = srs.distributed
    for rows, world in [(11546, 8), (11546, 4), (11546, 2), (11546, 1), (10, 8), (7, 3)]:
        bands = [d.band(r, world, rows) for r in range(world)]
        assert bands[0][0] == 0 and bands[-1][1] == rows
        assert all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
        for r in range(world):
            f0, f1 = d.fuse_rows(r, world, rows)
            b0, b1 = bands[r]
            # band + 5-row SSIM halo, the start rounded down to an even row (parity of the gather's fast path)
            assert (f0, f1) == ((max(b0 - 5, 0) & ~1, min(b1 + 5, rows)) if b1 > b0 else (b0, b0))
    # the 200 MP grid: 6 x 4 tiles of 4096 rows at y = 0, 3278, 6556, 9834
    pos = [(y, x) for y in (0, 3278, 6556, 9834) for x in range(0, 6 * 3278, 3278)]
    hs = [4096] * len(pos)
    sel = d.tiles_for_rows(pos, hs, *d.fuse_rows(0, 8, 11546))
    assert sel == list(range(6))                      # band 0 (rows 0..1448) only sees the first tile row
    sel = d.tiles_for_rows(pos, hs, *d.fuse_rows(2, 8, 11546))
    assert sel == list(range(12))                     # rows 2883..3897 touch tile rows 0 and 1 (overlap zone)
    assert d.tiles_for_rows(pos, hs, 0, 11546) == list(range(24))


def test_balanced_bands_cover_and_balance(srs):
    d = srs.distributed
    pos = [(y, x) for y in (0, 3278, 6556, 9834) for x in range(0, 6 * 3278, 3278)]
    sizes = [(4096, 4096)] * len(pos)
    H, W = 11546, 17320

    def model_cost(band, row_cost=1.13, tile_cost=100.0):  # the documented cost model, restated
        r0, r1 = band
        c = row_cost * (r1 - r0)
        for (y, x), (h, w) in zip(pos, sizes):
            wc = max(0, min(W, x + w) - x)
            c += wc / W * max(0, min(r1, y + h, H) - max(r0, y))
            if r1 > r0 and y < r1 + 5 and y + h > r0 - 5:
                c += tile_cost * wc / 4096.0
        return c

    for world in (1, 2, 3, 4, 8):
        bands = d.balanced_bands(world, H, W, pos, sizes)
        assert len(bands) == world and bands[0][0] == 0 and bands[-1][1] == H
        assert all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
        costs = [model_cost(b) for b in bands]
        equal = [model_cost(d.band(r, world, H)) for r in range(world)]
        assert max(costs) <= max(equal) + 1e-6                     # never worse than equal-height bands
        assert max(costs) <= 1.10 * (sum(costs) / world)          # the per-tile constants make exact balance impossible
    b8 = d.balanced_bands(8, H, W, pos, sizes)
    assert b8[0][1] + 5 <= 3278                                    # band 0 stops short of grazing the second tile row
    assert d.balanced_bands(4, 3, 10, [(0, 0)], [(3, 10)])[-1][1] == 3  # more ranks than rows: empty bands allowed


def test_band_starts_are_even_and_bands_random_layouts(srs):
    """property checks of the band arithmetic on random tile layouts: the cuts tile the canvas, every fused range starts
    on an even row (the gather's fast path), and a rank's tile selection contains every tile its rows intersect"""
    d = srs.distributed
    rng = np.random.default_rng(7)
    for _ in range(40):
        H, W = int(rng.integers(200, 9000)), int(rng.integers(200, 9000))
        th, tw = int(rng.integers(64, 2048)), int(rng.integers(64, 2048))
        sy, sx = max(1, int(th * rng.uniform(0.7, 0.9))), max(1, int(tw * rng.uniform(0.7, 0.9)))
        pos = [(y, x) for y in range(0, H, sy) for x in range(0, W, sx)]
        sizes = [(th, tw)] * len(pos)
        for world in (1, 2, 3, 5, 8):
            bands = d.balanced_bands(world, H, W, pos, sizes)
            assert len(bands) == world and bands[0][0] == 0 and bands[-1][1] == H
            assert all(a[1] == b[0] and a[0] <= a[1] for a, b in zip(bands, bands[1:]))
            for b in bands:
                e0, e1 = d.widen(b, H)
                if b[1] > b[0]:
                    assert e0 % 2 == 0 and e0 <= max(b[0] - 5, 0) and e1 == min(b[1] + 5, H)
                    sel = d.tiles_for_rows(pos, [th] * len(pos), e0, e1)
                    assert sel == [i for i, (y, _) in enumerate(pos) if y < e1 and y + th > e0]


def test_fuse_input_rows_is_a_sufficient_cone(srs, oracle):
    """srf_fuse_input_rows (host arithmetic behind BandedHostFuser: which tile rows a band's ranks must hold): scrambling
    every tile pixel OUTSIDE the rows / columns it returns must not change one byte of the band in the reference algorithm,
    and the cone must stay near the reach SURVEY section 8c measured (about 3.25 * 2^(levels-1) rows a side; the planner
    rounds to whole rows of the coarsest level and to its kernels' row batches, so allow 6 * 2^(levels-1) + 8)"""
    import ctypes as C
    lib = srs._native.load()
    rng = np.random.default_rng(11)
    from helpers import synth
    for case in range(10):
        levels = int(rng.integers(2, 6))
        th, tw = int(rng.integers(90, 200)) & ~1, int(rng.integers(90, 200)) & ~1
        sy, sx = int(th * rng.uniform(0.6, 0.9)), int(tw * rng.uniform(0.6, 0.9))
        pos = [(j * sy, i * sx) for j in range(3) for i in range(2)]
        H, W = 2 * sy + th - int(rng.integers(0, 30)), sx + tw - int(rng.integers(0, 30))  # last row / column overhangs
        tiles = [synth(rng, th, tw) for _ in pos]
        rb = int(rng.integers(0, H - 20))
        re = int(rng.integers(rb + 1, H + 1))
        want = oracle.fusion_oracle.laplacian_fuse(tiles, pos, (H, W), levels, "cosine", "c")
        scrambled = []
        for t, (y, x) in zip(tiles, pos):
            a, b, c = C.c_int(), C.c_int(), C.c_int()
            assert lib.srf_fuse_input_rows(th, tw, y, x, levels, H, W, rb, re, C.byref(a), C.byref(b), C.byref(c)) == 0
            r0, r1, cols = a.value, b.value, c.value
            assert 0 <= r0 <= r1 <= th and 0 <= cols <= tw
            s = rng.integers(0, 256, t.shape, dtype=np.uint8)
            if r1 > r0:
                s[r0:r1, :cols] = t[r0:r1, :cols]
                lo, hi = max(rb, y) - y, min(re, y + th, H) - y  # the tile rows the band itself covers
                reach = 6 * 2 ** (levels - 1) + 8
                assert r0 >= lo - reach and r1 <= hi + reach, (case, levels, r0, r1, lo, hi)
            else:
                assert y >= re or y + th <= rb or x >= W
            scrambled.append(s)
        got = oracle.fusion_oracle.laplacian_fuse(scrambled, pos, (H, W), levels, "cosine", "c")
        assert np.array_equal(got[rb:re], want[rb:re]), (case, levels, rb, re)
