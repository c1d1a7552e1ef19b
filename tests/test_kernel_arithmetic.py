"""CPU suite, part 5: the numerical arguments the kernels rely on, checked by emulating the device arithmetic in NumPy
(float32 / uint32 with the same operation order).  Nothing here touches the GPU; the GPU suite checks the kernels, this
file checks WHY their shortcuts are exact."""
import numpy as np

f32 = np.float32


def _ulp_step(x, k):
    """x moved by k float32 ulps"""
    return (x.view(np.int32) + np.int32(k)).view(np.float32)


def test_gather_epilogue_guard_never_changes_a_byte():
    """k_final epilogue (csrc/srf_fuse.cu): out = trunc(clip(RN(a / b))) is replaced by trunc_sat(a * rcp(b)) unless the
    product lies within FINAL_NEAR of an integer (then the IEEE division runs).  MUFU.RCP is within 1 ulp of 1/b; take
    the worst case of -1, 0, +1 ulp and require byte equality for every unflagged sample."""
    rng = np.random.default_rng(0)
    n = 2_000_000
    near = f32(2.5e-4)
    b = np.concatenate([rng.uniform(1e-6, 4.0, n // 2), rng.uniform(0.5, 2.5, n // 2)]).astype(f32)
    target = rng.uniform(-20, 290, n)                       # quotients around the byte range, incl. clipped ones
    target[: n // 4] = np.round(target[: n // 4]) + rng.uniform(-3e-4, 3e-4, n // 4)  # a quarter right next to integers
    a = (target * b.astype(np.float64)).astype(f32)
    exact = (a / b).astype(f32)                              # IEEE float32 division, what the reference computes
    want = np.clip(exact, 0, 255).astype(np.uint8)           # clip, then truncate
    r0 = (f32(1) / b).astype(f32)
    checked = 0
    for k in (-1, 0, 1):
        q = (a * _ulp_step(r0, k)).astype(f32)
        flagged = np.abs(q - np.rint(q)) < near
        got = np.clip(np.trunc(q), 0, 255).astype(np.uint8)  # cvt.rzi.u8.f32 saturates
        ok = ~flagged
        checked += int(ok.sum())
        assert np.array_equal(got[ok], want[ok]), f"rcp error {k} ulp"
    assert checked > n  # the guard leaves most samples on the fast path
    # power-of-two divisors: rcp is exact, the product IS the quotient -- also when it is an integer (singly covered pixels)
    for p2 in (0.5, 1.0, 2.0, 4.0):
        a2 = (np.round(rng.uniform(0, 255, 100000)) * p2).astype(f32)
        q = (a2 * (f32(1) / f32(p2))).astype(f32)
        assert np.array_equal(q, (a2 / f32(p2)).astype(f32))


def test_gray_from_packed_bytes_is_cv2_gray():
    """k_psnr_ssim (csrc/srf_iqa.cu): gray - 128 = ((dp4a(px, LO, K) + (dp4a(px, HI, 0) << 8)) as int32) >> 15 with the
    cv2 coefficients split into byte halves and K = 16384 - (128 << 15)"""
    cv2_ok = True
    try:
        import cv2
    except Exception:  # pragma: no cover
        cv2_ok = False
    rng = np.random.default_rng(1)
    rgb = rng.integers(0, 256, (400000, 3), dtype=np.uint8)
    rgb[:8] = [[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 255, 0], [0, 0, 255], [1, 1, 1], [254, 255, 253], [128, 127, 129]]
    lo, hi = (70, 35, 151), (38, 75, 14)
    assert [l + 256 * h for l, h in zip(lo, hi)] == [9798, 19235, 3735]
    p = rgb.astype(np.uint32)
    k = np.uint32((16384 - (128 << 15)) & 0xFFFFFFFF)
    x = (p[:, 0] * lo[0] + p[:, 1] * lo[1] + p[:, 2] * lo[2] + k).astype(np.uint32)
    y = (p[:, 0] * hi[0] + p[:, 1] * hi[1] + p[:, 2] * hi[2]).astype(np.uint32)
    g = ((x + (y << np.uint32(8))).astype(np.uint32).view(np.int32) >> 15).astype(np.int64)
    want = ((rgb[:, 0].astype(np.int64) * 9798 + rgb[:, 1].astype(np.int64) * 19235 + rgb[:, 2].astype(np.int64) * 3735 + 16384) >> 15) - 128
    assert np.array_equal(g, want)
    # the shipped form: two IDP.2A with the 16-bit coefficients -- (R, G) against the low byte pair, (B, next byte x 0) against the
    # high one -- accumulate the same 32-bit sum without the shift-add
    g2 = ((p[:, 0] * 9798 + p[:, 1] * 19235 + k).astype(np.uint32) + (p[:, 2] * 3735).astype(np.uint32)).astype(np.uint32).view(np.int32) >> 15
    assert np.array_equal(g2.astype(np.int64), want)
    if cv2_ok:
        assert np.array_equal(cv2.cvtColor(rgb[None], cv2.COLOR_RGB2GRAY)[0].astype(np.int64) - 128, want)


def test_pixel_pair_extraction_for_every_alignment():
    """iq_load_pair / iq_gray_pair: a pixel pair that starts `off` bytes into a 16-byte aligned row is rebuilt from three
    aligned 32-bit words by funnel shifts -- lo = bytes 0..3, hi = bytes 4..7, second pixel = funnel(lo, hi, 24)"""
    rng = np.random.default_rng(2)
    row = rng.integers(0, 256, 900, dtype=np.uint8)
    words = row[:896].view("<u4").astype(np.uint64)

    def funnel_r(lo, hi, sh):  # __funnelshift_r: low 32 bits of (hi:lo) >> (sh & 31)
        return int(((int(hi) << 32 | int(lo)) >> (sh & 31)) & 0xFFFFFFFF)

    for m in range(16):                     # row misalignment of the staged image
        for i in range(0, 266, 2):          # staged pixel pair (2t, 2t+1) and the halo pairs
            off = m + 3 * i
            w0, w1, w2 = words[off >> 2], words[(off >> 2) + 1], words[(off >> 2) + 2]
            sh = (off & 3) * 8
            lo, hi = funnel_r(w0, w1, sh), funnel_r(w1, w2, sh)
            p1 = funnel_r(lo, hi, 24)
            assert [lo & 255, lo >> 8 & 255, lo >> 16 & 255] == row[off:off + 3].tolist()
            assert [p1 & 255, p1 >> 8 & 255, p1 >> 16 & 255] == row[off + 3:off + 6].tolist()
            # ownership masks of the squared-error sum: pixel 0 = bytes 0..2 of lo, pixel 1 = byte 3 of lo + bytes 0..1 of hi
            assert [(lo >> 24) & 255, hi & 255, hi >> 8 & 255] == row[off + 3:off + 6].tolist()


def test_vertical_accumulate_form_keeps_the_summation_order():
    """k_psnr_ssim vertical pass: A[j] = fma(h, tap[2R-1-j], A[j+1]) row by row gives, for every output row, exactly the
    float32 sum  fma(h_{r+R}, t_{2R}, ... fma(h_{r-R+1}, t_1, h_{r-R} * t_0))  of the classic register-ring gather"""
    rng = np.random.default_rng(3)
    R = 5
    taps = np.exp(-np.arange(-R, R + 1) ** 2 / 4.5)
    taps = (taps / taps.sum()).astype(f32)
    h = rng.normal(0, 50, 64).astype(f32)

    def fma(a, b, c):  # float32 fma emulated in float64 (exact product, one rounding)
        return f32(np.float64(a) * np.float64(b) + np.float64(c))

    A = [f32(0)] * (2 * R)
    outs = {}
    for r_in, hv in enumerate(h):
        outs[r_in - R] = fma(hv, taps[2 * R], A[0])
        for j in range(2 * R - 1):
            A[j] = fma(hv, taps[2 * R - 1 - j], A[j + 1])
        A[2 * R - 1] = f32(hv * taps[0])
    for r in range(R, len(h) - R):
        v = f32(h[r - R] * taps[0])
        for k in range(1, 2 * R + 1):
            v = fma(h[r - R + k], taps[k], v)
        assert outs[r] == v


def test_sum_difference_form_of_ssim_in_float32_matches_the_float64_oracle():
    """k_psnr_ssim filters the packed signal (s, d) = (a + b, a - b) of the centred grays and its square (s^2, d^2) -- one
    packed multiply per sample instead of a^2 + b^2 and a b -- and rebuilds S from the four filtered values with every
    factor doubled:  2 ux uy + C1 ~ p.s^2 - p.d^2 + 2 C1,  ux^2 + uy^2 + C1 ~ p.s^2 + p.d^2 + 2 C1  (p = (u.s + 256, u.d)),
    2 vxy + C2 ~ v.s - v.d + 2 C2,  vx + vy + C2 ~ v.s + v.d + 2 C2  (v = e - u^2).  Emulated here in float32 (horizontal
    pass, then the vertical one, as the kernel) on textured, flat, saturated and anti-correlated pairs."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
    from oracle import iqa_oracle
    from helpers import synth

    taps64 = iqa_oracle.gaussian_taps()
    taps = taps64.astype(f32)
    R = 5

    def filt(x):  # float32, horizontal then vertical, mul + add in tap order
        h, w = x.shape
        t = np.zeros((h, w - 2 * R), f32)
        for k in range(2 * R + 1):
            t = (t + x[:, k:k + w - 2 * R] * taps[k]).astype(f32)
        o = np.zeros((h - 2 * R, w - 2 * R), f32)
        for k in range(2 * R + 1):
            o = (o + t[k:k + h - 2 * R] * taps[k]).astype(f32)
        return o

    def kernel_ssim(ga, gb, data_range=255.0):
        a, b = ga.astype(np.int32) - 128, gb.astype(np.int32) - 128
        s, d = (a + b).astype(f32), (a - b).astype(f32)
        us, ud = filt(s), filt(d)
        es, ed = filt((s * s).astype(f32)), filt((d * d).astype(f32))
        k1, k2 = f32(2) * f32((0.01 * data_range) ** 2), f32(2) * f32((0.03 * data_range) ** 2)
        ps = (us + f32(256)).astype(f32)
        pps, ppd = (ps * ps).astype(f32), (ud * ud).astype(f32)
        vs, vd = (es - us * us).astype(f32), (ed - ud * ud).astype(f32)
        t1, t2 = (pps + k1).astype(f32), (vs + k2).astype(f32)
        a1, b1, a2, b2 = t1 - ppd, t1 + ppd, t2 - vd, t2 + vd
        smap = ((a1 * a2).astype(f32) / (b1 * b2).astype(f32)).astype(f32)
        return float(smap.astype(np.float64).mean())

    rng = np.random.default_rng(11)
    x = synth(rng, 96, 160, 1)
    noisy = np.clip(x.astype(np.int32) + rng.integers(-12, 13, x.shape), 0, 255).astype(np.uint8)
    cases = {
        "textured": (x, noisy),
        "flat bright": (np.full_like(x, 255), np.full_like(x, 254)),
        "flat dark": (np.zeros_like(x), np.ones_like(x)),
        "saturated halves": (np.where(np.arange(160) < 80, 255, 0).astype(np.uint8)[None].repeat(96, 0), noisy),
        "negative": (x, 255 - x),
        "identical": (x, x.copy()),
        "checkerboard": ((np.indices(x.shape).sum(0) % 2 * 255).astype(np.uint8), x),
    }
    for name, (ga, gb) in cases.items():
        want = float(iqa_oracle.ssim_map_interior(ga, gb).mean(dtype=np.float64))
        got = kernel_ssim(ga, gb)
        assert abs(got - want) <= 2e-6 * max(1.0, abs(want)), (name, got, want)
