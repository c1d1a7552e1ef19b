// srf_iqa.cu -- full-reference PSNR + SSIM in one pass over an image pair (replaces
// QualityAssessmentModule.calculate_psnr / calculate_ssim, /root/reference/quality_assessment_module.py:277-320,
// 322-389; arithmetic of scikit-image's peak_signal_noise_ratio / structural_similarity: 11x11 Gaussian window with
// population covariance for multiscale=True, the default 7x7 uniform window with sample covariance otherwise).
//
// Each CTA owns a 256-column strip x P.rb rows of the S map and streams DOWN the rows: per input row it
//   * reads the RGB bytes of both images once (sum of squared differences and the OR of the bytes are taken here),
//   * converts to gray with cv2's 15-bit fixed point, keeps (gray - 128) as float in a double-buffered smem row,
//   * runs the horizontal pass for 4 signals {a, b, a^2 + b^2, a*b} (SSIM needs vx + vy, never each alone),
//   * pushes the 4 results into a register shift ring and runs the vertical pass from registers.
// 6 B/px of HBM traffic against ~130 FP32 instructions/px: FP32-issue bound, not HBM bound.
// Per-CTA partials are reduced by a second single-CTA kernel in a fixed order, so results are deterministic.
#include "srf_common.cuh"

namespace {

constexpr int IQ_THREADS = 128;
constexpr int IQ_RB_MAX = 128; // S-map rows per CTA: 256, 128 or 64 (srf_iqa_enqueue picks by grid size)
constexpr int IQ_R = 5;      // staging halo = radius of the widest window (11-tap Gaussian)

// exp(-k^2 / (2 * 1.5^2)) normalised, k = -5..5 (scipy.ndimage._gaussian_kernel1d(1.5, 0, 5))
#define IQ_T0 1.028380084e-03f
#define IQ_T1 7.598758135e-03f
#define IQ_T2 3.600077213e-02f
#define IQ_T3 1.093606895e-01f
#define IQ_T4 2.130055377e-01f
#define IQ_T5 2.660117249e-01f

// window taps: R = 5 -> exp(-k^2 / (2 * 1.5^2)) normalised (the literals above); R = 3 -> uniform 1/7
template <int R>
__device__ __forceinline__ constexpr float iq_tap(int k)
{
    if (R == 3) return 1.f / 7.f;
    const int d = k < 5 ? 5 - k : k - 5;
    return d == 0 ? IQ_T5 : d == 1 ? IQ_T4 : d == 2 ? IQ_T3 : d == 3 ? IQ_T2 : d == 4 ? IQ_T1 : IQ_T0;
}

struct IqaPartial {
    double ssim;
    unsigned long long sq;
    unsigned long long cnt;
    int max_a, max_b; // bitwise OR of every owned byte: <= 1 exactly when the image's max() <= 1
};

struct IqaPair { // one image pair of a batched launch (grid.z = pair)
    const uint8_t *a, *b;
    size_t pitch_a, pitch_b;
};

struct IqaParams {
    const uint8_t *a, *b;
    size_t pitch_a, pitch_b;
    const IqaPair *pairs; // batched launch: pair blockIdx.z replaces (a, b, pitch_a, pitch_b); nullptr otherwise
    int h, w;
    int scale_a, scale_b;
    float c1, c2;
    int rb;         // S-map rows per CTA
    float cov_norm; // 1 (population covariance, Gaussian window) or 49/48 (sample covariance, uniform 7x7 window)
    int row_begin, row_end;
    IqaPartial *partials;
};


// ---- packed FP32 (sm_100: FFMA2 / FMUL2 issue two IEEE-rounded lanes per instruction; the kernel is issue bound) ----
__device__ __forceinline__ float2 f2fma(float2 a, float w, float2 c) { return __ffma2_rn(a, make_float2(w, w), c); }
__device__ __forceinline__ float2 f2mul(float2 a, float w) { return __fmul2_rn(a, make_float2(w, w)); }

// the 2*C bytes of a pixel pair that starts `byte_off` bytes into a 16-byte aligned staged row, shifted down to bit 0:
// lo = bytes 0..3, hi = bytes 4..7 (C = 3: R0 G0 B0 R1 | G1 B1 x x; C = 1: Y0 Y1 x x | -)
template <int C, bool ANYSCALE>
__device__ __forceinline__ void iq_load_pair(const uint8_t *row, int byte_off, int scale, unsigned &lo, unsigned &hi)
{
    const unsigned *wp = reinterpret_cast<const unsigned *>(row) + (byte_off >> 2);
    const int sh = byte_off * 8; // the funnel shift uses the low five bits: (byte_off & 3) * 8
    const unsigned w0 = wp[0], w1 = wp[1];
    lo = __funnelshift_r(w0, w1, sh);
    hi = 0u;
    if (C == 3) hi = __funnelshift_r(w1, wp[2], sh);
    if (ANYSCALE && scale) { // (image * 255).astype(uint8), quality_assessment_module.py:192-193: (t * 255) & 255 == (-t) & 255
        lo = __vsub4(0u, lo);
        hi = __vsub4(0u, hi);
    }
}

// gray - 128 of both pixels.  cv2 COLOR_RGB2GRAY (uint8): (R*9798 + G*19235 + B*3735 + 16384) >> 15; the coefficients
// are split into byte halves so that two IDP.4A per pixel form the sum; the -128 rides in the rounding constant
// (128 << 15 is a multiple of the shift, so the arithmetic shift gives gray - 128 exactly)
#ifndef SRF_IQ_DP2A
#define SRF_IQ_DP2A 1 // gray from two IDP.2A (16-bit coefficients) instead of two IDP.4A + shift-add: 1.287 -> 1.280 ms at 200 MP
#endif
template <int C>
__device__ __forceinline__ void iq_gray_pair(unsigned lo, unsigned hi, int &g0, int &g1)
{
    if (C == 3) {
        constexpr unsigned CLO = 70u | (35u << 8) | (151u << 16), CHI = 38u | (75u << 8) | (14u << 16);
        constexpr unsigned K = 16384u - (128u << 15);
        const unsigned p1 = __funnelshift_r(lo, hi, 24);
#if SRF_IQ_DP2A
        // 16-bit coefficients against the bytes: (R, G) with the low byte pair, (B, -) with the high one -- no shift-add
        constexpr unsigned CRG = 9798u | (19235u << 16), CB = 3735u;
        const int x0 = (int)__dp2a_hi(CB, lo, __dp2a_lo(CRG, lo, K));
        const int x1 = (int)__dp2a_hi(CB, p1, __dp2a_lo(CRG, p1, K));
        (void)CLO;
        (void)CHI;
#else
        const int x0 = (int)(__dp4a(lo, CLO, K) + (__dp4a(lo, CHI, 0u) << 8));
        const int x1 = (int)(__dp4a(p1, CLO, K) + (__dp4a(p1, CHI, 0u) << 8));
#endif
        g0 = x0 >> 15;
        g1 = x1 >> 15;
    } else {
        g0 = (int)(lo & 0xFFu) - 128;
        g1 = (int)((lo >> 8) & 0xFFu) - 128;
    }
}

// ------------------------------------------------------------------------------------------------------------
// k_psnr_ssim: per CTA 256 S-map columns x P.rb rows, streamed DOWN the rows with one barrier per row:
//   L(r+4) cp.async (LDGSTS) 16-byte copies of the raw bytes of both images into a 4-stage smem ring (104 loader
//          threads, no registers held across the latency; image-edge chunks take a guarded byte path)
//   G(r+1) raw bytes -> (gray_a - 128, gray_b - 128) float pairs, one pixel PAIR per thread from packed words (funnel
//          shifts + IDP.4A); the exact squared error (VABSDIFF4 + IDP.4A) and the OR of the bytes of the owned pixels
//   H(r)   horizontal taps of the packed signals (a, b) and (a^2+b^2, ab) -- SSIM needs vx + vy, never each alone --
//          for the thread's TWO S-map columns: FFMA2, two lanes per instruction
//   V(r)   vertical pass in ACCUMULATE form: the new horizontal row h adds h * tap[k] to the 2R+1 open output rows,
//          A[j] = fma(h, tap, A[j + 1]) -- the shift of the classic register ring is the FMA's destination, so there
//          are no moves; per output row the summation order is oldest row first, as in a gather
// ANYSCALE: an image is rescaled by 255 first (the reference's max() <= 1 rule, P.scale_a / P.scale_b) -- a separate
// instantiation so the common case carries no dead arithmetic.  R: window radius (5 Gaussian, 3 uniform).
// ------------------------------------------------------------------------------------------------------------
constexpr int IQ2_COLS = 2 * IQ_THREADS;                 // 256 S-map columns per CTA
constexpr int IQ2_PX = IQ2_COLS + 2 * IQ_R;              // 266 staged pixels per row
constexpr int IQ2_RAW = ((15 + IQ2_PX * 3 + 15) / 16) * 16; // 832 raw bytes per row and image
constexpr int IQ2_SROW = IQ2_PX + 6;                     // 272 (a, b) pairs per row

template <int C, bool ANYSCALE, int R>
__global__ void __launch_bounds__(IQ_THREADS, 4) k_psnr_ssim(IqaParams P)
{
    pdl_enter(); // programmatic dependent launch: nothing above this line may touch global data
    const uint8_t *img_a = P.a, *img_b = P.b;
    size_t pitch_a = P.pitch_a, pitch_b = P.pitch_b;
    if (P.pairs != nullptr) { // batched launch (QualityAssessmentModule.batch_evaluate): every z-slice is its own image pair
        const IqaPair q = P.pairs[blockIdx.z];
        img_a = q.a;
        img_b = q.b;
        pitch_a = q.pitch_a;
        pitch_b = q.pitch_b;
    }
    __shared__ __align__(16) uint8_t s_raw[4][2][IQ2_RAW];
    __shared__ __align__(16) float2 s_ab[2][IQ2_SROW]; // (s, d) = (a + b, a - b), a / b = gray - 128 of the two images, per staged pixel
#ifndef SRF_IQ_BULK
#define SRF_IQ_BULK 0 // 1: rows that cannot touch either end of the readable range arrive as two 832-byte TMA bulk copies issued by ONE
                      // thread (mbarrier per stage) instead of 104 threads' 16-byte cp.async chunks; 0: chunks only.  MEASURED SLOWER
                      // (200 MP: 1.416 ms against 1.287): every thread polls the stage's mbarrier and the issuing warp runs a
                      // 50-instruction elect / R2UR sequence per row inside a loop that is barrier-coupled already
#endif
    __shared__ __align__(8) uint64_t s_full[4];
    __shared__ double s_red[IQ_THREADS / 32];
    __shared__ unsigned long long s_sq[IQ_THREADS / 32];
    __shared__ unsigned long long s_cnt[IQ_THREADS / 32];
    __shared__ int s_mx[2][IQ_THREADS / 32];

    constexpr int TAPS = 2 * R + 1;
    const int t = threadIdx.x;
    const int cx0 = blockIdx.x * IQ2_COLS;
    const int ry0 = P.row_begin + blockIdx.y * P.rb;
    const int ry1 = min(ry0 + P.rb, P.row_end);
    const int h = P.h, w = P.w;
    const int r_first = ry0 - R, r_last = ry1 + R; // input rows [r_first, r_last)
    const long long col_off = (long long)(cx0 - IQ_R) * C;

    // ---- loader role: threads 0..51 -> image a, 52..103 -> image b, one 16-byte chunk each per row ----
    constexpr int NCH = IQ2_RAW / 16;
    const bool loader = t < 2 * NCH;
    const int limg = (t >= NCH) ? 1 : 0, lchunk = t - limg * NCH;
    const uint8_t *lbase = limg ? img_b : img_a;
    const long long lpitch = (long long)(limg ? pitch_b : pitch_a);
    const int vr0 = max(P.row_begin - IQ_R, 0), vr1 = min(P.row_end + IQ_R, h);
    const uint8_t *llo = lbase + (long long)vr0 * lpitch;
    const uint8_t *lhi = lbase + (long long)(vr1 - 1) * lpitch + (long long)w * C;
    // a chunk of row r starts at most 30 bytes before the row and ends at most col_off + IQ2_RAW bytes into it: rows in
    // [safe_lo, safe_hi] can touch neither end of the readable range [llo, lhi) and skip the per-chunk guards
    const long long over = col_off + IQ2_RAW - (long long)w * C;
    const int safe_lo = vr0 + (int)((32 + lpitch - 1) / lpitch);
    const int safe_hi = vr1 - 1 - (over > 0 ? (int)((over + lpitch - 1) / lpitch) : 0);
    // rows every chunk of BOTH images may be copied blindly (uniform over the CTA): those travel as bulk copies
    auto safe_range = [&](long long pitch, int &lo, int &hi) {
        lo = vr0 + (int)((32 + pitch - 1) / pitch);
        hi = vr1 - 1 - (over > 0 ? (int)((over + pitch - 1) / pitch) : 0);
    };
    int sla, sha, slb, shb;
    safe_range((long long)pitch_a, sla, sha);
    safe_range((long long)pitch_b, slb, shb);
    const int bulk_lo = SRF_IQ_BULK ? max(max(sla, slb), r_first) : 0x7fffffff, bulk_hi = min(min(sha, shb), r_last - 1);
    if (SRF_IQ_BULK) {
        if (t == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) mbar_init(&s_full[i], 1);
            mbar_fence_init();
        }
        __syncthreads();
    }
    constexpr int BULK_T = 96; // the issuing thread: lane 0 of the last warp (warp 0 converts the extra halo pairs)
    // bulk rows are consecutive: the issuing thread tracks its two source pointers, nobody else carries loader state
    const uint8_t *ba_ptr = img_a + (long long)bulk_lo * (long long)pitch_a + col_off;
    const uint8_t *bb_ptr = img_b + (long long)bulk_lo * (long long)pitch_b + col_off;
    auto L = [&](int r) {
        if (SRF_IQ_BULK && r >= bulk_lo && r <= bulk_hi) {
            if (t == BULK_T) {
                const int st = (r - r_first) & 3;
                mbar_expect_tx(&s_full[st], 2 * IQ2_RAW);
                tma_bulk_g2s(&s_raw[st][0][0], ba_ptr - (reinterpret_cast<uintptr_t>(ba_ptr) & 15), IQ2_RAW, &s_full[st]);
                tma_bulk_g2s(&s_raw[st][1][0], bb_ptr - (reinterpret_cast<uintptr_t>(bb_ptr) & 15), IQ2_RAW, &s_full[st]);
                ba_ptr += pitch_a;
                bb_ptr += pitch_b;
            }
        } else if (loader && r >= vr0 && r < vr1 && r < r_last) { // the few rows next to an end of the readable range: guarded chunks
            const uint8_t *l_ptr = lbase + (long long)r * lpitch + col_off + 16 * lchunk;
            const uint8_t *q = l_ptr - (reinterpret_cast<uintptr_t>(l_ptr) & 15);
            uint8_t *dst = &s_raw[(r - r_first) & 3][limg][16 * lchunk];
            if (r >= safe_lo && r <= safe_hi) {
                cp_async_16(dst, q);
            } else if (q >= llo && q + 16 <= lhi) {
                cp_async_16(dst, q);
            } else if (q + 16 > llo && q < lhi) { // image edge: guarded bytes
                uint4 v = make_uint4(0, 0, 0, 0);
                uint8_t *bv = reinterpret_cast<uint8_t *>(&v);
#pragma unroll 1
                for (int e = 0; e < 16; ++e)
                    if (q + e >= llo && q + e < lhi) bv[e] = q[e];
                *reinterpret_cast<uint4 *>(dst) = v;
            }
        }
        cp_async_commit();
    };
    // a row that travelled as bulk copies is complete when its stage's mbarrier flips: bulk rows are consecutive, so the
    // k-th of them is the (k / 4)-th use of its stage
    auto wait_row = [&](int r) {
        if (SRF_IQ_BULK && r >= bulk_lo && r <= bulk_hi) mbar_wait(&s_full[(r - r_first) & 3], (unsigned)(((r - bulk_lo) >> 2) & 1));
    };

    unsigned int sq = 0; // <= 266 rows * 4 px * 3 * 65025 < 2^32
    unsigned int bits_a = 0, bits_b = 0;
    const int mis_a0 = (int)((reinterpret_cast<uintptr_t>(img_a) + (unsigned long long)col_off) & 15);
    const int mis_b0 = (int)((reinterpret_cast<uintptr_t>(img_b) + (unsigned long long)col_off) & 15);
    const int pa16 = (int)(pitch_a & 15), pb16 = (int)(pitch_b & 15);
    // staged pixel i sits at image column cx0 - 5 + i; thread t converts the pairs (2t, 2t+1) and, t < 5, (256+2t, 257+2t).
    // byte masks of the pixels this CTA OWNS (columns [cx0, cx0 + 256) inside the image) within a pair's lo / hi words
    auto own = [&](int i) { const int colp = cx0 - IQ_R + i; return colp >= cx0 && colp < cx0 + IQ2_COLS && colp < w; };
    auto pair_masks = [&](int i, unsigned &mlo, unsigned &mhi) {
        const bool o0 = own(i), o1 = own(i + 1);
        if (C == 3) {
            mlo = (o0 ? 0x00FFFFFFu : 0u) | (o1 ? 0xFF000000u : 0u);
            mhi = o1 ? 0x0000FFFFu : 0u;
        } else {
            mlo = (o0 ? 0x000000FFu : 0u) | (o1 ? 0x0000FF00u : 0u);
            mhi = 0u;
        }
    };
    unsigned mlo0, mhi0, mlo1 = 0u, mhi1 = 0u;
    pair_masks(2 * t, mlo0, mhi0);
    if (t < IQ_R) pair_masks(IQ2_COLS + 2 * t, mlo1, mhi1);
    asm volatile("" : "+r"(mlo0), "+r"(mhi0)); // per-thread constants: keep them, do not re-derive them from the thread index every row
    auto G = [&](int r) {
        const int buf = r & 1;
        float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0; // (s, d) of the pair's two pixels
        if (r >= 0 && r < h) {
            wait_row(r);
            const int st = (r - r_first) & 3;
            const int ma = (mis_a0 + r * pa16) & 15, mb = (mis_b0 + r * pb16) & 15;
            const bool rown = r >= ry0 && r < ry1;
            auto pair = [&](int i, unsigned mlo, unsigned mhi, float4 &o) {
                unsigned alo, ahi, blo, bhi;
                iq_load_pair<C, ANYSCALE>(s_raw[st][0], ma + i * C, P.scale_a, alo, ahi);
                iq_load_pair<C, ANYSCALE>(s_raw[st][1], mb + i * C, P.scale_b, blo, bhi);
                int a0, a1, b0, b1;
                iq_gray_pair<C>(alo, ahi, a0, a1);
                iq_gray_pair<C>(blo, bhi, b0, b1);
                o = make_float4((float)(a0 + b0), (float)(a0 - b0), (float)(a1 + b1), (float)(a1 - b1));
                if (rown) {
                    const unsigned dlo = __vabsdiffu4(alo, blo) & mlo;
                    sq = __dp4a(dlo, dlo, sq);
                    bits_a |= alo & mlo;
                    bits_b |= blo & mlo;
                    if (C == 3) {
                        const unsigned dhi = __vabsdiffu4(ahi, bhi) & mhi;
                        sq = __dp4a(dhi, dhi, sq);
                        bits_a |= ahi & mhi;
                        bits_b |= bhi & mhi;
                    }
                }
            };
            pair(2 * t, mlo0, mhi0, o0);
            if (t < IQ_R) pair(IQ2_COLS + 2 * t, mlo1, mhi1, o1);
        }
        *reinterpret_cast<float4 *>(&s_ab[buf][2 * t]) = o0;
        if (t < IQ_R) *reinterpret_cast<float4 *>(&s_ab[buf][IQ2_COLS + 2 * t]) = o1;
    };

    // open output rows: A[q][0][j] = partial (s, d) sums, A[q][1][j] = partial (s^2, d^2) sums of S-map column q
    float2 A[2][2][TAPS - 1];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int k = 0; k < TAPS - 1; ++k) A[q][s][k] = make_float2(0.f, 0.f);

    float ssum = 0.f;
    const int col0 = cx0 + 2 * t; // this thread's two S-map columns: col0, col0 + 1
    const bool ci0 = col0 >= R && col0 < w - R, ci1 = col0 + 1 >= R && col0 + 1 < w - R;
    // rows outside the map skip the pointwise stage altogether, columns outside it are skipped by a per-thread flag (a
    // predicated FFMA: S of a column outside the image need not be finite when the caller's data_range makes C1, C2 tiny);
    // the COUNT of accumulated values is rows x columns -- nothing is counted inside the loop
    int cf0 = ci0 ? 1 : 0, cf1 = ci1 ? 1 : 0;
    asm volatile("" : "+r"(cf0), "+r"(cf1)); // per-thread constants: kept in registers, not re-derived every row
    const float K1 = 2.f * P.c1, K2 = 2.f * P.c2;

    L(r_first);
    L(r_first + 1);
    L(r_first + 2);
    L(r_first + 3);
    cp_async_wait<2>();
    __syncthreads();
    G(r_first);
    __syncthreads();

    // ---- the stages of one row as lambdas; the loop below orders them ----
    // H: staged pixels 2t + (5 - R) .. + TAPS (TAPS + 1 of them, two per LDS.128): column 0 takes samples 0 .. TAPS-1,
    // column 1 samples 1 .. TAPS, both with taps 0 .. TAPS-1.  The second-order signals of a sample are ONE packed multiply
    // (s^2, d^2): a^2 + b^2 = (s^2 + d^2) / 2 and a b = (s^2 - d^2) / 4 are taken after the filters, which are linear
    auto H = [&](int buf, float2 (&hn)[2][2]) {
        const float4 *pv = reinterpret_cast<const float4 *>(&s_ab[buf][2 * t + (IQ_R - R)]);
        static_assert(((IQ_R - R) & 1) == 0 && ((TAPS + 1) & 1) == 0, "pairs must stay 16-byte aligned");
#pragma unroll
        for (int j = 0; j < (TAPS + 1) / 2; ++j) {
            const float4 v = pv[j];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int k = 2 * j + e;
                const float2 sd = e ? make_float2(v.z, v.w) : make_float2(v.x, v.y);
                const float2 sq2 = __fmul2_rn(sd, sd);
                if (k == 0) {
                    hn[0][0] = f2mul(sd, iq_tap<R>(0));
                    hn[0][1] = f2mul(sq2, iq_tap<R>(0));
                } else if (k < TAPS) {
                    hn[0][0] = f2fma(sd, iq_tap<R>(k), hn[0][0]);
                    hn[0][1] = f2fma(sq2, iq_tap<R>(k), hn[0][1]);
                }
                if (k == 1) {
                    hn[1][0] = f2mul(sd, iq_tap<R>(0));
                    hn[1][1] = f2mul(sq2, iq_tap<R>(0));
                } else if (k > 1) {
                    hn[1][0] = f2fma(sd, iq_tap<R>(k - 1), hn[1][0]);
                    hn[1][1] = f2fma(sq2, iq_tap<R>(k - 1), hn[1][1]);
                }
            }
        }
    };
    // V: vertical pass in accumulate form; f = the output row that closes with this input row
    auto V = [&](const float2 (&hn)[2][2], float2 (&f)[2][2]) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                f[q][s] = f2fma(hn[q][s], iq_tap<R>(TAPS - 1), A[q][s][0]);
#pragma unroll
                for (int j = 0; j < TAPS - 2; ++j) A[q][s][j] = f2fma(hn[q][s], iq_tap<R>(TAPS - 2 - j), A[q][s][j + 1]);
                A[q][s][TAPS - 2] = f2mul(hn[q][s], iq_tap<R>(0));
            }
        }
    };
    // PW: S of one column, accumulated when `take` is set.  With u = (mean s, mean d) and e = (mean s^2, mean d^2) of the
    // (x - 128) signals:
    //   ux + uy = u.s + 256, ux - uy = u.d;  2 ux uy = (p.s^2 - p.d^2) / 2, ux^2 + uy^2 = (p.s^2 + p.d^2) / 2
    //   v = e - u^2 = (var s, var d);       2 vxy = (v.s - v.d) / 2,     vx + vy = (v.s + v.d) / 2
    // The four factors of S are used DOUBLED (the 2s cancel in the quotient), so the constants enter as 2 C1, 2 C2.
    auto PW = [&](const float2 (&fq)[2], int take) {
        const float2 u = fq[0];
        const float2 pp0 = __fadd2_rn(u, make_float2(256.f, 0.f));
        const float2 pp = __fmul2_rn(pp0, pp0);
        float2 v = __ffma2_rn(make_float2(-u.x, -u.y), u, fq[1]);
        if (R == 3) v = f2mul(v, (float)(49.0 / 48.0)); // sample covariance for the 7x7 uniform window
        const float t1 = pp.x + K1, t2 = v.x + K2;
        const float A1 = t1 - pp.y, B1 = t1 + pp.y, A2 = t2 - v.y, B2 = t2 + v.y;
        float inv;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(B1 * B2)); // B1 B2 >= 4 C1 C2 = 1521: no range handling needed
        if (take) ssum = fmaf(A1 * A2, inv, ssum);
    };

#pragma unroll 1
    for (int r_in = r_first; r_in < r_last; ++r_in) {
        const int buf = r_in & 1;
        L(r_in + 4);
        float2 hn[2][2], f[2][2]; // this row's horizontal sums [q][(s, d) | (s^2, d^2)] and the output row they close
        const int r = r_in - R;
        const bool emit = r >= ry0 && r >= R && r < h - R; // uniform: an S-map row this CTA owns
        H(buf, hn);
        V(hn, f);
        if (emit) {
            PW(f[0], cf0);
            PW(f[1], cf1);
        }
        if (r_in + 1 < r_last) G(r_in + 1);
        cp_async_wait<2>();
        __syncthreads();
    }
    cp_async_wait<0>();

    double ds = (double)ssum;
    const int own_rows = max(min(ry1, h - R) - max(ry0, R), 0); // S-map rows this CTA accumulated
    unsigned long long dsq = sq, dcnt = (unsigned long long)own_rows * (unsigned)((ci0 ? 1 : 0) + (ci1 ? 1 : 0));
    int mxa = (int)((bits_a | (bits_a >> 8) | (bits_a >> 16) | (bits_a >> 24)) & 0xFFu);
    int mxb = (int)((bits_b | (bits_b >> 8) | (bits_b >> 16) | (bits_b >> 24)) & 0xFFu);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ds += __shfl_down_sync(0xffffffffu, ds, o);
        dsq += __shfl_down_sync(0xffffffffu, dsq, o);
        dcnt += __shfl_down_sync(0xffffffffu, dcnt, o);
        mxa |= __shfl_down_sync(0xffffffffu, mxa, o);
        mxb |= __shfl_down_sync(0xffffffffu, mxb, o);
    }
    if ((t & 31) == 0) {
        s_red[t >> 5] = ds;
        s_sq[t >> 5] = dsq;
        s_cnt[t >> 5] = dcnt;
        s_mx[0][t >> 5] = mxa;
        s_mx[1][t >> 5] = mxb;
    }
    __syncthreads();
    if (t == 0) {
        IqaPartial p;
        p.ssim = 0;
        p.sq = 0;
        p.cnt = 0;
        p.max_a = 0;
        p.max_b = 0;
        for (int k = 0; k < IQ_THREADS / 32; ++k) {
            p.ssim += s_red[k];
            p.sq += s_sq[k];
            p.cnt += s_cnt[k];
            p.max_a |= s_mx[0][k];
            p.max_b |= s_mx[1][k];
        }
        P.partials[(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = p;
    }
}

// single CTA, fixed summation order: thread k sums partials k, k+256, ... then a shared-memory tree
__global__ void __launch_bounds__(256) k_iqa_reduce(const IqaPartial *__restrict__ parts, int n, IqaPartial *__restrict__ out, double *__restrict__ dsum)
{
    pdl_enter(); // programmatic dependent launch: nothing above this line may touch global data
    __shared__ IqaPartial s[256];
    parts += (size_t)blockIdx.x * n; // batched: CTA z reduces the n partials of pair z
    out += blockIdx.x;
    if (dsum) dsum += 5 * blockIdx.x;
    IqaPartial p;
    p.ssim = 0;
    p.sq = 0;
    p.cnt = 0;
    p.max_a = 0;
    p.max_b = 0;
    for (int i = threadIdx.x; i < n; i += 256) {
        const IqaPartial q = parts[i];
        p.ssim += q.ssim;
        p.sq += q.sq;
        p.cnt += q.cnt;
        p.max_a |= q.max_a;
        p.max_b |= q.max_b;
    }
    s[threadIdx.x] = p;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            IqaPartial &x = s[threadIdx.x];
            const IqaPartial &y = s[threadIdx.x + o];
            x.ssim += y.ssim;
            x.sq += y.sq;
            x.cnt += y.cnt;
            x.max_a |= y.max_a;
            x.max_b |= y.max_b;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *out = s[0];
        if (dsum) { // additive form for a device-side all-reduce: {sq_err, ssim_sum, ssim_count, any byte of a > 1, of b > 1}
            dsum[0] = (double)s[0].sq;
            dsum[1] = s[0].ssim;
            dsum[2] = (double)s[0].cnt;
            dsum[3] = s[0].max_a > 1 ? 1.0 : 0.0;
            dsum[4] = s[0].max_b > 1 ? 1.0 : 0.0;
        }
    }
}

} // namespace

// one launch of k_psnr_ssim (+ the fixed-order reduction) over ONE pair (h_pairs == nullptr) or over n_pairs equally
// sized pairs (grid.z = pair; h_pairs are the host copies of the descriptors).  Per-pair results: IqaPartial n_pairs x
// at ws_iqa's tail (copied to ctx->pinned[slot ...] when slot >= 0) and, if d_sums != nullptr, 5 doubles per pair.
static int iqa_launch(srf_ctx *ctx, const uint8_t *a, size_t pa, const uint8_t *b, size_t pb, const IqaPair *h_pairs, int n_pairs, int h, int w,
                      int C, int scale_a, int scale_b, double data_range, int row_begin, int row_end, cudaStream_t st, int slot, double *d_sums,
                      int window)
{
    int rc;
    // S-map rows per CTA: tall CTAs amortise the vertical halo (2R rows re-filtered per CTA), but every CTA does the same
    // work, so the grid runs in lock-step waves of 4 CTAs per SM and a half-empty last wave costs a whole one.  Pick the
    // height that minimises (rows + halo) x waves (200 MP: 224 rows, 1.448 -> 1.426 ms).
    const int gx = srf_ceil_div(w, IQ2_COLS);
    const long long slots = 4LL * ctx->sm_count;
    int rb = IQ_RB_MAX;
    long long best = -1;
    for (int cand = 32; cand <= 2 * IQ_RB_MAX; cand += 2) {
        const long long ctas = (long long)gx * srf_ceil_div(row_end - row_begin, cand) * n_pairs;
        // fewer than three waves are charged as three: with one or two waves nothing overlaps the CTAs' fill and drain
        // (measured on 1 600-row bands: one wave of 208-row CTAs is slower than three waves of 64-row ones)
        const long long cost = (long long)(cand + 2 * IQ_R) * std::max<long long>((ctas + slots - 1) / slots, 3);
        if (best < 0 || cost <= best) {
            best = cost;
            rb = cand;
        }
    }
    const int gy = srf_ceil_div(row_end - row_begin, rb);
    const size_t nparts = (size_t)gx * gy; // per pair
    const size_t o_final = srf_align_up(nparts * n_pairs * sizeof(IqaPartial), 256);
    const size_t o_pairs = o_final + srf_align_up((size_t)n_pairs * sizeof(IqaPartial), 256);
    if ((rc = srf_reserve(ctx, ctx->ws_iqa, o_pairs + (size_t)n_pairs * sizeof(IqaPair)))) return rc;
    if ((rc = srf_reserve_pinned(ctx, std::max<size_t>(4096, (size_t)(std::max(slot, 0) + n_pairs) * sizeof(IqaPartial))))) return rc;
    IqaPartial *parts = (IqaPartial *)ctx->ws_iqa.ptr;
    IqaPartial *final_d = (IqaPartial *)((uint8_t *)ctx->ws_iqa.ptr + o_final);
    IqaPair *pairs_d = (IqaPair *)((uint8_t *)ctx->ws_iqa.ptr + o_pairs);
    if (h_pairs) // pageable source: staged before the call returns
        SRF_CHECK_CUDA(ctx, cudaMemcpyAsync(pairs_d, h_pairs, (size_t)n_pairs * sizeof(IqaPair), cudaMemcpyHostToDevice, st));
    IqaParams P;
    P.a = a;
    P.b = b;
    P.pitch_a = pa;
    P.pitch_b = pb;
    P.pairs = h_pairs ? pairs_d : nullptr;
    P.h = h;
    P.w = w;
    P.scale_a = scale_a;
    P.scale_b = scale_b;
    P.c1 = (float)((0.01 * data_range) * (0.01 * data_range));
    P.c2 = (float)((0.03 * data_range) * (0.03 * data_range));
    P.rb = rb;
    P.cov_norm = window ? (float)(49.0 / 48.0) : 1.f;
    P.row_begin = row_begin;
    P.row_end = row_end;
    P.partials = parts;
    if (nparts > 0) {
        dim3 grid(gx, gy, n_pairs);
        srf_launch_scope scope(ctx, "psnr_ssim", 2.0 * (double)(row_end - row_begin) * w * C * n_pairs, st);
        const bool any = scale_a || scale_b; // the rare max() <= 1 rescale: generic instantiation
#define IQ_LAUNCH(CC, ANY, RR) srf_launch_pdl(k_psnr_ssim<CC, ANY, RR>, grid, dim3(IQ_THREADS), 0, st, P)
        if (window == 0) {
            if (C == 3) { if (any) IQ_LAUNCH(3, true, 5); else IQ_LAUNCH(3, false, 5); }
            else        { if (any) IQ_LAUNCH(1, true, 5); else IQ_LAUNCH(1, false, 5); }
        } else {
            if (C == 3) { if (any) IQ_LAUNCH(3, true, 3); else IQ_LAUNCH(3, false, 3); }
            else        { if (any) IQ_LAUNCH(1, true, 3); else IQ_LAUNCH(1, false, 3); }
        }
#undef IQ_LAUNCH
    }
    {
        srf_launch_scope scope(ctx, "iqa_reduce", 0.0, st);
        srf_launch_pdl(k_iqa_reduce, dim3(n_pairs), dim3(256), 0, st, (const IqaPartial *)parts, (int)nparts, final_d, d_sums);
    }
    SRF_CHECK_CUDA(ctx, cudaGetLastError());
    if (slot >= 0)
        SRF_CHECK_CUDA(ctx, cudaMemcpyAsync((IqaPartial *)ctx->pinned + slot, final_d, (size_t)n_pairs * sizeof(IqaPartial), cudaMemcpyDeviceToHost, st));
    return SRF_OK;
}

int srf_iqa_enqueue(srf_ctx *ctx, const uint8_t *a, size_t pa, const uint8_t *b, size_t pb, int h, int w, int C, int scale_a, int scale_b,
                    double data_range, int row_begin, int row_end, cudaStream_t st, int slot, double *d_sums, int window)
{
    return iqa_launch(ctx, a, pa, b, pb, nullptr, 1, h, w, C, scale_a, scale_b, data_range, row_begin, row_end, st, slot, d_sums, window);
}

void srf_iqa_collect(srf_ctx *ctx, int w, int C, int row_begin, int row_end, srf_iqa_sums *out, int slot)
{
    const IqaPartial *f = (const IqaPartial *)ctx->pinned + slot;
    out->sq_err = (int64_t)f->sq;
    out->n_elems = (int64_t)(row_end - row_begin) * w * C;
    out->ssim_sum = f->ssim;
    out->ssim_count = (int64_t)f->cnt;
    out->bits_a = f->max_a;
    out->bits_b = f->max_b;
}

static int iqa_device(srf_ctx *ctx, const uint8_t *a, size_t pa, const uint8_t *b, size_t pb, int h, int w, int C, int scale_a, int scale_b,
                      double data_range, int row_begin, int row_end, srf_iqa_sums *out, cudaStream_t st, bool record_events, int window)
{
    if (record_events) SRF_CHECK_CUDA(ctx, cudaEventRecord(ctx->ev0, st));
    int rc = srf_iqa_enqueue(ctx, a, pa, b, pb, h, w, C, scale_a, scale_b, data_range, row_begin, row_end, st, 0, nullptr, window);
    if (rc) return rc;
    if (record_events) {
        SRF_CHECK_CUDA(ctx, cudaEventRecord(ctx->ev1, st));
        ctx->ev_valid = true;
    }
    SRF_CHECK_CUDA(ctx, cudaStreamSynchronize(st));
    srf_iqa_collect(ctx, w, C, row_begin, row_end, out, 0);
    return SRF_OK;
}

static int check_iqa_args(srf_ctx *ctx, const void *a, size_t pa, const void *b, size_t pb, int h, int w, int C, int row_begin, int row_end,
                          const void *out, int window)
{
    SRF_REQUIRE(ctx, a && b && out, "psnr_ssim: null pointer");
    SRF_REQUIRE(ctx, C == 1 || C == 3, "psnr_ssim: channels must be 1 or 3 (got %d)", C);
    SRF_REQUIRE(ctx, h >= 1 && w >= 1, "psnr_ssim: empty image");
    SRF_REQUIRE(ctx, pa >= (size_t)w * C && pb >= (size_t)w * C, "psnr_ssim: pitch < row bytes");
    SRF_REQUIRE(ctx, window == 0 || window == 1, "psnr_ssim: window must be 0 (11x11 Gaussian) or 1 (7x7 uniform), got %d", window);
    SRF_REQUIRE(ctx, row_begin >= 0 && row_begin <= row_end && row_end <= h, "psnr_ssim: bad row band [%d,%d) of %d", row_begin, row_end, h);
    return SRF_OK;
}

extern "C" int srf_psnr_ssim(srf_ctx *ctx, const uint8_t *d_a, size_t pitch_a, const uint8_t *d_b, size_t pitch_b, int h, int w,
                             int channels, int scale_a, int scale_b, double data_range, int window, int row_begin, int row_end, srf_iqa_sums *h_out, void *stream)
{
    if (!ctx) return SRF_ERR_ARG;
    int rc = check_iqa_args(ctx, d_a, pitch_a, d_b, pitch_b, h, w, channels, row_begin, row_end, h_out, window);
    if (rc) return rc;
    SRF_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    return iqa_device(ctx, d_a, pitch_a, d_b, pitch_b, h, w, channels, scale_a, scale_b, data_range, row_begin, row_end, h_out, (cudaStream_t)stream, true, window);
}

extern "C" int srf_psnr_ssim_host(srf_ctx *ctx, const uint8_t *h_a, size_t pitch_a, const uint8_t *h_b, size_t pitch_b, int h, int w,
                                  int channels, int scale_a, int scale_b, double data_range, int window, int row_begin, int row_end, srf_iqa_sums *h_out)
{
    if (!ctx) return SRF_ERR_ARG;
    int rc = check_iqa_args(ctx, h_a, pitch_a, h_b, pitch_b, h, w, channels, row_begin, row_end, h_out, window);
    if (rc) return rc;
    SRF_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (int hrc = srf_host_enter(ctx)) return hrc;
    // stage only the rows the band reads: [row_begin - 5, row_end + 5) clipped
    const int r0 = row_begin - IQ_R < 0 ? 0 : row_begin - IQ_R, r1 = row_end + IQ_R > h ? h : row_end + IQ_R;
    const size_t dp = srf_align_up((size_t)w * channels, 256);
    const size_t one = dp * (size_t)(r1 - r0 > 0 ? r1 - r0 : 1);
    if ((rc = srf_reserve(ctx, ctx->ws_stage_in, 2 * one))) return rc;
    cudaStream_t st = ctx->own_stream;
    uint8_t *da = (uint8_t *)ctx->ws_stage_in.ptr, *db = da + one;
    if (r1 > r0) {
        SRF_CHECK_CUDA(ctx, cudaMemcpy2DAsync(da, dp, h_a + (size_t)r0 * pitch_a, pitch_a, (size_t)w * channels, r1 - r0, cudaMemcpyHostToDevice, st));
        SRF_CHECK_CUDA(ctx, cudaMemcpy2DAsync(db, dp, h_b + (size_t)r0 * pitch_b, pitch_b, (size_t)w * channels, r1 - r0, cudaMemcpyHostToDevice, st));
    }
    ctx->last_h2d = 2 * (size_t)w * channels * (size_t)(r1 > r0 ? r1 - r0 : 0);
    ctx->last_d2h = sizeof(IqaPartial);
    return iqa_device(ctx, da - (size_t)r0 * dp, dp, db - (size_t)r0 * dp, dp, h, w, channels, scale_a, scale_b, data_range, row_begin, row_end, h_out, st, true, window);
}

extern "C" int srf_psnr_ssim_enqueue(srf_ctx *ctx, const uint8_t *d_a, size_t pitch_a, const uint8_t *d_b, size_t pitch_b, int h, int w,
                                     int channels, int scale_a, int scale_b, double data_range, int window, int row_begin, int row_end,
                                     double *d_sums5, void *stream)
{
    if (!ctx) return SRF_ERR_ARG;
    int rc = check_iqa_args(ctx, d_a, pitch_a, d_b, pitch_b, h, w, channels, row_begin, row_end, d_sums5, window);
    if (rc) return rc;
    SRF_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    rc = srf_iqa_enqueue(ctx, d_a, pitch_a, d_b, pitch_b, h, w, channels, scale_a, scale_b, data_range, row_begin, row_end,
                         (cudaStream_t)stream, -1, d_sums5, window);
    return rc ? rc : srf_note_device_work(ctx, (cudaStream_t)stream);
}

// ---- batched pairs (QualityAssessmentModule.batch_evaluate, quality_assessment_module.py:1350-1371: one result per pair) ----
static int check_batch_args(srf_ctx *ctx, const srf_image_pair *pairs, int n_pairs, int h, int w, int C, const void *out, int window)
{
    SRF_REQUIRE(ctx, pairs && out, "psnr_ssim_batch: null pointer");
    SRF_REQUIRE(ctx, n_pairs >= 1 && n_pairs <= 65535, "psnr_ssim_batch: n_pairs must be 1..65535 (got %d)", n_pairs);
    SRF_REQUIRE(ctx, C == 1 || C == 3, "psnr_ssim_batch: channels must be 1 or 3 (got %d)", C);
    SRF_REQUIRE(ctx, h >= 1 && w >= 1, "psnr_ssim_batch: empty image");
    SRF_REQUIRE(ctx, window == 0 || window == 1, "psnr_ssim_batch: window must be 0 (11x11 Gaussian) or 1 (7x7 uniform), got %d", window);
    for (int i = 0; i < n_pairs; ++i) {
        SRF_REQUIRE(ctx, pairs[i].a && pairs[i].b, "psnr_ssim_batch: pair %d has a null image", i);
        SRF_REQUIRE(ctx, pairs[i].pitch_a >= (size_t)w * C && pairs[i].pitch_b >= (size_t)w * C, "psnr_ssim_batch: pair %d pitch < row bytes", i);
    }
    return SRF_OK;
}

static void batch_collect(srf_ctx *ctx, int n_pairs, int h, int w, int C, srf_iqa_sums *out)
{
    for (int i = 0; i < n_pairs; ++i) srf_iqa_collect(ctx, w, C, 0, h, out + i, i);
}

extern "C" int srf_psnr_ssim_batch(srf_ctx *ctx, const srf_image_pair *pairs, int n_pairs, int h, int w, int channels, double data_range,
                                   int window, srf_iqa_sums *h_out, void *stream)
{
    if (!ctx) return SRF_ERR_ARG;
    int rc = check_batch_args(ctx, pairs, n_pairs, h, w, channels, h_out, window);
    if (rc) return rc;
    SRF_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    static_assert(sizeof(IqaPair) == sizeof(srf_image_pair), "descriptor layouts must agree");
    std::vector<IqaPair> hp(n_pairs);
    for (int i = 0; i < n_pairs; ++i) hp[i] = IqaPair{pairs[i].a, pairs[i].b, pairs[i].pitch_a, pairs[i].pitch_b};
    cudaStream_t st = (cudaStream_t)stream;
    if ((rc = iqa_launch(ctx, nullptr, 0, nullptr, 0, hp.data(), n_pairs, h, w, channels, 0, 0, data_range, 0, h, st, 0, nullptr, window))) return rc;
    SRF_CHECK_CUDA(ctx, cudaStreamSynchronize(st));
    batch_collect(ctx, n_pairs, h, w, channels, h_out);
    return SRF_OK;
}

extern "C" int srf_psnr_ssim_batch_enqueue(srf_ctx *ctx, const srf_image_pair *pairs, int n_pairs, int h, int w, int channels, double data_range,
                                           int window, double *d_sums5n, void *stream)
{
    if (!ctx) return SRF_ERR_ARG;
    int rc = check_batch_args(ctx, pairs, n_pairs, h, w, channels, d_sums5n, window);
    if (rc) return rc;
    SRF_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<IqaPair> hp(n_pairs);
    for (int i = 0; i < n_pairs; ++i) hp[i] = IqaPair{pairs[i].a, pairs[i].b, pairs[i].pitch_a, pairs[i].pitch_b};
    rc = iqa_launch(ctx, nullptr, 0, nullptr, 0, hp.data(), n_pairs, h, w, channels, 0, 0, data_range, 0, h, (cudaStream_t)stream, -1, d_sums5n,
                    window);
    return rc ? rc : srf_note_device_work(ctx, (cudaStream_t)stream);
}

extern "C" int srf_psnr_ssim_batch_host(srf_ctx *ctx, const srf_image_pair *pairs, int n_pairs, int h, int w, int channels, double data_range,
                                        int window, srf_iqa_sums *h_out)
{
    if (!ctx) return SRF_ERR_ARG;
    int rc = check_batch_args(ctx, pairs, n_pairs, h, w, channels, h_out, window);
    if (rc) return rc;
    SRF_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (int hrc = srf_host_enter(ctx)) return hrc;
    const size_t dp = srf_align_up((size_t)w * channels, 256), one = dp * (size_t)h;
    if ((rc = srf_reserve(ctx, ctx->ws_stage_in, 2 * one * (size_t)n_pairs))) return rc;
    cudaStream_t st = ctx->own_stream;
    uint8_t *base = (uint8_t *)ctx->ws_stage_in.ptr;
    std::vector<IqaPair> hp(n_pairs);
    for (int i = 0; i < n_pairs; ++i) {
        uint8_t *da = base + (size_t)(2 * i) * one, *db = da + one;
        SRF_CHECK_CUDA(ctx, cudaMemcpy2DAsync(da, dp, pairs[i].a, pairs[i].pitch_a, (size_t)w * channels, h, cudaMemcpyHostToDevice, st));
        SRF_CHECK_CUDA(ctx, cudaMemcpy2DAsync(db, dp, pairs[i].b, pairs[i].pitch_b, (size_t)w * channels, h, cudaMemcpyHostToDevice, st));
        hp[i] = IqaPair{da, db, dp, dp};
    }
    SRF_CHECK_CUDA(ctx, cudaEventRecord(ctx->ev0, st));
    if ((rc = iqa_launch(ctx, nullptr, 0, nullptr, 0, hp.data(), n_pairs, h, w, channels, 0, 0, data_range, 0, h, st, 0, nullptr, window))) return rc;
    SRF_CHECK_CUDA(ctx, cudaEventRecord(ctx->ev1, st));
    ctx->ev_valid = true;
    SRF_CHECK_CUDA(ctx, cudaStreamSynchronize(st));
    batch_collect(ctx, n_pairs, h, w, channels, h_out);
    ctx->last_h2d = 2 * (size_t)w * channels * h * (size_t)n_pairs;
    ctx->last_d2h = (size_t)n_pairs * sizeof(IqaPartial);
    return SRF_OK;
}
