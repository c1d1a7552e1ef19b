/*
 * oracle/pyr_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C restatement of the two OpenCV primitives the reference's Laplacian fusion is built on:
 *   cv2.pyrDown(f32)                 called at /root/reference/blending_module.py:263
 *   cv2.pyrUp(f32, dstsize=(w,h))    called at /root/reference/blending_module.py:306, 358
 * The arithmetic lives in a third-party dependency that is not vendored in the reference
 * (requirements.txt: opencv-python>=4.8.0; the build probed here is opencv-python-headless 4.13.0.92,
 * x86-64 pip wheel, baseline SSE3 = 4 float lanes in imgproc/pyramids).  The float op ORDER below was
 * pinned empirically against that build (tests/test_oracle_pyr.py asserts bit equality with cv2 when cv2
 * is importable):
 *   pyrDown horizontal, vector body : r2*6 + ((r1+r3)*4 + (r0+r4))
 *   pyrDown horizontal, scalar cols : ((r2*6 + (r1+r3)*4) + r0) + r4      (col 0, cols >= vec_end)
 *   pyrDown vertical,   vector body : (((r1+r3)+r2)*4 + ((r0+r4)+(r2+r2))) * (1/256)
 *   pyrDown vertical,   scalar tail : (((r2*6 + (r1+r3)*4) + r0) + r4) * (1/256)   (float idx >= 4*floor(ow*cn/4))
 *   pyrUp even sample               : (p*6 + m) + n ; left edge  p*6 + n*2 ; right edge m + p*7
 *   pyrUp odd  sample               : (p + n)*4     ; right edge p*8
 *   pyrUp vertical uses the same even/odd forms with rows m=reflect101(j-1), n=min(j+1, h-1), then *(1/64)
 * Borders: pyrDown BORDER_REFLECT_101 on both axes; pyrUp reflect-101 low / replicate high.
 * Layout: HWC interleaved float32, like the NumPy arrays the reference passes to cv2.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -- contraction would change the rounding).
 */
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

static int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * (n - 1) - i;
    return i;
}

/* first output column (in pixels) that the 4-lane SIMD body of OpenCV's horizontal pass does NOT cover */
int orc_down_vec_end(int w, int ow, int cn)
{
    int width0 = (w - 3) / 2 + 1;            /* columns < width0 need no right border */
    if (width0 > ow) width0 = ow;
    if (width0 < 1) return 1;                /* everything is border -> scalar */
    if (cn == 1) return 1 + 4 * ((width0 - 1) / 4);
    if (cn == 3) return (width0 - 1 > 1) ? width0 - 1 : 1;
    if (cn == 2) return 1 + 2 * ((width0 - 1) / 2);
    if (cn == 4) return width0;
    return 1;
}

static inline float h_vec(float r0, float r1, float r2, float r3, float r4)
{
    float a = (r1 + r3) * 4.f;
    float b = r0 + r4;
    float c = a + b;
    return r2 * 6.f + c;
}
static inline float h_sca(float r0, float r1, float r2, float r3, float r4)
{
    float a = r2 * 6.f;
    float b = (r1 + r3) * 4.f;
    return ((a + b) + r0) + r4;
}
static inline float v_vec(float r0, float r1, float r2, float r3, float r4)
{
    float a = ((r1 + r3) + r2) * 4.f;
    float b = (r0 + r4) + (r2 + r2);
    return (a + b) * (1.f / 256.f);
}
static inline float v_sca(float r0, float r1, float r2, float r3, float r4)
{
    return h_sca(r0, r1, r2, r3, r4) * (1.f / 256.f);
}

/* dst is ((h+1)/2, (w+1)/2, cn) */
int orc_pyr_down_f32(const float *src, int h, int w, int cn, float *dst)
{
    const int oh = (h + 1) / 2, ow = (w + 1) / 2;
    const int vec_end = orc_down_vec_end(w, ow, cn);
    const int vtail = 4 * ((ow * cn) / 4);
    float *rows = (float *)malloc((size_t)h * ow * cn * sizeof(float));
    if (!rows) return -1;
    for (int y = 0; y < h; ++y) {
        const float *s = src + (size_t)y * w * cn;
        float *r = rows + (size_t)y * ow * cn;
        for (int x = 0; x < ow; ++x) {
            int i0 = reflect101(2 * x - 2, w), i1 = reflect101(2 * x - 1, w), i2 = reflect101(2 * x, w);
            int i3 = reflect101(2 * x + 1, w), i4 = reflect101(2 * x + 2, w);
            int scalar = (x < 1) || (x >= vec_end);
            for (int c = 0; c < cn; ++c) {
                float a0 = s[i0 * cn + c], a1 = s[i1 * cn + c], a2 = s[i2 * cn + c], a3 = s[i3 * cn + c],
                      a4 = s[i4 * cn + c];
                r[x * cn + c] = scalar ? h_sca(a0, a1, a2, a3, a4) : h_vec(a0, a1, a2, a3, a4);
            }
        }
    }
    for (int y = 0; y < oh; ++y) {
        const float *r0 = rows + (size_t)reflect101(2 * y - 2, h) * ow * cn;
        const float *r1 = rows + (size_t)reflect101(2 * y - 1, h) * ow * cn;
        const float *r2 = rows + (size_t)reflect101(2 * y, h) * ow * cn;
        const float *r3 = rows + (size_t)reflect101(2 * y + 1, h) * ow * cn;
        const float *r4 = rows + (size_t)reflect101(2 * y + 2, h) * ow * cn;
        float *d = dst + (size_t)y * ow * cn;
        for (int k = 0; k < ow * cn; ++k)
            d[k] = (k < vtail) ? v_vec(r0[k], r1[k], r2[k], r3[k], r4[k]) : v_sca(r0[k], r1[k], r2[k], r3[k], r4[k]);
    }
    free(rows);
    return 0;
}

/* horizontal pyrUp pass of one source row: out has 2*w samples per channel */
static void up_row(const float *s, int w, int cn, float *out)
{
    for (int c = 0; c < cn; ++c) {
        if (w == 1) {
            out[c] = s[c] * 8.f;
            out[cn + c] = s[c] * 8.f;
            continue;
        }
        for (int i = 0; i < w; ++i) {
            float p = s[i * cn + c];
            float e, o;
            if (i == 0) e = p * 6.f + s[cn + c] * 2.f;
            else if (i == w - 1) e = s[(i - 1) * cn + c] + p * 7.f;
            else e = (p * 6.f + s[(i - 1) * cn + c]) + s[(i + 1) * cn + c];
            if (i == w - 1) o = p * 8.f;
            else o = (p + s[(i + 1) * cn + c]) * 4.f;
            out[(2 * i) * cn + c] = e;
            out[(2 * i + 1) * cn + c] = o;
        }
    }
}

/* dst is (dh, dw, cn) with dh in {2h, 2h-1} and dw in {2w, 2w-1}: the only sizes the reference's pyramids
 * produce (G_{l+1} = ceil(G_l / 2)).  OpenCV also accepts 2n+1; that case is not on the path and is
 * rejected here (-2) rather than guessed at. */
int orc_pyr_up_f32(const float *src, int h, int w, int cn, float *dst, int dh, int dw)
{
    if (dh > 2 * h || dh < 2 * h - 1 || dw > 2 * w || dw < 2 * w - 1) return -2;
    const int rw = 2 * w;
    float *rows = (float *)malloc((size_t)h * rw * cn * sizeof(float));
    if (!rows) return -1;
    for (int y = 0; y < h; ++y) up_row(src + (size_t)y * w * cn, w, cn, rows + (size_t)y * rw * cn);
    for (int j = 0; j < h; ++j) {
        const float *rm = rows + (size_t)reflect101(j - 1, h) * rw * cn; /* h==1 -> row 0 */
        const float *rp = rows + (size_t)j * rw * cn;
        const float *rn = rows + (size_t)((j + 1 < h) ? j + 1 : h - 1) * rw * cn;
        for (int par = 0; par < 2; ++par) {
            int ty = 2 * j + par;
            if (ty >= dh) break;
            float *d = dst + (size_t)ty * dw * cn;
            for (int x = 0; x < dw; ++x) {
                int sx = x;
                for (int c = 0; c < cn; ++c) {
                    float p = rp[sx * cn + c], m = rm[sx * cn + c], n = rn[sx * cn + c];
                    float v = par ? (p + n) * 4.f : (p * 6.f + m) + n;
                    d[x * cn + c] = v * (1.f / 64.f);
                }
            }
        }
    }
    free(rows);
    return 0;
}
