// srf_resize.cu -- cv2.resize(..., interpolation=cv2.INTER_CUBIC) on uint8 images: the arithmetic of
// QualityAssessmentModule.downsample_bicubic / upsample_bicubic (/root/reference/quality_assessment_module.py:226-275),
// which _evaluate_downsample_comparison (518-555) applies to both images at 0.1x / 0.2x / 0.4x before PSNR and the
// 7x7-window SSIM, and which is also the bicubic stand-in for the remote SR call.
//
// The opencv-python wheel routes uint8 INTER_CUBIC to Intel IPP, whose result is the Keys cubic (a = -0.75) evaluated at
// float64 sample positions, rounded half-to-even and saturated (oracle/resize_oracle.py: identical to cv2 4.13 on all but
// ~4e-6 of the pixels, bit-exact for the 2x up-scale).  The per-axis tap tables (4 clamped source indices + 4 weights
// per destination index) are built on the host in float64 with OpenCV's interpolateCubic formula and rounded to float32;
// the kernel gathers the 4 x 4 neighbourhood and sums rows first, then columns, every product and sum rounded on its
// own (__fmul_rn / __fadd_rn: no contraction), which is the oracle's float32 order bit for bit.
#include "srf_common.cuh"

#include <math.h>

namespace {

struct __align__(16) ResizeTap {
    int i[4];   // clamped source indices s-1 .. s+2
    float w[4]; // Keys weights of the fractional position
};

constexpr int RS_BX = 32, RS_BY = 8;

template <int C>
__global__ void __launch_bounds__(RS_BX *RS_BY) k_resize_cubic(const uint8_t *__restrict__ src, size_t spitch, uint8_t *__restrict__ dst,
                                                                size_t dpitch, int dh, int dw, const ResizeTap *__restrict__ tx,
                                                                const ResizeTap *__restrict__ ty)
{
    const int x = blockIdx.x * RS_BX + threadIdx.x, y = blockIdx.y * RS_BY + threadIdx.y;
    if (x >= dw || y >= dh) return;
    const int4 xi = *reinterpret_cast<const int4 *>(tx[x].i), yi = *reinterpret_cast<const int4 *>(ty[y].i);
    const float4 xw = *reinterpret_cast<const float4 *>(tx[x].w), yw = *reinterpret_cast<const float4 *>(ty[y].w);
    const uint8_t *r0 = src + (size_t)yi.x * spitch, *r1 = src + (size_t)yi.y * spitch;
    const uint8_t *r2 = src + (size_t)yi.z * spitch, *r3 = src + (size_t)yi.w * spitch;
    const int xo[4] = {xi.x * C, xi.y * C, xi.z * C, xi.w * C};
    const float wx[4] = {xw.x, xw.y, xw.z, xw.w};
    float o[C];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int b = xo[k] + c;
            float t = __fmul_rn((float)__ldg(r0 + b), yw.x); // rows first
            t = __fadd_rn(t, __fmul_rn((float)__ldg(r1 + b), yw.y));
            t = __fadd_rn(t, __fmul_rn((float)__ldg(r2 + b), yw.z));
            t = __fadd_rn(t, __fmul_rn((float)__ldg(r3 + b), yw.w));
            const float v = __fmul_rn(t, wx[k]);
            o[c] = (k == 0) ? v : __fadd_rn(o[c], v);
        }
    }
    uint8_t *out = dst + (size_t)y * dpitch + (size_t)x * C;
#pragma unroll
    for (int c = 0; c < C; ++c) out[c] = (uint8_t)min(max(__float2int_rn(o[c]), 0), 255); // rint, then saturate
}

// ------------------------------------------------------------------------------------------------------------
// k_resize_cubic_up: the same arithmetic for ENLARGING resizes (the bicubic SR stand-in: 2048^2 -> 4096^2 per tile), tiled.
// The general kernel above gathers 16 source bytes per output value straight from global memory and repeats the column
// sums of a source column for every output pixel that taps it (8 times at 2x).  Here a CTA stages the source patch of its
// 128 x 16 output block in shared memory once, and per output row forms the COLUMN sums T[sx] = sum_j src[yi_j][sx] * yw_j
// once per source column (the first, "rows first" half of the reference order), then every thread combines four of them
// with its x weights -- the identical products and sums, in the identical order, so the bytes equal the general kernel's.
// ------------------------------------------------------------------------------------------------------------
constexpr int RU_W = 128, RU_H = 16;             // output block
constexpr int RU_SW = RU_W + 8, RU_SH = RU_H + 8; // source patch capacity (scale >= 1: at most block + 4 taps, rounded up)

template <int C>
__global__ void __launch_bounds__(RU_W) k_resize_cubic_up(const uint8_t *__restrict__ src, size_t spitch, uint8_t *__restrict__ dst, size_t dpitch,
                                                           int dh, int dw, const ResizeTap *__restrict__ tx, const ResizeTap *__restrict__ ty)
{
    __shared__ uint8_t s_src[RU_SH][RU_SW * C];
    __shared__ float s_t[RU_SW * C];
    const int x0 = blockIdx.x * RU_W, y0 = blockIdx.y * RU_H;
    const int x1 = min(x0 + RU_W, dw), y1 = min(y0 + RU_H, dh);
    // source patch: taps are clamped and monotone, so the first tap of the first output and the last tap of the last bound it
    const int sx0 = tx[x0].i[0], sx1 = tx[x1 - 1].i[3], sy0 = ty[y0].i[0], sy1 = ty[y1 - 1].i[3];
    const int ncols = sx1 - sx0 + 1, nrows = sy1 - sy0 + 1; // <= RU_SW, RU_SH for every scale >= 1 (checked by the launcher)
    const int nbytes = ncols * C;
    for (int r = 0; r < nrows; ++r) {
        const uint8_t *g = src + (size_t)(sy0 + r) * spitch + (size_t)sx0 * C;
        for (int b = threadIdx.x; b < nbytes; b += RU_W) s_src[r][b] = __ldg(g + b);
    }
    const int x = x0 + threadIdx.x;
    int4 xi = make_int4(0, 0, 0, 0);
    float4 xw = make_float4(0.f, 0.f, 0.f, 0.f);
    if (x < dw) {
        xi = *reinterpret_cast<const int4 *>(tx[x].i);
        xw = *reinterpret_cast<const float4 *>(tx[x].w);
    }
    const int xo[4] = {(xi.x - sx0) * C, (xi.y - sx0) * C, (xi.z - sx0) * C, (xi.w - sx0) * C};
    const float wx[4] = {xw.x, xw.y, xw.z, xw.w};
    __syncthreads();
    for (int y = y0; y < y1; ++y) {
        const int4 yi = *reinterpret_cast<const int4 *>(ty[y].i); // uniform
        const float4 yw = *reinterpret_cast<const float4 *>(ty[y].w);
        const uint8_t *r0 = s_src[yi.x - sy0], *r1 = s_src[yi.y - sy0], *r2 = s_src[yi.z - sy0], *r3 = s_src[yi.w - sy0];
        for (int b = threadIdx.x; b < nbytes; b += RU_W) { // rows first: the column sum of every source column, once
            float t = __fmul_rn((float)r0[b], yw.x);
            t = __fadd_rn(t, __fmul_rn((float)r1[b], yw.y));
            t = __fadd_rn(t, __fmul_rn((float)r2[b], yw.z));
            t = __fadd_rn(t, __fmul_rn((float)r3[b], yw.w));
            s_t[b] = t;
        }
        __syncthreads();
        if (x < dw) {
            uint8_t *out = dst + (size_t)y * dpitch + (size_t)x * C;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                float o = __fmul_rn(s_t[xo[0] + c], wx[0]);
                o = __fadd_rn(o, __fmul_rn(s_t[xo[1] + c], wx[1]));
                o = __fadd_rn(o, __fmul_rn(s_t[xo[2] + c], wx[2]));
                o = __fadd_rn(o, __fmul_rn(s_t[xo[3] + c], wx[3]));
                out[c] = (uint8_t)min(max(__float2int_rn(o), 0), 255); // rint, then saturate
            }
        }
        __syncthreads();
    }
}

// OpenCV interpolateCubic (a = -0.75) in float64, the operation order of oracle/resize_oracle.py::cubic_weights
void cubic_weights(double x, float *w)
{
    const double A = -0.75;
    const double c0 = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
    const double c1 = ((A + 2) * x - (A + 3)) * x * x + 1;
    const double c2 = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
    const double c3 = 1.0 - c0 - c1 - c2;
    w[0] = (float)c0;
    w[1] = (float)c1;
    w[2] = (float)c2;
    w[3] = (float)c3;
}

void axis_table(int src, int dst, ResizeTap *tab)
{
    const double inv = (double)dst / (double)src; // OpenCV: inv_scale = dsize / ssize, scale = 1 / inv_scale
    const double scale = 1.0 / inv;
    for (int d = 0; d < dst; ++d) {
        const double f = ((double)d + 0.5) * scale - 0.5;
        const double s = floor(f);
        const long long s0 = (long long)s;
        for (int k = 0; k < 4; ++k) {
            long long i = s0 + k - 1;
            tab[d].i[k] = (int)(i < 0 ? 0 : (i > src - 1 ? src - 1 : i));
        }
        cubic_weights(f - s, tab[d].w);
    }
}

int check_resize_args(srf_ctx *ctx, const void *src, size_t spitch, int h, int w, int C, const void *dst, size_t dpitch, int dh, int dw)
{
    SRF_REQUIRE(ctx, src && dst, "resize_cubic: null pointer");
    SRF_REQUIRE(ctx, C == 1 || C == 3, "resize_cubic: channels must be 1 or 3 (got %d)", C);
    SRF_REQUIRE(ctx, h >= 1 && w >= 1, "resize_cubic: empty source image (%dx%d)", h, w);
    SRF_REQUIRE(ctx, dh >= 1 && dw >= 1, "resize_cubic: empty destination size (%dx%d)", dh, dw); // cv2.resize asserts !dsize.empty()
    SRF_REQUIRE(ctx, spitch >= (size_t)w * C && dpitch >= (size_t)dw * C, "resize_cubic: pitch < row bytes");
    return SRF_OK;
}

int resize_device(srf_ctx *ctx, const uint8_t *d_src, size_t spitch, int h, int w, int C, uint8_t *d_dst, size_t dpitch, int dh, int dw,
                  cudaStream_t st)
{
    std::vector<ResizeTap> tab((size_t)dw + dh);
    axis_table(w, dw, tab.data());
    axis_table(h, dh, tab.data() + dw);
    int rc;
    if ((rc = srf_reserve(ctx, ctx->ws_tab, tab.size() * sizeof(ResizeTap)))) return rc;
    // pageable source: the copy is staged before the call returns, so `tab` may go out of scope
    SRF_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->ws_tab.ptr, tab.data(), tab.size() * sizeof(ResizeTap), cudaMemcpyHostToDevice, st));
    const ResizeTap *tx = (const ResizeTap *)ctx->ws_tab.ptr, *ty = tx + dw;
    dim3 grid(srf_ceil_div(dw, RS_BX), srf_ceil_div(dh, RS_BY)), block(RS_BX, RS_BY);
    {
        // every source pixel a tap can reach is read once from HBM (the 16-tap gathers overlap in L1/L2) + the output
        const double src_px = (double)(dw < w ? std::min((double)w, 4.0 * dw) : w) * (double)(dh < h ? std::min((double)h, 4.0 * dh) : h);
        srf_launch_scope scope(ctx, "resize_cubic", C * (src_px + (double)dw * dh), st);
        if (dw >= w && dh >= h) { // enlarging (the SR stand-in): the tiled kernel -- same bytes, a quarter of the instructions
            dim3 ugrid(srf_ceil_div(dw, RU_W), srf_ceil_div(dh, RU_H));
            if (C == 3)
                k_resize_cubic_up<3><<<ugrid, RU_W, 0, st>>>(d_src, spitch, d_dst, dpitch, dh, dw, tx, ty);
            else
                k_resize_cubic_up<1><<<ugrid, RU_W, 0, st>>>(d_src, spitch, d_dst, dpitch, dh, dw, tx, ty);
        } else if (C == 3)
            k_resize_cubic<3><<<grid, block, 0, st>>>(d_src, spitch, d_dst, dpitch, dh, dw, tx, ty);
        else
            k_resize_cubic<1><<<grid, block, 0, st>>>(d_src, spitch, d_dst, dpitch, dh, dw, tx, ty);
    }
    SRF_CHECK_CUDA(ctx, cudaGetLastError());
    return SRF_OK;
}

} // namespace

extern "C" int srf_resize_cubic_taps(int src, int dst, int *idx4, float *w4)
{
    if (src < 1 || dst < 1 || !idx4 || !w4) return SRF_ERR_ARG;
    std::vector<ResizeTap> tab((size_t)dst);
    axis_table(src, dst, tab.data());
    for (int d = 0; d < dst; ++d)
        for (int k = 0; k < 4; ++k) {
            idx4[4 * d + k] = tab[d].i[k];
            w4[4 * d + k] = tab[d].w[k];
        }
    return SRF_OK;
}

extern "C" int srf_resize_cubic_u8(srf_ctx *ctx, const uint8_t *d_src, size_t src_pitch, int h, int w, int channels, uint8_t *d_dst,
                                   size_t dst_pitch, int dst_h, int dst_w, void *stream)
{
    if (!ctx) return SRF_ERR_ARG;
    int rc = check_resize_args(ctx, d_src, src_pitch, h, w, channels, d_dst, dst_pitch, dst_h, dst_w);
    if (rc) return rc;
    SRF_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    rc = resize_device(ctx, d_src, src_pitch, h, w, channels, d_dst, dst_pitch, dst_h, dst_w, (cudaStream_t)stream);
    return rc ? rc : srf_note_device_work(ctx, (cudaStream_t)stream);
}

extern "C" int srf_resize_cubic_u8_host(srf_ctx *ctx, const uint8_t *h_src, size_t src_pitch, int h, int w, int channels, uint8_t *h_dst,
                                        size_t dst_pitch, int dst_h, int dst_w)
{
    if (!ctx) return SRF_ERR_ARG;
    int rc = check_resize_args(ctx, h_src, src_pitch, h, w, channels, h_dst, dst_pitch, dst_h, dst_w);
    if (rc) return rc;
    SRF_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (int hrc = srf_host_enter(ctx)) return hrc;
    const size_t sp = srf_align_up((size_t)w * channels, 256), dp = srf_align_up((size_t)dst_w * channels, 256);
    if ((rc = srf_reserve(ctx, ctx->ws_stage_in, sp * (size_t)h))) return rc;
    if ((rc = srf_reserve(ctx, ctx->ws_stage_out, dp * (size_t)dst_h))) return rc;
    cudaStream_t st = ctx->own_stream;
    uint8_t *ds = (uint8_t *)ctx->ws_stage_in.ptr, *dd = (uint8_t *)ctx->ws_stage_out.ptr;
    SRF_CHECK_CUDA(ctx, cudaMemcpy2DAsync(ds, sp, h_src, src_pitch, (size_t)w * channels, h, cudaMemcpyHostToDevice, st));
    if ((rc = resize_device(ctx, ds, sp, h, w, channels, dd, dp, dst_h, dst_w, st))) return rc;
    SRF_CHECK_CUDA(ctx, cudaMemcpy2DAsync(h_dst, dst_pitch, dd, dp, (size_t)dst_w * channels, dst_h, cudaMemcpyDeviceToHost, st));
    SRF_CHECK_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->last_h2d = (size_t)w * channels * h;
    ctx->last_d2h = (size_t)dst_w * channels * dst_h;
    return SRF_OK;
}
