// srf_seam.cu -- the per-window SSIM of BlendingModule.detect_seams (/root/reference/blending_module.py:765-851) with
// _compute_ssim (853-903): for a tile placed at (tile_x, tile_y) on the fused result, every window x window patch (stride
// `stride`) of the region the two share is converted to gray and scored with the GLOBAL (whole-patch) SSIM
//     ((2 mu1 mu2 + c1)(2 s12 + c2)) / ((mu1^2 + mu2^2 + c1)(s1 + s2 + c2)),     float64,
// tile first, result second.  The reference converts with cv2.COLOR_BGR2GRAY although the arrays are RGB, i.e. channel 0
// gets the blue weight: gray = (ch0*3735 + ch1*19235 + ch2*9798 + 16384) >> 15 -- kept.  All five patch sums are exact
// integers, and so are NumPy's float64 mean / var / covariance of a patch of bytes (every intermediate is a small dyadic
// rational), so mu = S/n, var = (n*Sxx - S^2)/n^2, cov = (n*Sxy - Sx*Sy)/n^2 reproduce them bit for bit; the final
// expression follows the reference's operation order with unfused float64 instructions.
// One warp per window: lanes stride over the patch, five integer warp reductions, lane 0 writes the score.
#include "srf_common.cuh"

namespace {

struct SeamParams {
    const uint8_t *tile, *result; // both already offset to the top-left pixel of the shared region
    size_t tpitch, rpitch;
    int roi_h, roi_w;
    int window, stride, ny, nx;
    double c1, c2;
    double *out; // ny x nx
};

template <int C>
__device__ __forceinline__ int seam_gray(const uint8_t *p)
{
    if (C == 1) return p[0];
    return ((int)p[0] * 3735 + (int)p[1] * 19235 + (int)p[2] * 9798 + 16384) >> 15; // cv2 BGR2GRAY on the stored order
}

template <int C>
__global__ void __launch_bounds__(256) k_window_ssim(SeamParams P)
{
    pdl_enter();
    const int lane = threadIdx.x & 31;
    const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wid >= (long long)P.ny * P.nx) return;
    const int wy = (int)(wid / P.nx), wx = (int)(wid - (long long)wy * P.nx);
    const uint8_t *t0 = P.tile + (size_t)(wy * P.stride) * P.tpitch + (size_t)(wx * P.stride) * C;
    const uint8_t *r0 = P.result + (size_t)(wy * P.stride) * P.rpitch + (size_t)(wx * P.stride) * C;
    const int n = P.window * P.window;
    long long s1 = 0, s2 = 0, s11 = 0, s22 = 0, s12 = 0;
    for (int i = lane; i < n; i += 32) {
        const int y = i / P.window, x = i - y * P.window;
        const int a = seam_gray<C>(t0 + (size_t)y * P.tpitch + (size_t)x * C);
        const int b = seam_gray<C>(r0 + (size_t)y * P.rpitch + (size_t)x * C);
        s1 += a;
        s2 += b;
        s11 += a * a;
        s22 += b * b;
        s12 += a * b;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_down_sync(0xffffffffu, s1, o);
        s2 += __shfl_down_sync(0xffffffffu, s2, o);
        s11 += __shfl_down_sync(0xffffffffu, s11, o);
        s22 += __shfl_down_sync(0xffffffffu, s22, o);
        s12 += __shfl_down_sync(0xffffffffu, s12, o);
    }
    if (lane != 0) return;
    const double dn = (double)n, dn2 = __dmul_rn(dn, dn);
    const double mu1 = __ddiv_rn((double)s1, dn), mu2 = __ddiv_rn((double)s2, dn);
    // exact integers below 2^53; the quotients equal NumPy's np.var / covariance of the patch whenever n is a power of two
    // (16 x 16 by default) and agree to the last bits otherwise
    const double v1 = __ddiv_rn((double)(n * s11 - s1 * s1), dn2);
    const double v2 = __ddiv_rn((double)(n * s22 - s2 * s2), dn2);
    const double cv = __ddiv_rn((double)(n * s12 - s1 * s2), dn2);
    const double num = __dmul_rn(__dadd_rn(__dmul_rn(__dmul_rn(2.0, mu1), mu2), P.c1), __dadd_rn(__dmul_rn(2.0, cv), P.c2));
    const double den = __dmul_rn(__dadd_rn(__dadd_rn(__dmul_rn(mu1, mu1), __dmul_rn(mu2, mu2)), P.c1), __dadd_rn(__dadd_rn(v1, v2), P.c2));
    P.out[wid] = __ddiv_rn(num, den);
}

int seam_geometry(srf_ctx *ctx, int rh, int rw, int th, int tw, int C, int ty, int tx, int window, int stride, int *roi_h, int *roi_w, int *ny,
                  int *nx)
{
    SRF_REQUIRE(ctx, C == 1 || C == 3, "window_ssim: channels must be 1 or 3 (got %d)", C);
    SRF_REQUIRE(ctx, rh >= 1 && rw >= 1 && th >= 1 && tw >= 1, "window_ssim: empty image");
    // n * Sxx must stay below 2^53 for the integer form to be exact in float64: window <= 128 gives 2^14 * 2^30
    SRF_REQUIRE(ctx, window >= 1 && window <= 128 && stride >= 1, "window_ssim: bad window %d (1..128) / stride %d", window, stride);
    SRF_REQUIRE(ctx, tx >= 0 && ty >= 0, "window_ssim: negative tile position (%d,%d)", ty, tx);
    *roi_h = std::max(0, std::min(ty + th, rh) - ty); // blending_module.py:809-814: the part of the tile inside the result
    *roi_w = std::max(0, std::min(tx + tw, rw) - tx);
    *ny = *roi_h >= window ? (*roi_h - window) / stride + 1 : 0; // range(0, roi_h - window + 1, stride)
    *nx = *roi_w >= window ? (*roi_w - window) / stride + 1 : 0;
    return SRF_OK;
}

int seam_launch(srf_ctx *ctx, const uint8_t *d_result, size_t rpitch, const uint8_t *d_tile, size_t tpitch, int C, int roi_h, int roi_w, int window,
                int stride, int ny, int nx, double c1, double c2, double *d_out, cudaStream_t st)
{
    if (ny == 0 || nx == 0) return SRF_OK;
    SeamParams P;
    P.tile = d_tile;
    P.result = d_result;
    P.tpitch = tpitch;
    P.rpitch = rpitch;
    P.roi_h = roi_h;
    P.roi_w = roi_w;
    P.window = window;
    P.stride = stride;
    P.ny = ny;
    P.nx = nx;
    P.c1 = c1;
    P.c2 = c2;
    P.out = d_out;
    const long long windows = (long long)ny * nx;
    dim3 grid((unsigned)((windows + 7) / 8)), block(256);
    srf_launch_scope scope(ctx, "window_ssim", 2.0 * C * (double)roi_h * roi_w + 8.0 * (double)windows, st);
    if (C == 3)
        srf_launch_pdl(k_window_ssim<3>, grid, block, 0, st, P);
    else
        srf_launch_pdl(k_window_ssim<1>, grid, block, 0, st, P);
    SRF_CHECK_CUDA(ctx, cudaGetLastError());
    return SRF_OK;
}

} // namespace

extern "C" int srf_window_ssim_shape(int result_h, int result_w, int tile_h, int tile_w, int tile_y, int tile_x, int window, int stride, int *ny,
                                     int *nx)
{
    if (!ny || !nx || window < 1 || stride < 1 || tile_x < 0 || tile_y < 0) return SRF_ERR_ARG;
    const int roi_h = std::max(0, std::min(tile_y + tile_h, result_h) - tile_y), roi_w = std::max(0, std::min(tile_x + tile_w, result_w) - tile_x);
    *ny = roi_h >= window ? (roi_h - window) / stride + 1 : 0;
    *nx = roi_w >= window ? (roi_w - window) / stride + 1 : 0;
    return SRF_OK;
}

extern "C" int srf_window_ssim(srf_ctx *ctx, const uint8_t *d_result, size_t result_pitch, int result_h, int result_w, const uint8_t *d_tile,
                               size_t tile_pitch, int tile_h, int tile_w, int channels, int tile_y, int tile_x, int window, int stride, double c1,
                               double c2, double *d_out, void *stream)
{
    if (!ctx) return SRF_ERR_ARG;
    SRF_REQUIRE(ctx, d_result && d_tile && d_out, "window_ssim: null pointer");
    int roi_h, roi_w, ny, nx;
    int rc = seam_geometry(ctx, result_h, result_w, tile_h, tile_w, channels, tile_y, tile_x, window, stride, &roi_h, &roi_w, &ny, &nx);
    if (rc) return rc;
    SRF_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    return seam_launch(ctx, d_result + (size_t)tile_y * result_pitch + (size_t)tile_x * channels, result_pitch, d_tile, tile_pitch, channels, roi_h,
                       roi_w, window, stride, ny, nx, c1, c2, d_out, (cudaStream_t)stream);
}

extern "C" int srf_window_ssim_host(srf_ctx *ctx, const uint8_t *h_result, size_t result_pitch, int result_h, int result_w, const uint8_t *h_tile,
                                    size_t tile_pitch, int tile_h, int tile_w, int channels, int tile_y, int tile_x, int window, int stride,
                                    double c1, double c2, double *h_out)
{
    if (!ctx) return SRF_ERR_ARG;
    SRF_REQUIRE(ctx, h_result && h_tile && h_out, "window_ssim: null pointer");
    int roi_h, roi_w, ny, nx;
    int rc = seam_geometry(ctx, result_h, result_w, tile_h, tile_w, channels, tile_y, tile_x, window, stride, &roi_h, &roi_w, &ny, &nx);
    if (rc) return rc;
    if (ny == 0 || nx == 0) return SRF_OK;
    SRF_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (int hrc = srf_host_enter(ctx)) return hrc;
    // only the shared region travels: both images share one staging buffer, the scores the other
    const size_t dp = srf_align_up((size_t)roi_w * channels, 256), one = dp * (size_t)roi_h, ob = sizeof(double) * (size_t)ny * nx;
    if ((rc = srf_reserve(ctx, ctx->ws_stage_in, 2 * one))) return rc;
    if ((rc = srf_reserve(ctx, ctx->ws_stage_out, ob))) return rc;
    cudaStream_t st = ctx->own_stream;
    uint8_t *dr = (uint8_t *)ctx->ws_stage_in.ptr, *dt = dr + one;
    SRF_CHECK_CUDA(ctx, cudaMemcpy2DAsync(dr, dp, h_result + (size_t)tile_y * result_pitch + (size_t)tile_x * channels, result_pitch,
                                          (size_t)roi_w * channels, roi_h, cudaMemcpyHostToDevice, st));
    SRF_CHECK_CUDA(ctx, cudaMemcpy2DAsync(dt, dp, h_tile, tile_pitch, (size_t)roi_w * channels, roi_h, cudaMemcpyHostToDevice, st));
    if ((rc = seam_launch(ctx, dr, dp, dt, dp, channels, roi_h, roi_w, window, stride, ny, nx, c1, c2, (double *)ctx->ws_stage_out.ptr, st))) return rc;
    SRF_CHECK_CUDA(ctx, cudaMemcpyAsync(h_out, ctx->ws_stage_out.ptr, ob, cudaMemcpyDeviceToHost, st));
    SRF_CHECK_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->last_h2d = 2 * (size_t)roi_w * channels * roi_h;
    ctx->last_d2h = ob;
    return SRF_OK;
}
