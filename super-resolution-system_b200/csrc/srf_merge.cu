// srf_merge.cu -- TilingModule.merge_tiles (/root/reference/tiling_module.py:1074-1135): feathered weighted average of the
// processed tiles on the output canvas,
//     out = astype(uint8)( sum_t tile_t * w_t / max(sum_t w_t, 1e-6) ),   float32, tiles in list order,
// with the weight maps of _create_blend_weight (1137-1175): float32 ones multiplied in place by float64 linspace ramps
// over the (scaled) overlap rows and columns.  Such a map is w(y, x) = f32(f32(wy[y] * wl[x]) * wr[x]) with the products
// taken in float64 (NumPy's in-place float32 *= float64), so the caller passes the three 1-D tables per tile and the
// kernel reproduces every rounding; the full map is never built.
// A tile whose stored size differs from its canvas size is first squashed with cv2.resize's default INTER_LINEAR, as
// the reference does for the padded edge tiles (1107-1109): OpenCV's own 11-bit fixed-point bilinear code for uint8
// (IPP is not used for 8-bit linear), restated in oracle/resize_oracle.py::resize_linear_u8 and bit-exact against
// cv2 4.13 in both directions (the vertical pass keeps the fraction above the first / below the last source row and clamps
// its two row indices separately -- only enlarging reaches those rows).
#include "srf_common.cuh"

#include <math.h>

#include <algorithm>

namespace {

struct MergeTileDev {
    const uint8_t *data; // (h, w, C) at its canvas size
    size_t pitch;
    int h, w, y, x;
    const float *wy;  // h row factors, or nullptr: weight 1 everywhere (blending=False)
    const double *wl; // w left-ramp factors (1 outside the ramp)
    const double *wr; // w right-ramp factors
};

struct LinTap {
    int i0, i1, a0, a1; // source index pair and the two 11-bit coefficients
};

constexpr int MG_BX = 32, MG_BY = 8;

template <int C>
__global__ void __launch_bounds__(MG_BX *MG_BY) k_resize_linear_u8(const uint8_t *__restrict__ src, size_t spitch, uint8_t *__restrict__ dst,
                                                                    size_t dpitch, int dh, int dw, const LinTap *__restrict__ tx,
                                                                    const LinTap *__restrict__ ty)
{
    pdl_enter();
    const int x = blockIdx.x * MG_BX + threadIdx.x, y = blockIdx.y * MG_BY + threadIdx.y;
    if (x >= dw || y >= dh) return;
    const LinTap ax = tx[x], ay = ty[y];
    const uint8_t *r0 = src + (size_t)ay.i0 * spitch, *r1 = src + (size_t)ay.i1 * spitch;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int h0 = (int)__ldg(r0 + ax.i0 * C + c) * ax.a0 + (int)__ldg(r0 + ax.i1 * C + c) * ax.a1; // horizontal pass, int32
        const int h1 = (int)__ldg(r1 + ax.i0 * C + c) * ax.a0 + (int)__ldg(r1 + ax.i1 * C + c) * ax.a1;
        const int v = (((ay.a0 * (h0 >> 4)) >> 16) + ((ay.a1 * (h1 >> 4)) >> 16) + 2) >> 2;         // OpenCV's VResizeLinear for uchar
        dst[(size_t)y * dpitch + (size_t)x * C + c] = (uint8_t)min(max(v, 0), 255);
    }
}

template <int C>
__global__ void __launch_bounds__(MG_BX *MG_BY) k_merge_tiles(const MergeTileDev *__restrict__ tiles, int n_tiles, uint8_t *__restrict__ canvas,
                                                               size_t pitch, int out_h, int out_w)
{
    pdl_enter();
    const int px = blockIdx.x * MG_BX + threadIdx.x, py = blockIdx.y * MG_BY + threadIdx.y;
    if (px >= out_w || py >= out_h) return;
    float acc[C], wsum = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    for (int i = 0; i < n_tiles; ++i) { // list order = the reference's accumulation order
        const MergeTileDev &t = tiles[i];
        const int tx = px - t.x, ty = py - t.y;
        if (tx < 0 || ty < 0 || tx >= t.w || ty >= t.h) continue;
        float w = 1.f;
        if (t.wy != nullptr) {
            w = (float)((double)__ldg(t.wy + ty) * __ldg(t.wl + tx)); // float32 *= float64: product in double, stored as float
            w = (float)((double)w * __ldg(t.wr + tx));
        }
        const uint8_t *p = t.data + (size_t)ty * t.pitch + (size_t)tx * C;
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = __fadd_rn(acc[c], __fmul_rn((float)__ldg(p + c), w));
        wsum = __fadd_rn(wsum, w);
    }
    const float den = fmaxf(wsum, 1e-6f);
    uint8_t *o = canvas + (size_t)py * pitch + (size_t)px * C;
#pragma unroll
    for (int c = 0; c < C; ++c) o[c] = (uint8_t)min(max(__float2int_rz(__fdiv_rn(acc[c], den)), 0), 255); // astype(uint8): truncation
}

// ---- BlendingModule.weighted_average_fusion with CUSTOM weight maps (/root/reference/blending_module.py:661-760, the
// `weights=` argument): per tile an (h, w) plane of float32 or float64 weights.  NumPy's arithmetic for the two cases:
//   float32 map: acc += f32(tile) * w          -> float32 product, float32 sum
//   float64 map: acc += f32(tile) * w          -> float64 product, the in-place add runs in float64 and is stored as float32
// and the same for the weight accumulator; then acc / max(wsum, 1e-6), clip(0, 255), astype(uint8).  (Without custom maps
// the call is the gather kernel with one pyramid level -- srf_laplacian_fuse(levels = 1).)
struct WavgTileDev {
    const uint8_t *data;
    size_t pitch;
    int h, w, y, x; // extent already cropped to the canvas
    const void *wgt;
    size_t wpitch; // bytes
    int f64;
};

template <int C>
__global__ void __launch_bounds__(MG_BX *MG_BY) k_weighted_average(const WavgTileDev *__restrict__ tiles, int n_tiles, uint8_t *__restrict__ canvas,
                                                                    size_t pitch, int out_h, int out_w)
{
    pdl_enter();
    const int px = blockIdx.x * MG_BX + threadIdx.x, py = blockIdx.y * MG_BY + threadIdx.y;
    if (px >= out_w || py >= out_h) return;
    float acc[C], wsum = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    for (int i = 0; i < n_tiles; ++i) { // list order = the reference's accumulation order
        const WavgTileDev &t = tiles[i];
        const int tx = px - t.x, ty = py - t.y;
        if (tx < 0 || ty < 0 || tx >= t.w || ty >= t.h) continue;
        const uint8_t *p = t.data + (size_t)ty * t.pitch + (size_t)tx * C;
        const char *wrow = (const char *)t.wgt + (size_t)ty * t.wpitch;
        if (t.f64) {
            const double w = __ldg((const double *)wrow + tx);
#pragma unroll
            for (int c = 0; c < C; ++c) acc[c] = (float)__dadd_rn((double)acc[c], __dmul_rn((double)__ldg(p + c), w));
            wsum = (float)__dadd_rn((double)wsum, w);
        } else {
            const float w = __ldg((const float *)wrow + tx);
#pragma unroll
            for (int c = 0; c < C; ++c) acc[c] = __fadd_rn(acc[c], __fmul_rn((float)__ldg(p + c), w));
            wsum = __fadd_rn(wsum, w);
        }
    }
    const float den = fmaxf(wsum, 1e-6f);
    uint8_t *o = canvas + (size_t)py * pitch + (size_t)px * C;
#pragma unroll
    for (int c = 0; c < C; ++c) o[c] = (uint8_t)__float2int_rz(fminf(fmaxf(__fdiv_rn(acc[c], den), 0.f), 255.f)); // clip, then astype(uint8)
}

// OpenCV's INTER_LINEAR taps for uint8 (oracle/resize_oracle.py::linear_table): float32 sample position, coefficients =
// cvRound(frac * 2048) as shorts.  Horizontal table: fraction reset to 0 at the clamped ends.  Vertical pass: fraction
// kept, the two row indices clamped one by one (both taps then read the same row, each product truncated on its own).
void linear_table(int src, int dst, LinTap *tab, bool vertical)
{
    const double inv = (double)dst / (double)src;
    const double scale = 1.0 / inv;
    for (int d = 0; d < dst; ++d) {
        float f = (float)(((double)d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        float fr = f - (float)s;
        if (!vertical) {
            if (s < 0) {
                fr = 0.f;
                s = 0;
            }
            if (s >= src - 1) {
                fr = 0.f;
                s = src - 1;
            }
        }
        tab[d].i0 = std::min(std::max(s, 0), src - 1);
        tab[d].i1 = std::min(std::max(s + 1, 0), src - 1);
        tab[d].a1 = (int)lrintf(fr * 2048.f);
        tab[d].a0 = (int)lrintf((1.f - fr) * 2048.f);
    }
}

int resize_linear_device(srf_ctx *ctx, const uint8_t *d_src, size_t spitch, int h, int w, int C, uint8_t *d_dst, size_t dpitch, int dh, int dw,
                         LinTap *d_tab, cudaStream_t st)
{
    std::vector<LinTap> tab((size_t)dw + dh);
    linear_table(w, dw, tab.data(), false);
    linear_table(h, dh, tab.data() + dw, true);
    SRF_CHECK_CUDA(ctx, cudaMemcpyAsync(d_tab, tab.data(), tab.size() * sizeof(LinTap), cudaMemcpyHostToDevice, st));
    dim3 grid(srf_ceil_div(dw, MG_BX), srf_ceil_div(dh, MG_BY)), block(MG_BX, MG_BY);
    srf_launch_scope scope(ctx, "resize_linear", (double)C * ((double)h * w + (double)dh * dw), st);
    if (C == 3)
        srf_launch_pdl(k_resize_linear_u8<3>, grid, block, 0, st, d_src, spitch, d_dst, dpitch, dh, dw, (const LinTap *)d_tab, (const LinTap *)(d_tab + dw));
    else
        srf_launch_pdl(k_resize_linear_u8<1>, grid, block, 0, st, d_src, spitch, d_dst, dpitch, dh, dw, (const LinTap *)d_tab, (const LinTap *)(d_tab + dw));
    SRF_CHECK_CUDA(ctx, cudaGetLastError());
    return SRF_OK;
}

int check_merge_args(srf_ctx *ctx, const srf_merge_tile *tiles, int n, int C, int out_h, int out_w, const void *canvas, size_t pitch)
{
    SRF_REQUIRE(ctx, tiles && n >= 1, "merge_tiles: no tiles");
    SRF_REQUIRE(ctx, C == 1 || C == 3, "merge_tiles: channels must be 1 or 3 (got %d)", C);
    SRF_REQUIRE(ctx, out_h >= 1 && out_w >= 1 && canvas, "merge_tiles: empty canvas");
    SRF_REQUIRE(ctx, pitch >= (size_t)out_w * C, "merge_tiles: canvas pitch < row bytes");
    for (int i = 0; i < n; ++i) {
        const srf_merge_tile &t = tiles[i];
        SRF_REQUIRE(ctx, t.data && t.h >= 1 && t.w >= 1 && t.out_h >= 1 && t.out_w >= 1, "merge_tiles: tile %d is empty", i);
        SRF_REQUIRE(ctx, t.pitch >= (size_t)t.w * C, "merge_tiles: tile %d pitch < row bytes", i);
        SRF_REQUIRE(ctx, t.y >= 0 && t.x >= 0, "merge_tiles: tile %d has a negative position (%d,%d)", i, t.y, t.x);
        SRF_REQUIRE(ctx, (t.wy == nullptr) == (t.wl == nullptr) && (t.wl == nullptr) == (t.wr == nullptr),
                    "merge_tiles: tile %d needs all three weight tables or none", i);
    }
    return SRF_OK;
}

// tiles[i].data are DEVICE pointers here; the weight tables are host pointers
int merge_device(srf_ctx *ctx, const srf_merge_tile *tiles, int n, int C, int out_h, int out_w, uint8_t *d_canvas, size_t pitch, cudaStream_t st)
{
    // workspace layout: [MergeTileDev x n | per tile: wy (float) wl wr (double) | LinTap tables | resized tiles]
    size_t off = srf_align_up(sizeof(MergeTileDev) * (size_t)n, 256);
    std::vector<size_t> o_w(n, 0), o_tab(n, 0), o_img(n, 0), rp(n, 0);
    for (int i = 0; i < n; ++i)
        if (tiles[i].wy) {
            o_w[i] = off;
            off += srf_align_up((size_t)tiles[i].out_h * sizeof(float), 16) + 2 * srf_align_up((size_t)tiles[i].out_w * sizeof(double), 16);
            off = srf_align_up(off, 256);
        }
    for (int i = 0; i < n; ++i)
        if (tiles[i].out_h != tiles[i].h || tiles[i].out_w != tiles[i].w) {
            o_tab[i] = off;
            off = srf_align_up(off + sizeof(LinTap) * ((size_t)tiles[i].out_w + tiles[i].out_h), 256);
            rp[i] = srf_align_up((size_t)tiles[i].out_w * C, 16);
            o_img[i] = off;
            off = srf_align_up(off + rp[i] * (size_t)tiles[i].out_h, 256);
        }
    int rc;
    if ((rc = srf_reserve(ctx, ctx->ws_merge, off))) return rc;
    uint8_t *base = (uint8_t *)ctx->ws_merge.ptr;
    std::vector<MergeTileDev> dev(n);
    for (int i = 0; i < n; ++i) {
        const srf_merge_tile &t = tiles[i];
        MergeTileDev &d = dev[i];
        d.h = t.out_h;
        d.w = t.out_w;
        d.y = t.y;
        d.x = t.x;
        d.wy = nullptr;
        d.wl = d.wr = nullptr;
        if (t.wy) {
            float *dwy = (float *)(base + o_w[i]);
            double *dwl = (double *)(base + o_w[i] + srf_align_up((size_t)t.out_h * sizeof(float), 16));
            double *dwr = (double *)((uint8_t *)dwl + srf_align_up((size_t)t.out_w * sizeof(double), 16));
            SRF_CHECK_CUDA(ctx, cudaMemcpyAsync(dwy, t.wy, sizeof(float) * t.out_h, cudaMemcpyHostToDevice, st));
            SRF_CHECK_CUDA(ctx, cudaMemcpyAsync(dwl, t.wl, sizeof(double) * t.out_w, cudaMemcpyHostToDevice, st));
            SRF_CHECK_CUDA(ctx, cudaMemcpyAsync(dwr, t.wr, sizeof(double) * t.out_w, cudaMemcpyHostToDevice, st));
            d.wy = dwy;
            d.wl = dwl;
            d.wr = dwr;
        }
        if (o_img[i]) { // squash to the canvas size first (cv2.resize, INTER_LINEAR)
            uint8_t *img = base + o_img[i];
            if ((rc = resize_linear_device(ctx, t.data, t.pitch, t.h, t.w, C, img, rp[i], t.out_h, t.out_w, (LinTap *)(base + o_tab[i]), st))) return rc;
            d.data = img;
            d.pitch = rp[i];
        } else {
            d.data = t.data;
            d.pitch = t.pitch;
        }
    }
    SRF_CHECK_CUDA(ctx, cudaMemcpyAsync(base, dev.data(), sizeof(MergeTileDev) * n, cudaMemcpyHostToDevice, st));
    dim3 grid(srf_ceil_div(out_w, MG_BX), srf_ceil_div(out_h, MG_BY)), block(MG_BX, MG_BY);
    double cov = 0;
    for (int i = 0; i < n; ++i) cov += (double)tiles[i].out_h * tiles[i].out_w;
    srf_launch_scope scope(ctx, "merge_tiles", C * (cov + (double)out_h * out_w), st);
    if (C == 3)
        srf_launch_pdl(k_merge_tiles<3>, grid, block, 0, st, (const MergeTileDev *)base, n, d_canvas, pitch, out_h, out_w);
    else
        srf_launch_pdl(k_merge_tiles<1>, grid, block, 0, st, (const MergeTileDev *)base, n, d_canvas, pitch, out_h, out_w);
    SRF_CHECK_CUDA(ctx, cudaGetLastError());
    return SRF_OK;
}

int check_wavg_args(srf_ctx *ctx, const srf_wavg_tile *tiles, int n, int C, int out_h, int out_w, const void *canvas, size_t pitch)
{
    SRF_REQUIRE(ctx, tiles && n >= 1, "weighted_average: no tiles");
    SRF_REQUIRE(ctx, C == 1 || C == 3, "weighted_average: channels must be 1 or 3 (got %d)", C);
    SRF_REQUIRE(ctx, out_h >= 1 && out_w >= 1 && canvas, "weighted_average: empty canvas");
    SRF_REQUIRE(ctx, pitch >= (size_t)out_w * C, "weighted_average: canvas pitch < row bytes");
    for (int i = 0; i < n; ++i) {
        const srf_wavg_tile &t = tiles[i];
        SRF_REQUIRE(ctx, t.data && t.weight && t.h >= 1 && t.w >= 1, "weighted_average: tile %d is empty", i);
        SRF_REQUIRE(ctx, t.pitch >= (size_t)t.w * C, "weighted_average: tile %d pitch < row bytes", i);
        SRF_REQUIRE(ctx, t.y >= 0 && t.x >= 0, "weighted_average: tile %d has a negative position (%d,%d)", i, t.y, t.x);
        const int ch = std::max(std::min(t.y + t.h, out_h) - t.y, 0), cw = std::max(std::min(t.x + t.w, out_w) - t.x, 0);
        // the reference multiplies tile[:ch, :cw] by weight[:ch, :cw]: a smaller map does not broadcast (ValueError there)
        SRF_REQUIRE(ctx, t.weight_h >= ch && t.weight_w >= cw, "weighted_average: weight map %d is %dx%d, the tile's part of the canvas %dx%d", i,
                    t.weight_w, t.weight_h, cw, ch);
        SRF_REQUIRE(ctx, t.weight_pitch >= (size_t)t.weight_w * (t.weight_f64 ? 8 : 4), "weighted_average: weight map %d pitch < row bytes", i);
    }
    return SRF_OK;
}

// tiles[i].data / .weight are DEVICE pointers here
int wavg_device(srf_ctx *ctx, const srf_wavg_tile *tiles, int n, int C, int out_h, int out_w, uint8_t *d_canvas, size_t pitch, cudaStream_t st)
{
    int rc;
    if ((rc = srf_reserve(ctx, ctx->ws_merge, sizeof(WavgTileDev) * (size_t)n))) return rc;
    std::vector<WavgTileDev> dev(n);
    double cov = 0;
    for (int i = 0; i < n; ++i) {
        const srf_wavg_tile &t = tiles[i];
        WavgTileDev &d = dev[i];
        d.data = t.data;
        d.pitch = t.pitch;
        d.y = t.y;
        d.x = t.x;
        d.h = std::max(std::min(t.y + t.h, out_h) - t.y, 0);
        d.w = std::max(std::min(t.x + t.w, out_w) - t.x, 0);
        d.wgt = t.weight;
        d.wpitch = t.weight_pitch;
        d.f64 = t.weight_f64;
        cov += (double)d.h * d.w * (C + (t.weight_f64 ? 8 : 4));
    }
    SRF_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->ws_merge.ptr, dev.data(), sizeof(WavgTileDev) * n, cudaMemcpyHostToDevice, st));
    dim3 grid(srf_ceil_div(out_w, MG_BX), srf_ceil_div(out_h, MG_BY)), block(MG_BX, MG_BY);
    srf_launch_scope scope(ctx, "weighted_average", cov + (double)C * out_h * out_w, st);
    if (C == 3)
        srf_launch_pdl(k_weighted_average<3>, grid, block, 0, st, (const WavgTileDev *)ctx->ws_merge.ptr, n, d_canvas, pitch, out_h, out_w);
    else
        srf_launch_pdl(k_weighted_average<1>, grid, block, 0, st, (const WavgTileDev *)ctx->ws_merge.ptr, n, d_canvas, pitch, out_h, out_w);
    SRF_CHECK_CUDA(ctx, cudaGetLastError());
    return SRF_OK;
}

} // namespace

extern "C" int srf_weighted_average(srf_ctx *ctx, const srf_wavg_tile *tiles, int n_tiles, int channels, int out_h, int out_w, uint8_t *d_canvas,
                                    size_t canvas_pitch, void *stream)
{
    if (!ctx) return SRF_ERR_ARG;
    int rc = check_wavg_args(ctx, tiles, n_tiles, channels, out_h, out_w, d_canvas, canvas_pitch);
    if (rc) return rc;
    SRF_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    rc = wavg_device(ctx, tiles, n_tiles, channels, out_h, out_w, d_canvas, canvas_pitch, (cudaStream_t)stream);
    return rc ? rc : srf_note_device_work(ctx, (cudaStream_t)stream);
}

extern "C" int srf_weighted_average_host(srf_ctx *ctx, const srf_wavg_tile *tiles, int n_tiles, int channels, int out_h, int out_w,
                                         uint8_t *h_canvas, size_t canvas_pitch)
{
    if (!ctx) return SRF_ERR_ARG;
    int rc = check_wavg_args(ctx, tiles, n_tiles, channels, out_h, out_w, h_canvas, canvas_pitch);
    if (rc) return rc;
    SRF_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (int hrc = srf_host_enter(ctx)) return hrc;
    std::vector<srf_wavg_tile> dt(tiles, tiles + n_tiles);
    std::vector<size_t> off(n_tiles), woff(n_tiles);
    size_t total = 0, h2d = 0;
    for (int i = 0; i < n_tiles; ++i) { // only the part of a tile (and of its map) that lies on the canvas travels
        const int ch = std::max(std::min(tiles[i].y + tiles[i].h, out_h) - tiles[i].y, 0), cw = std::max(std::min(tiles[i].x + tiles[i].w, out_w) - tiles[i].x, 0);
        const bool on_canvas = ch > 0 && cw > 0; // a tile that lies entirely off the canvas contributes nothing and is not staged
        dt[i].h = on_canvas ? ch : 0;
        dt[i].w = on_canvas ? cw : 0;
        off[i] = total;
        dt[i].pitch = srf_align_up((size_t)dt[i].w * channels, 16);
        total += srf_align_up(dt[i].pitch * dt[i].h, 256);
        woff[i] = total;
        dt[i].weight_pitch = srf_align_up((size_t)dt[i].w * (tiles[i].weight_f64 ? 8 : 4), 16);
        dt[i].weight_h = dt[i].h;
        dt[i].weight_w = dt[i].w;
        total += srf_align_up(dt[i].weight_pitch * dt[i].h, 256);
    }
    const size_t cpitch = srf_align_up((size_t)out_w * channels, 256);
    if ((rc = srf_reserve(ctx, ctx->ws_stage_in, std::max<size_t>(total, 256)))) return rc;
    if ((rc = srf_reserve(ctx, ctx->ws_stage_out, cpitch * (size_t)out_h))) return rc;
    cudaStream_t st = ctx->own_stream;
    uint8_t *din = (uint8_t *)ctx->ws_stage_in.ptr, *dcanvas = (uint8_t *)ctx->ws_stage_out.ptr;
    for (int i = 0; i < n_tiles; ++i) {
        const size_t esz = tiles[i].weight_f64 ? 8 : 4;
        dt[i].data = din + off[i];
        dt[i].weight = din + woff[i];
        if (dt[i].h == 0) continue;
        SRF_CHECK_CUDA(ctx, cudaMemcpy2DAsync(din + off[i], dt[i].pitch, tiles[i].data, tiles[i].pitch, (size_t)dt[i].w * channels, dt[i].h,
                                              cudaMemcpyHostToDevice, st));
        SRF_CHECK_CUDA(ctx, cudaMemcpy2DAsync(din + woff[i], dt[i].weight_pitch, tiles[i].weight, tiles[i].weight_pitch, (size_t)dt[i].w * esz, dt[i].h,
                                              cudaMemcpyHostToDevice, st));
        h2d += (size_t)dt[i].w * dt[i].h * (channels + esz);
    }
    if ((rc = wavg_device(ctx, dt.data(), n_tiles, channels, out_h, out_w, dcanvas, cpitch, st))) return rc;
    SRF_CHECK_CUDA(ctx, cudaMemcpy2DAsync(h_canvas, canvas_pitch, dcanvas, cpitch, (size_t)out_w * channels, out_h, cudaMemcpyDeviceToHost, st));
    SRF_CHECK_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->last_h2d = h2d;
    ctx->last_d2h = (size_t)out_w * channels * out_h;
    return SRF_OK;
}

static int linear_taps_out(int src, int dst, bool vertical, int *idx2, int *coef2)
{
    if (src < 1 || dst < 1 || !idx2 || !coef2) return SRF_ERR_ARG;
    std::vector<LinTap> tab((size_t)dst);
    linear_table(src, dst, tab.data(), vertical);
    for (int d = 0; d < dst; ++d) {
        idx2[2 * d] = tab[d].i0;
        idx2[2 * d + 1] = tab[d].i1;
        coef2[2 * d] = tab[d].a0;
        coef2[2 * d + 1] = tab[d].a1;
    }
    return SRF_OK;
}

extern "C" int srf_resize_linear_taps(int src, int dst, int *idx2, int *coef2) { return linear_taps_out(src, dst, false, idx2, coef2); }

extern "C" int srf_resize_linear_taps_axis(int src, int dst, int vertical, int *idx2, int *coef2)
{
    return linear_taps_out(src, dst, vertical != 0, idx2, coef2);
}

extern "C" int srf_merge_tiles(srf_ctx *ctx, const srf_merge_tile *tiles, int n_tiles, int channels, int out_h, int out_w, uint8_t *d_canvas,
                               size_t canvas_pitch, void *stream)
{
    if (!ctx) return SRF_ERR_ARG;
    int rc = check_merge_args(ctx, tiles, n_tiles, channels, out_h, out_w, d_canvas, canvas_pitch);
    if (rc) return rc;
    SRF_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    rc = merge_device(ctx, tiles, n_tiles, channels, out_h, out_w, d_canvas, canvas_pitch, (cudaStream_t)stream);
    return rc ? rc : srf_note_device_work(ctx, (cudaStream_t)stream);
}

extern "C" int srf_merge_tiles_host(srf_ctx *ctx, const srf_merge_tile *tiles, int n_tiles, int channels, int out_h, int out_w, uint8_t *h_canvas,
                                    size_t canvas_pitch)
{
    if (!ctx) return SRF_ERR_ARG;
    int rc = check_merge_args(ctx, tiles, n_tiles, channels, out_h, out_w, h_canvas, canvas_pitch);
    if (rc) return rc;
    SRF_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (int hrc = srf_host_enter(ctx)) return hrc;
    std::vector<srf_merge_tile> dt(tiles, tiles + n_tiles);
    std::vector<size_t> off(n_tiles);
    size_t total = 0, h2d = 0;
    for (int i = 0; i < n_tiles; ++i) {
        off[i] = total;
        dt[i].pitch = srf_align_up((size_t)tiles[i].w * channels, 16);
        total += srf_align_up(dt[i].pitch * tiles[i].h, 256);
    }
    const size_t cpitch = srf_align_up((size_t)out_w * channels, 256);
    if ((rc = srf_reserve(ctx, ctx->ws_stage_in, total))) return rc;
    if ((rc = srf_reserve(ctx, ctx->ws_stage_out, cpitch * (size_t)out_h))) return rc;
    cudaStream_t st = ctx->own_stream;
    uint8_t *din = (uint8_t *)ctx->ws_stage_in.ptr, *dcanvas = (uint8_t *)ctx->ws_stage_out.ptr;
    for (int i = 0; i < n_tiles; ++i) {
        dt[i].data = din + off[i];
        SRF_CHECK_CUDA(ctx, cudaMemcpy2DAsync(din + off[i], dt[i].pitch, tiles[i].data, tiles[i].pitch, (size_t)tiles[i].w * channels, tiles[i].h,
                                              cudaMemcpyHostToDevice, st));
        h2d += (size_t)tiles[i].w * channels * tiles[i].h;
    }
    if ((rc = merge_device(ctx, dt.data(), n_tiles, channels, out_h, out_w, dcanvas, cpitch, st))) return rc;
    SRF_CHECK_CUDA(ctx, cudaMemcpy2DAsync(h_canvas, canvas_pitch, dcanvas, cpitch, (size_t)out_w * channels, out_h, cudaMemcpyDeviceToHost, st));
    SRF_CHECK_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->last_h2d = h2d;
    ctx->last_d2h = (size_t)out_w * channels * out_h;
    return SRF_OK;
}
