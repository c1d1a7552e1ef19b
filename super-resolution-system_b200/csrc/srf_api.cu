// srf_api.cu -- context management, host-exact tiling geometry, tile extraction/padding, and the
// interleaved-float32 pyrDown / pyrUp entry points (the API-compat ones; the fusion hot path has its own
// planar kernels in srf_fuse.cu).
#include "srf_common.cuh"

#include <math.h>
#include <stdarg.h>
#include <string.h>
#include <unistd.h>

// ----------------------------------------------------------------------------------------------------------
// errors / context
// ----------------------------------------------------------------------------------------------------------
static thread_local std::string g_global_error;

void srf_set_error(srf_ctx *ctx, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->last_error = buf;
    g_global_error = buf;
}
void srf_set_error(const srf_ctx *ctx, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) const_cast<srf_ctx *>(ctx)->last_error = buf;
    g_global_error = buf;
}

int srf_reserve(srf_ctx *ctx, srf_buffer &b, size_t bytes)
{
    if (bytes <= b.bytes) return SRF_OK;
    ctx->ws_epoch++; // pointers handed to captured graphs / cached descriptors are about to change
    if (&b == &ctx->ws_pyr) ctx->pyr_epoch++;
    ctx->desc_sig = 0;
    if (b.ptr) {
        SRF_CHECK_CUDA(ctx, cudaDeviceSynchronize());
        SRF_CHECK_CUDA(ctx, cudaFree(b.ptr));
        b.ptr = nullptr;
        b.bytes = 0;
    }
    size_t want = srf_align_up(bytes + bytes / 8, (size_t)1 << 20);
    cudaError_t e = cudaMalloc(&b.ptr, want);
    if (e != cudaSuccess) {
        want = srf_align_up(bytes, 256);
        e = cudaMalloc(&b.ptr, want);
    }
    if (e != cudaSuccess) {
        b.ptr = nullptr;
        srf_set_error(ctx, "cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
        (void)cudaGetLastError();
        return SRF_ERR_OOM;
    }
    b.bytes = want;
    return SRF_OK;
}

int srf_reserve_pinned(srf_ctx *ctx, size_t bytes)
{
    if (bytes <= ctx->pinned_bytes) return SRF_OK;
    if (ctx->pinned) {
        SRF_CHECK_CUDA(ctx, cudaDeviceSynchronize());
        cudaFreeHost(ctx->pinned);
        ctx->pinned = nullptr;
        ctx->pinned_bytes = 0;
    }
    SRF_CHECK_CUDA(ctx, cudaMallocHost(&ctx->pinned, bytes));
    ctx->pinned_bytes = bytes;
    return SRF_OK;
}

// one warp per row: 16-byte loads from the mapped host pointer while both ends are 16-byte aligned, bytes for the tail
__global__ void __launch_bounds__(256) k_copy_rows_from_host(uint8_t *__restrict__ dst, size_t dpitch, const uint8_t *__restrict__ src, size_t spitch,
                                                             unsigned width, unsigned rows)
{
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (unsigned r = warp; r < rows; r += nwarps) {
        const uint8_t *s = src + (size_t)r * spitch;
        uint8_t *d = dst + (size_t)r * dpitch;
        unsigned body = 0;
        if ((((uintptr_t)s | (uintptr_t)d) & 15) == 0) {
            body = width & ~15u;
            for (unsigned o = lane * 16; o < body; o += 32 * 16) *reinterpret_cast<uint4 *>(d + o) = *reinterpret_cast<const uint4 *>(s + o);
        }
        for (unsigned o = body + lane; o < width; o += 32) d[o] = s[o];
    }
}

int srf_copy2d_h2d(srf_ctx *ctx, void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t rows, cudaStream_t st)
{
    if (width == 0 || rows == 0) return SRF_OK;
    if (rows >= 64 && width * 2 <= spitch && width < (1u << 31) && rows < (1u << 31)) { // narrow region: is the source mapped pinned memory?
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, src) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer != nullptr) {
            const unsigned blocks = (unsigned)std::min<size_t>((rows + 7) / 8, 2048);
            k_copy_rows_from_host<<<blocks, 256, 0, st>>>((uint8_t *)dst, dpitch, (const uint8_t *)at.devicePointer, spitch, (unsigned)width, (unsigned)rows);
            SRF_CHECK_CUDA(ctx, cudaGetLastError());
            ctx->launches++;
            return SRF_OK;
        }
        (void)cudaGetLastError(); // pageable memory: not an error, the copy engine takes it
    }
    SRF_CHECK_CUDA(ctx, cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, cudaMemcpyHostToDevice, st));
    return SRF_OK;
}

int srf_note_device_work(srf_ctx *ctx, cudaStream_t st)
{
    if (st == ctx->own_stream) return SRF_OK; // a host entry point calling its device sibling: same stream, already ordered
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) (void)cudaGetLastError();
    if (cs != cudaStreamCaptureStatusNone) return SRF_OK; // the caller captures a graph: ordering is the caller's business
    SRF_CHECK_CUDA(ctx, cudaEventRecord(ctx->ev_dev, st));
    ctx->ev_dev_set = true;
    return SRF_OK;
}

int srf_host_enter(srf_ctx *ctx)
{
    if (!ctx->ev_dev_set) return SRF_OK;
    SRF_CHECK_CUDA(ctx, cudaStreamWaitEvent(ctx->own_stream, ctx->ev_dev, 0));
    SRF_CHECK_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_dev, 0));
    SRF_CHECK_CUDA(ctx, cudaStreamWaitEvent(ctx->down_stream, ctx->ev_dev, 0));
    ctx->ev_dev_set = false;
    return SRF_OK;
}

extern "C" int srf_version(void) { return SRF_VERSION; }

extern "C" const char *srf_status_string(int s)
{
    switch (s) {
    case SRF_OK: return "ok";
    case SRF_ERR_ARG: return "invalid argument";
    case SRF_ERR_CUDA: return "CUDA error";
    case SRF_ERR_NO_DEVICE: return "no usable CUDA device (libsrfuse has no CPU fallback)";
    case SRF_ERR_OOM: return "out of device memory";
    case SRF_ERR_UNSUPPORTED: return "unsupported";
    default: return "unknown status";
    }
}

extern "C" const char *srf_last_error(const srf_ctx *ctx) { return ctx ? ctx->last_error.c_str() : g_global_error.c_str(); }

extern "C" int srf_create(int device, srf_ctx **out)
{
    if (!out) return SRF_ERR_ARG;
    *out = nullptr;
    // a driverless machine can stall inside the first CUDA call; the device node is the cheap, certain test
    if (access("/dev/nvidiactl", F_OK) != 0) {
        srf_set_error((srf_ctx *)nullptr, "no NVIDIA device node (/dev/nvidiactl): libsrfuse has no CPU fallback");
        return SRF_ERR_NO_DEVICE;
    }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        srf_set_error((srf_ctx *)nullptr, "no CUDA device: %s", e == cudaSuccess ? "count is 0" : cudaGetErrorString(e));
        (void)cudaGetLastError();
        return SRF_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) {
        srf_set_error((srf_ctx *)nullptr, "device %d out of range (have %d)", device, n);
        return SRF_ERR_ARG;
    }
    cudaDeviceProp p = {};
    e = cudaGetDeviceProperties(&p, device);
    if (e != cudaSuccess) {
        srf_set_error((srf_ctx *)nullptr, "cudaGetDeviceProperties(%d) failed: %s", device, cudaGetErrorString(e));
        (void)cudaGetLastError();
        return SRF_ERR_NO_DEVICE;
    }
    if (p.major < 10) {
        srf_set_error((srf_ctx *)nullptr, "device %d is sm_%d%d; libsrfuse is built for sm_100a only", device, p.major, p.minor);
        return SRF_ERR_NO_DEVICE;
    }
    srf_ctx *c = new srf_ctx();
    c->device = device;
    c->sm_count = p.multiProcessorCount;
    const int rc = [&]() -> int { // any failure below destroys the half-built context (streams and events included)
        SRF_CHECK_CUDA(c, cudaSetDevice(device));
        SRF_CHECK_CUDA(c, cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
        SRF_CHECK_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        SRF_CHECK_CUDA(c, cudaStreamCreateWithFlags(&c->down_stream, cudaStreamNonBlocking));
        SRF_CHECK_CUDA(c, cudaEventCreate(&c->ev0));
        SRF_CHECK_CUDA(c, cudaEventCreate(&c->ev1));
        SRF_CHECK_CUDA(c, cudaEventCreateWithFlags(&c->ev_dev, cudaEventDisableTiming));
        SRF_CHECK_CUDA(c, cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        for (int i = 0; i < srf_ctx::kForkStreams; ++i) {
            SRF_CHECK_CUDA(c, cudaStreamCreateWithFlags(&c->fork_stream[i], cudaStreamNonBlocking));
            SRF_CHECK_CUDA(c, cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming));
        }
        for (cudaEvent_t &ev : c->ev_aux) SRF_CHECK_CUDA(c, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        return SRF_OK;
    }();
    if (rc != SRF_OK) {
        srf_set_error((srf_ctx *)nullptr, "srf_create: %s", c->last_error.c_str());
        srf_destroy(c);
        return rc;
    }
    *out = c;
    return SRF_OK;
}

extern "C" void srf_destroy(srf_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    srf_comm_free(c);
    srf_buffer *bs[] = {&c->ws_pyr, &c->ws_desc, &c->ws_stage_in, &c->ws_stage_out, &c->ws_iqa, &c->ws_tab, &c->ws_merge};
    for (srf_buffer *b : bs)
        if (b->ptr) cudaFree(b->ptr);
    if (c->fuse_graph) cudaGraphExecDestroy(c->fuse_graph);
    if (c->pinned) cudaFreeHost(c->pinned);
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_aux)
        if (e) cudaEventDestroy(e);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev_dev) cudaEventDestroy(c->ev_dev);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    for (int i = 0; i < srf_ctx::kForkStreams; ++i) {
        if (c->ev_join[i]) cudaEventDestroy(c->ev_join[i]);
        if (c->fork_stream[i]) cudaStreamDestroy(c->fork_stream[i]);
    }
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->down_stream) cudaStreamDestroy(c->down_stream);
    delete c;
}

extern "C" int srf_sync(srf_ctx *ctx, void *stream)
{
    if (!ctx) return SRF_ERR_ARG;
    SRF_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    SRF_CHECK_CUDA(ctx, cudaStreamSynchronize((cudaStream_t)stream));
    return SRF_OK;
}

extern "C" int64_t srf_launch_count(const srf_ctx *ctx) { return ctx ? ctx->launches : -1; }
extern "C" int64_t srf_graph_replays(const srf_ctx *ctx) { return ctx ? ctx->graph_replays : -1; }

extern "C" int srf_last_transfer_bytes(const srf_ctx *c, size_t *h2d, size_t *d2h)
{
    if (!c) return SRF_ERR_ARG;
    if (h2d) *h2d = c->last_h2d;
    if (d2h) *d2h = c->last_d2h;
    return SRF_OK;
}

extern "C" size_t srf_workspace_bytes(const srf_ctx *c)
{
    if (!c) return 0;
    return c->ws_pyr.bytes + c->ws_desc.bytes + c->ws_stage_in.bytes + c->ws_stage_out.bytes + c->ws_iqa.bytes + c->ws_tab.bytes + c->ws_merge.bytes;
}

extern "C" int srf_profile_enable(srf_ctx *c, int on)
{
    if (!c) return SRF_ERR_ARG;
    c->profile = on != 0;
    return SRF_OK;
}

extern "C" int srf_profile_read(srf_ctx *c, srf_profile_entry *out, int cap)
{
    if (!c) return SRF_ERR_ARG;
    SRF_CHECK_CUDA(c, cudaSetDevice(c->device));
    SRF_CHECK_CUDA(c, cudaDeviceSynchronize());
    const int n = (int)c->prof.size();
    for (int i = 0; i < n; ++i) {
        srf_ctx::ProfRec &r = c->prof[i];
        if (out && i < cap) {
            memset(&out[i], 0, sizeof out[i]);
            strncpy(out[i].name, r.name, sizeof(out[i].name) - 1);
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) {
                (void)cudaGetLastError();
                ms = -1.f;
            }
            out[i].ms = ms;
            out[i].bytes = r.bytes;
        }
        c->ev_pool.push_back(r.e0);
        c->ev_pool.push_back(r.e1);
    }
    c->prof.clear();
    return n;
}

extern "C" float srf_last_device_ms(srf_ctx *c)
{
    if (!c || !c->ev_valid) return -1.f;
    float ms = -1.f;
    if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) != cudaSuccess) {
        (void)cudaGetLastError();
        return -1.f;
    }
    return ms;
}

// ----------------------------------------------------------------------------------------------------------
// tiling geometry (host, exact integer restatement of tiling_module.py:572-646)
// ----------------------------------------------------------------------------------------------------------
static int tiles_along(int dim, int block, int overlap)
{
    const int step = block - overlap;
    // max(1, ceil((dim - overlap) / step)) with a possibly negative numerator (ceil toward +inf)
    const int num = dim - overlap;
    int n = (num > 0) ? (num + step - 1) / step : 0;
    return n < 1 ? 1 : n;
}

extern "C" int srf_tile_positions(int width, int height, int block, int overlap, int *xywh, int cap)
{
    if (width < 1 || height < 1 || block < 1 || overlap < 0 || overlap >= block) {
        srf_set_error((srf_ctx *)nullptr, "srf_tile_positions: bad geometry W=%d H=%d block=%d overlap=%d", width, height, block, overlap);
        return SRF_ERR_ARG;
    }
    const int step = block - overlap;
    const int nx = tiles_along(width, block, overlap), ny = tiles_along(height, block, overlap);
    int k = 0;
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i, ++k) {
            if (xywh && k < cap) {
                const int x = i * step, y = j * step;
                xywh[4 * k + 0] = x;
                xywh[4 * k + 1] = y;
                xywh[4 * k + 2] = block < width - x ? block : width - x;
                xywh[4 * k + 3] = block < height - y ? block : height - y;
            }
        }
    return nx * ny;
}

extern "C" int srf_tile_overlaps(int x, int y, int w, int h, int width, int height, int block, int overlap, int *o)
{
    if (!o) return SRF_ERR_ARG;
    int top = y > 0 ? overlap : 0, left = x > 0 ? overlap : 0;
    int bottom = (y + h < height) ? overlap : 0, right = (x + w < width) ? overlap : 0;
    if (y + block >= height) {
        bottom = block - (height - y) - top;
        if (bottom < 0) bottom = 0;
    }
    if (x + block >= width) {
        right = block - (width - x) - left;
        if (right < 0) right = 0;
    }
    o[0] = top;
    o[1] = bottom;
    o[2] = left;
    o[3] = right;
    return SRF_OK;
}

// ----------------------------------------------------------------------------------------------------------
// tile extraction + bottom/right padding (pure gather; HBM-bound)
// ----------------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int border_index(int i, int n, int mode)
{
    if (i >= 0 && i < n) return i;
    if (mode == SRF_BORDER_CONSTANT) return -1;
    if (mode == SRF_BORDER_REPLICATE) return i < 0 ? 0 : n - 1;
    if (n == 1) return 0;
    const int delta = (mode == SRF_BORDER_MIRROR) ? 1 : 0; // REFLECT_101 skips the edge sample, REFLECT repeats it
    while (i < 0 || i >= n) i = (i < 0) ? (-i - 1 + delta) : (2 * n - 1 - i - delta);
    return i;
}

// cv2.copyMakeBorder of a crop: output tile t is (out_h, out_w, C); its pixel (r, px) is source pixel
// (border_index(r - top, h), border_index(px - left, w)) of the crop src[y0:y0+h, x0:x0+w] (0 for BORDER_CONSTANT).
// One CTA per (output row, tile); the row is out_w*C contiguous bytes produced 16 B per thread.  The BODY of the row (the
// bytes copied verbatim from the source row) is read as five aligned 32-bit words per chunk and funnel-shifted into
// place -- the crop's first byte sits at an arbitrary address -- so body traffic is word loads and 16-byte stores; only
// the chunks that touch the border take the per-byte gather.  vec_ok: every output row is 16-byte aligned.
template <int C>
__global__ void __launch_bounds__(256) k_tile_extract_pad(const uint8_t *__restrict__ src, size_t pitch, const int4 *__restrict__ xywh,
                                                          int out_h, int out_w, int top, int left, int mode, uint8_t *__restrict__ out, int vec_ok)
{
    const int t = blockIdx.y;
    const int r = blockIdx.x;
    const int4 q = xywh[t];
    const int x0 = q.x, y0 = q.y, w = q.z, h = q.w;
    const int sr = border_index(r - top, h, mode);
    uint8_t *orow = out + ((size_t)t * out_h + r) * (size_t)out_w * C;
    const uint8_t *srow = (sr >= 0) ? src + (size_t)(y0 + sr) * pitch + (size_t)x0 * C : nullptr;
    const int nbytes = out_w * C;
    const int body0 = left * C, body1 = body0 + w * C; // output bytes [body0, body1) are source bytes [0, w*C) of srow
    auto gather = [&](int byte) -> uint8_t {
        if (!srow) return 0;
        const int px = byte / C, c = byte - px * C;
        const int sx = border_index(px - left, w, mode);
        return (sx >= 0) ? srow[sx * C + c] : (uint8_t)0;
    };
    if (vec_ok) {
        for (int k = threadIdx.x * 16; k < nbytes; k += blockDim.x * 16) {
            uint4 v;
            const uint8_t *p = srow + (k - body0);
            const int sh = (int)(reinterpret_cast<uintptr_t>(p) & 3) * 8;
            // a misaligned chunk reads a fifth word, i.e. up to 3 bytes beyond its own 16: allowed only while those
            // still belong to the body of this row (the last chunk of a misaligned row takes the byte path instead)
            if (srow && k >= body0 && k + 16 <= body1 && (sh == 0 || k + 20 <= body1)) {
                const uint32_t *wp = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)3);
                const uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2), w3 = __ldg(wp + 3);
                const uint32_t w4 = sh ? __ldg(wp + 4) : 0u;
                v = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
            } else {
                uint32_t words[4];
                uint8_t *b = reinterpret_cast<uint8_t *>(words);
                if (k + 16 <= nbytes) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int byte = k + e;
                        b[e] = (srow && byte >= body0 && byte < body1) ? srow[byte - body0] : gather(byte);
                    }
                    v = make_uint4(words[0], words[1], words[2], words[3]);
                } else { // ragged end of a row whose length is not a multiple of 16 (cannot happen when vec_ok)
                    for (int byte = k; byte < nbytes; ++byte) orow[byte] = gather(byte);
                    continue;
                }
            }
            *reinterpret_cast<uint4 *>(orow + k) = v;
        }
    } else {
        for (int byte = threadIdx.x; byte < nbytes; byte += blockDim.x) orow[byte] = gather(byte);
    }
}

static int launch_extract_pad(srf_ctx *ctx, const uint8_t *d_src, int channels, size_t pitch, const int *xywh, int n, int out_h, int out_w, int top,
                              int left, int border, uint8_t *d_out, cudaStream_t st)
{
    {
        const int rc = srf_reserve(ctx, ctx->ws_desc, (size_t)n * sizeof(int4));
        if (rc) return rc;
    }
    ctx->desc_sig = 0; // ws_desc is shared with the fusion descriptors
    ctx->fuse_sig = 0;
    SRF_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->ws_desc.ptr, xywh, (size_t)n * sizeof(int4), cudaMemcpyHostToDevice, st));
    const int vec_ok = ((out_w * channels) % 16 == 0) && (((uintptr_t)d_out) % 16 == 0);
    dim3 grid(out_h, n);
    {
        double bytes = (double)n * out_h * out_w * channels;
        for (int i = 0; i < n; ++i) bytes += (double)xywh[4 * i + 2] * xywh[4 * i + 3] * channels;
        srf_launch_scope scope(ctx, "tile_extract_pad", bytes, st);
        if (channels == 3)
            k_tile_extract_pad<3><<<grid, 256, 0, st>>>(d_src, pitch, (const int4 *)ctx->ws_desc.ptr, out_h, out_w, top, left, border, d_out, vec_ok);
        else
            k_tile_extract_pad<1><<<grid, 256, 0, st>>>(d_src, pitch, (const int4 *)ctx->ws_desc.ptr, out_h, out_w, top, left, border, d_out, vec_ok);
    }
    SRF_CHECK_CUDA(ctx, cudaGetLastError());
    return srf_note_device_work(ctx, st);
}

extern "C" int srf_tile_extract_pad(srf_ctx *ctx, const uint8_t *d_src, int height, int width, int channels, size_t pitch,
                                    const int *xywh, int n, int block, int border, uint8_t *d_out, void *stream)
{
    if (!ctx) return SRF_ERR_ARG;
    SRF_REQUIRE(ctx, d_src && d_out && xywh, "srf_tile_extract_pad: null pointer");
    SRF_REQUIRE(ctx, channels == 1 || channels == 3, "srf_tile_extract_pad: channels must be 1 or 3 (got %d)", channels);
    SRF_REQUIRE(ctx, border >= 0 && border <= 3, "srf_tile_extract_pad: bad border mode %d", border);
    SRF_REQUIRE(ctx, n >= 0 && block >= 1 && block <= 65535 && height >= 1 && width >= 1, "srf_tile_extract_pad: bad sizes");
    SRF_REQUIRE(ctx, pitch >= (size_t)width * channels, "srf_tile_extract_pad: pitch < row bytes");
    if (n == 0) return SRF_OK;
    for (int i = 0; i < n; ++i) {
        const int x = xywh[4 * i], y = xywh[4 * i + 1], w = xywh[4 * i + 2], h = xywh[4 * i + 3];
        SRF_REQUIRE(ctx, x >= 0 && y >= 0 && w >= 1 && h >= 1 && x + w <= width && y + h <= height && w <= block && h <= block,
                    "srf_tile_extract_pad: tile %d (%d,%d,%d,%d) outside image %dx%d / block %d", i, x, y, w, h, width, height, block);
    }
    SRF_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    return launch_extract_pad(ctx, d_src, channels, pitch, xywh, n, block, block, 0, 0, border, d_out, (cudaStream_t)stream);
}

extern "C" int srf_tile_extract_pad_host(srf_ctx *ctx, const uint8_t *h_src, int height, int width, int channels, size_t pitch,
                                         const int *xywh, int n, int block, int border, uint8_t *h_out)
{
    if (!ctx) return SRF_ERR_ARG;
    SRF_REQUIRE(ctx, h_src && h_out && xywh, "srf_tile_extract_pad_host: null pointer");
    SRF_REQUIRE(ctx, channels == 1 || channels == 3, "srf_tile_extract_pad_host: channels must be 1 or 3 (got %d)", channels);
    SRF_REQUIRE(ctx, n >= 0 && block >= 1 && height >= 1 && width >= 1, "srf_tile_extract_pad_host: bad sizes (n=%d block=%d image %dx%d)", n,
                block, width, height);
    SRF_REQUIRE(ctx, pitch >= (size_t)width * channels, "srf_tile_extract_pad_host: pitch < row bytes");
    SRF_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (int hrc = srf_host_enter(ctx)) return hrc;
    const size_t dpitch = srf_align_up((size_t)width * channels, 256);
    const size_t out_bytes = (size_t)n * block * block * channels;
    int rc = srf_reserve(ctx, ctx->ws_stage_in, dpitch * height);
    if (rc) return rc;
    rc = srf_reserve(ctx, ctx->ws_stage_out, out_bytes ? out_bytes : 16);
    if (rc) return rc;
    cudaStream_t st = ctx->own_stream;
    SRF_CHECK_CUDA(ctx, cudaMemcpy2DAsync(ctx->ws_stage_in.ptr, dpitch, h_src, pitch, (size_t)width * channels, height,
                                          cudaMemcpyHostToDevice, st));
    SRF_CHECK_CUDA(ctx, cudaEventRecord(ctx->ev0, st));
    rc = srf_tile_extract_pad(ctx, (const uint8_t *)ctx->ws_stage_in.ptr, height, width, channels, dpitch, xywh, n, block, border,
                              (uint8_t *)ctx->ws_stage_out.ptr, st);
    if (rc) return rc;
    SRF_CHECK_CUDA(ctx, cudaEventRecord(ctx->ev1, st));
    ctx->ev_valid = true;
    if (out_bytes) SRF_CHECK_CUDA(ctx, cudaMemcpyAsync(h_out, ctx->ws_stage_out.ptr, out_bytes, cudaMemcpyDeviceToHost, st));
    SRF_CHECK_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->last_h2d = (size_t)width * channels * height;
    ctx->last_d2h = out_bytes;
    return SRF_OK;
}

// cv2.copyMakeBorder(image, top, bottom, left, right, mode) -- TilingModule._apply_padding in full generality
// (tiling_module.py:522-570; split_image itself only ever pads bottom / right)
static int check_border_args(srf_ctx *ctx, const void *s, const void *d, int h, int w, int C, size_t pitch, int top, int bottom, int left, int right,
                             int border)
{
    SRF_REQUIRE(ctx, s && d, "srf_copy_make_border: null pointer");
    SRF_REQUIRE(ctx, C == 1 || C == 3, "srf_copy_make_border: channels must be 1 or 3 (got %d)", C);
    SRF_REQUIRE(ctx, h >= 1 && w >= 1 && pitch >= (size_t)w * C, "srf_copy_make_border: bad image %dx%d / pitch", w, h);
    SRF_REQUIRE(ctx, top >= 0 && bottom >= 0 && left >= 0 && right >= 0, "srf_copy_make_border: negative padding");
    SRF_REQUIRE(ctx, border >= 0 && border <= 3, "srf_copy_make_border: bad border mode %d", border);
    SRF_REQUIRE(ctx, (long long)h + top + bottom <= 65535 * 32LL && (long long)w + left + right <= (1 << 28), "srf_copy_make_border: result too large");
    return SRF_OK;
}

extern "C" int srf_copy_make_border(srf_ctx *ctx, const uint8_t *d_src, int h, int w, int channels, size_t pitch, int top, int bottom, int left,
                                    int right, int border, uint8_t *d_dst, void *stream)
{
    if (!ctx) return SRF_ERR_ARG;
    int rc = check_border_args(ctx, d_src, d_dst, h, w, channels, pitch, top, bottom, left, right, border);
    if (rc) return rc;
    SRF_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    const int xywh[4] = {0, 0, w, h};
    return launch_extract_pad(ctx, d_src, channels, pitch, xywh, 1, h + top + bottom, w + left + right, top, left, border, d_dst, (cudaStream_t)stream);
}

extern "C" int srf_copy_make_border_host(srf_ctx *ctx, const uint8_t *h_src, int h, int w, int channels, size_t pitch, int top, int bottom,
                                         int left, int right, int border, uint8_t *h_dst)
{
    if (!ctx) return SRF_ERR_ARG;
    int rc = check_border_args(ctx, h_src, h_dst, h, w, channels, pitch, top, bottom, left, right, border);
    if (rc) return rc;
    SRF_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (int hrc = srf_host_enter(ctx)) return hrc;
    const size_t dpitch = srf_align_up((size_t)w * channels, 256);
    const size_t out_bytes = (size_t)(h + top + bottom) * (size_t)(w + left + right) * channels;
    if ((rc = srf_reserve(ctx, ctx->ws_stage_in, dpitch * h))) return rc;
    if ((rc = srf_reserve(ctx, ctx->ws_stage_out, out_bytes))) return rc;
    cudaStream_t st = ctx->own_stream;
    SRF_CHECK_CUDA(ctx, cudaMemcpy2DAsync(ctx->ws_stage_in.ptr, dpitch, h_src, pitch, (size_t)w * channels, h, cudaMemcpyHostToDevice, st));
    const int xywh[4] = {0, 0, w, h};
    if ((rc = launch_extract_pad(ctx, (const uint8_t *)ctx->ws_stage_in.ptr, channels, dpitch, xywh, 1, h + top + bottom, w + left + right, top, left,
                                 border, (uint8_t *)ctx->ws_stage_out.ptr, st)))
        return rc;
    SRF_CHECK_CUDA(ctx, cudaMemcpyAsync(h_dst, ctx->ws_stage_out.ptr, out_bytes, cudaMemcpyDeviceToHost, st));
    SRF_CHECK_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->last_h2d = (size_t)w * channels * h;
    ctx->last_d2h = out_bytes;
    return SRF_OK;
}

// ----------------------------------------------------------------------------------------------------------
// interleaved float32 pyrDown / pyrUp (API-compat entry points; simple one-thread-per-output kernels that read
// through L1/L2 -- they exist so build_gaussian_pyramid & co. stay drop-in and unit-testable against cv2)
// ----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pyr_down_hwc(const float *__restrict__ src, int h, int w, int cn, float *__restrict__ dst, int oh,
                                                      int ow, int vec_end, int vtail)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x; // float index within the output row
    const int y = blockIdx.y;
    if (k >= ow * cn) return;
    const int x = k / cn, c = k - x * cn;
    int xi[5], yi[5];
#pragma unroll
    for (int t = 0; t < 5; ++t) {
        xi[t] = srf_reflect101(2 * x + t - 2, w) * cn + c;
        yi[t] = srf_reflect101(2 * y + t - 2, h);
    }
    const bool hs = (x < 1) || (x >= vec_end);
    float r[5];
#pragma unroll
    for (int t = 0; t < 5; ++t) {
        const float *s = src + (size_t)yi[t] * w * cn;
        const float a0 = s[xi[0]], a1 = s[xi[1]], a2 = s[xi[2]], a3 = s[xi[3]], a4 = s[xi[4]];
        r[t] = hs ? dn_h_sca(a0, a1, a2, a3, a4) : dn_h_vec(a0, a1, a2, a3, a4);
    }
    dst[(size_t)y * ow * cn + k] = (k < vtail) ? dn_v_vec(r[0], r[1], r[2], r[3], r[4]) : dn_v_sca(r[0], r[1], r[2], r[3], r[4]);
}

__global__ void __launch_bounds__(256) k_pyr_up_hwc(const float *__restrict__ src, int h, int w, int cn, float *__restrict__ dst, int dh,
                                                    int dw)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int ty = blockIdx.y;
    if (k >= dw * cn) return;
    const int tx = k / cn, c = k - tx * cn;
    const int j = ty >> 1;
    const int jm = srf_reflect101(j - 1, h), jn = (j + 1 < h) ? j + 1 : h - 1;
    auto hrow = [&](int row) {
        const float *s = src + (size_t)row * w * cn + c;
        return up_h_sample(tx, w, [&](int i) { return s[(size_t)i * cn]; });
    };
    const float p = hrow(j);
    float v;
    if (ty & 1)
        v = up_odd(p, hrow(jn));
    else
        v = up_even(hrow(jm), p, hrow(jn));
    dst[(size_t)ty * dw * cn + k] = __fmul_rn(v, 1.f / 64.f);
}

int srf_launch_pyr_down_hwc(srf_ctx *ctx, const float *src, int h, int w, int cn, float *dst, cudaStream_t st)
{
    const int oh = (h + 1) / 2, ow = (w + 1) / 2;
    dim3 grid(srf_ceil_div(ow * cn, 256), oh);
    {
        srf_launch_scope scope(ctx, "pyr_down_hwc", 4.0 * cn * ((double)h * w + (double)oh * ow), st);
        k_pyr_down_hwc<<<grid, 256, 0, st>>>(src, h, w, cn, dst, oh, ow, srf_down_vec_end(w, ow, cn), 4 * ((ow * cn) / 4));
    }
    SRF_CHECK_CUDA(ctx, cudaGetLastError());
    return SRF_OK;
}

int srf_launch_pyr_up_hwc(srf_ctx *ctx, const float *src, int h, int w, int cn, float *dst, int dh, int dw, cudaStream_t st)
{
    dim3 grid(srf_ceil_div(dw * cn, 256), dh);
    {
        srf_launch_scope scope(ctx, "pyr_up_hwc", 4.0 * cn * ((double)h * w + (double)dh * dw), st);
        k_pyr_up_hwc<<<grid, 256, 0, st>>>(src, h, w, cn, dst, dh, dw);
    }
    SRF_CHECK_CUDA(ctx, cudaGetLastError());
    return SRF_OK;
}

static int check_pyr_args(srf_ctx *ctx, const void *s, const void *d, int h, int w, int cn, const char *who)
{
    SRF_REQUIRE(ctx, s && d, "%s: null pointer", who);
    SRF_REQUIRE(ctx, h >= 1 && w >= 1, "%s: empty image %dx%d", who, h, w);
    SRF_REQUIRE(ctx, cn >= 1 && cn <= 4, "%s: channels must be 1..4 (got %d)", who, cn);
    SRF_REQUIRE(ctx, h <= 65535 * 2, "%s: image too tall", who);
    return SRF_OK;
}

extern "C" int srf_pyr_down_f32(srf_ctx *ctx, const float *d_src, int h, int w, int cn, float *d_dst, void *stream)
{
    if (!ctx) return SRF_ERR_ARG;
    int rc = check_pyr_args(ctx, d_src, d_dst, h, w, cn, "srf_pyr_down_f32");
    if (rc) return rc;
    SRF_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    return srf_launch_pyr_down_hwc(ctx, d_src, h, w, cn, d_dst, (cudaStream_t)stream);
}

extern "C" int srf_pyr_up_f32(srf_ctx *ctx, const float *d_src, int h, int w, int cn, float *d_dst, int dh, int dw, void *stream)
{
    if (!ctx) return SRF_ERR_ARG;
    int rc = check_pyr_args(ctx, d_src, d_dst, h, w, cn, "srf_pyr_up_f32");
    if (rc) return rc;
    SRF_REQUIRE(ctx, (dh == 2 * h || dh == 2 * h - 1) && (dw == 2 * w || dw == 2 * w - 1) && dh >= 1 && dw >= 1,
                "srf_pyr_up_f32: dst %dx%d must be 2n or 2n-1 of src %dx%d", dh, dw, h, w);
    SRF_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    return srf_launch_pyr_up_hwc(ctx, d_src, h, w, cn, d_dst, dh, dw, (cudaStream_t)stream);
}

extern "C" int srf_pyr_down_f32_host(srf_ctx *ctx, const float *h_src, int h, int w, int cn, float *h_dst)
{
    if (!ctx) return SRF_ERR_ARG;
    int rc = check_pyr_args(ctx, h_src, h_dst, h, w, cn, "srf_pyr_down_f32_host");
    if (rc) return rc;
    SRF_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (int hrc = srf_host_enter(ctx)) return hrc;
    const size_t ib = (size_t)h * w * cn * 4, ob = (size_t)((h + 1) / 2) * ((w + 1) / 2) * cn * 4;
    if ((rc = srf_reserve(ctx, ctx->ws_stage_in, ib))) return rc;
    if ((rc = srf_reserve(ctx, ctx->ws_stage_out, ob))) return rc;
    cudaStream_t st = ctx->own_stream;
    SRF_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->ws_stage_in.ptr, h_src, ib, cudaMemcpyHostToDevice, st));
    if ((rc = srf_launch_pyr_down_hwc(ctx, (const float *)ctx->ws_stage_in.ptr, h, w, cn, (float *)ctx->ws_stage_out.ptr, st))) return rc;
    SRF_CHECK_CUDA(ctx, cudaMemcpyAsync(h_dst, ctx->ws_stage_out.ptr, ob, cudaMemcpyDeviceToHost, st));
    SRF_CHECK_CUDA(ctx, cudaStreamSynchronize(st));
    return SRF_OK;
}

extern "C" int srf_pyr_up_f32_host(srf_ctx *ctx, const float *h_src, int h, int w, int cn, float *h_dst, int dh, int dw)
{
    if (!ctx) return SRF_ERR_ARG;
    int rc = check_pyr_args(ctx, h_src, h_dst, h, w, cn, "srf_pyr_up_f32_host");
    if (rc) return rc;
    SRF_REQUIRE(ctx, (dh == 2 * h || dh == 2 * h - 1) && (dw == 2 * w || dw == 2 * w - 1) && dh >= 1 && dw >= 1,
                "srf_pyr_up_f32_host: dst %dx%d must be 2n or 2n-1 of src %dx%d", dh, dw, h, w);
    SRF_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (int hrc = srf_host_enter(ctx)) return hrc;
    const size_t ib = (size_t)h * w * cn * 4, ob = (size_t)dh * dw * cn * 4;
    if ((rc = srf_reserve(ctx, ctx->ws_stage_in, ib))) return rc;
    if ((rc = srf_reserve(ctx, ctx->ws_stage_out, ob))) return rc;
    cudaStream_t st = ctx->own_stream;
    SRF_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->ws_stage_in.ptr, h_src, ib, cudaMemcpyHostToDevice, st));
    if ((rc = srf_launch_pyr_up_hwc(ctx, (const float *)ctx->ws_stage_in.ptr, h, w, cn, (float *)ctx->ws_stage_out.ptr, dh, dw, st))) return rc;
    SRF_CHECK_CUDA(ctx, cudaMemcpyAsync(h_dst, ctx->ws_stage_out.ptr, ob, cudaMemcpyDeviceToHost, st));
    SRF_CHECK_CUDA(ctx, cudaStreamSynchronize(st));
    return SRF_OK;
}

// ----------------------------------------------------------------------------------------------------------
// small host helpers
// ----------------------------------------------------------------------------------------------------------
extern "C" int srf_weight_profile(int weight_type, int feather, float *out)
{
    if (!out || feather < 1 || weight_type < 0 || weight_type > 2) return SRF_ERR_ARG;
    const double pi = 3.141592653589793;
    for (int d = 0; d <= feather; ++d) {
        double nd = (double)d / (double)feather;
        if (nd < 0) nd = 0;
        if (nd > 1) nd = 1;
        double v;
        if (weight_type == SRF_WEIGHT_COSINE)
            v = 0.5 * (1 - cos(pi * nd));
        else if (weight_type == SRF_WEIGHT_SIGMOID)
            v = 1 / (1 + exp(-10 * (nd - 0.5)));
        else
            v = nd;
        out[d] = (float)v;
    }
    return SRF_OK;
}

extern "C" int srf_iqa_finish(const srf_iqa_sums *s, double data_range, double *psnr, double *ssim)
{
    if (!s) return SRF_ERR_ARG;
    if (psnr) {
        if (s->n_elems <= 0)
            *psnr = NAN;
        else if (s->sq_err == 0)
            *psnr = INFINITY;
        else
            *psnr = 10.0 * log10((data_range * data_range) / ((double)s->sq_err / (double)s->n_elems));
    }
    if (ssim) *ssim = (s->ssim_count > 0) ? s->ssim_sum / (double)s->ssim_count : NAN;
    return SRF_OK;
}
