// srf_common.cuh -- shared declarations for libsrfuse (sm_100a).
//
// Float32 op ORDER is part of the contract: the reference path (cv2.pyrDown / cv2.pyrUp on float32 followed by a
// truncating uint8 cast, /root/reference/blending_module.py:263, 306, 358, 504-506) is reproduced rounding for
// rounding, so every arithmetic step below is an explicit __f*_rn intrinsic (never contracted into an FMA).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <string>
#include <vector>
#include <map>

#include "../../include/srfuse.h"

#define SRF_CHECK_CUDA(ctx, expr)                                                                   \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            srf_set_error((ctx), "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return (_e == cudaErrorMemoryAllocation) ? SRF_ERR_OOM : SRF_ERR_CUDA;                  \
        }                                                                                           \
    } while (0)

#define SRF_REQUIRE(ctx, cond, ...)           \
    do {                                      \
        if (!(cond)) {                        \
            srf_set_error((ctx), __VA_ARGS__); \
            return SRF_ERR_ARG;               \
        }                                     \
    } while (0)

struct srf_buffer {
    void *ptr = nullptr;
    size_t bytes = 0;
};

struct srf_comm; // srf_comm.cu: peer-memory inboxes of a multi-GPU job
void srf_comm_free(struct srf_ctx *ctx);

struct srf_ctx {
    srf_comm *comm = nullptr;
    int device = 0;
    int sm_count = 148;
    std::string last_error;
    int64_t launches = 0;
    int64_t graph_replays = 0;
    size_t last_h2d = 0, last_d2h = 0; // bytes the last *_host entry point moved over PCIe
    // grow-only device workspaces, by role
    srf_buffer ws_pyr;      // per-tile G/C pyramids + weight pyramids
    srf_buffer ws_desc;     // device copies of descriptor arrays
    srf_buffer ws_stage_in; // *_host entry points: device staging of inputs
    srf_buffer ws_stage_out;
    srf_buffer ws_iqa;      // IQA reduction partials
    srf_buffer ws_tab;      // per-axis tap tables of the bicubic resize
    srf_buffer ws_merge;    // merge_tiles: tile descriptors, weight tables, squashed edge tiles
    void *pinned = nullptr; // small pinned host block for result read-back
    size_t pinned_bytes = 0;
    cudaStream_t own_stream = nullptr; // used by the *_host entry points
    cudaStream_t copy_stream = nullptr;
    cudaStream_t down_stream = nullptr; // device->host copies of the pipelined host entry point
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_aux[3] = {nullptr, nullptr, nullptr};
    bool ev_valid = false;
    std::vector<cudaEvent_t> ev_pool;
    // weight pyramids of the previous fuse call still valid in ws_pyr: one key per shape group (srf_fuse.cu)
    std::vector<uint64_t> wcache_keys;
    const void *wcache_base = nullptr;
    uint64_t wcache_epoch = ~0ull; // ws_epoch at which the cached weight pyramids were written (cudaFree + a larger cudaMalloc
                                   // may hand back the same address, so the base pointer alone proves nothing)
    // host entry points run on own_stream while device entry points enqueue on the caller's stream over the SAME
    // workspaces: every device-path enqueue leaves an event here and every host entry point makes its streams wait on it
    cudaEvent_t ev_dev = nullptr;
    bool ev_dev_set = false;
    // the gather's class launches write disjoint canvas blocks: they fork onto these streams and join back (srf_fuse.cu)
    static const int kForkStreams = 8; // >= the gather's classes - 1
    cudaStream_t fork_stream[kForkStreams] = {};
    cudaEvent_t ev_fork = nullptr, ev_join[kForkStreams] = {};
    // CUDA-graph replay of a repeated srf_laplacian_fuse call (same tiles, canvas, band, stream): the second identical
    // call is captured, later ones are one cudaGraphLaunch.  Any workspace reallocation bumps ws_epoch and drops it.
    uint64_t fuse_sig = 0;        // signature of the last srf_laplacian_fuse call
    uint64_t desc_sig = 0;        // signature whose descriptors currently sit in ws_desc
    uint64_t ws_epoch = 0, graph_epoch = 0;
    uint64_t pyr_epoch = 0; // bumped only when ws_pyr itself is re-allocated (the weight cache lives there)
    cudaGraphExec_t fuse_graph = nullptr;
    int64_t fuse_graph_launches = 0;
    // per-kernel profiling (srf_profile_enable)
    bool profile = false;
    struct ProfRec {
        const char *name;
        cudaEvent_t e0, e1;
        double bytes;
    };
    std::vector<ProfRec> prof;
};

// brackets one kernel launch with events when profiling is on; always counts the launch
struct srf_launch_scope {
    srf_ctx *ctx;
    cudaStream_t st;
    cudaEvent_t e1 = nullptr;
    const char *name;
    srf_launch_scope(srf_ctx *c, const char *name_, double bytes, cudaStream_t s) : ctx(c), st(s), name(name_)
    {
        ctx->launches++;
        if (!ctx->profile) return;
        cudaEvent_t e0;
        auto get = [&]() {
            cudaEvent_t e;
            if (!ctx->ev_pool.empty()) {
                e = ctx->ev_pool.back();
                ctx->ev_pool.pop_back();
            } else {
                cudaEventCreate(&e);
            }
            return e;
        };
        e0 = get();
        e1 = get();
        cudaEventRecord(e0, st);
        ctx->prof.push_back(srf_ctx::ProfRec{name, e0, e1, bytes});
    }
    ~srf_launch_scope()
    {
        if (e1) cudaEventRecord(e1, st);
        static const bool debug_sync = getenv("SRF_DEBUG_SYNC") != nullptr; // bring-up aid: name the kernel that faults
        if (debug_sync) {
            const cudaError_t e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) fprintf(stderr, "[srfuse] kernel '%s' (launch %lld): %s\n", name, (long long)ctx->launches, cudaGetErrorString(e));
        }
    }
};

void srf_set_error(srf_ctx *ctx, const char *fmt, ...);
void srf_set_error(const srf_ctx *ctx, const char *fmt, ...);
int srf_reserve(srf_ctx *ctx, srf_buffer &b, size_t bytes);
int srf_reserve_pinned(srf_ctx *ctx, size_t bytes);
// stream hygiene between the two families of entry points (they share the ctx workspaces):
//   srf_note_device_work  a device-path entry point calls it after enqueuing on the caller's stream
//   srf_host_enter        a *_host entry point calls it first: own / copy / down streams wait for that work
// host -> device copy of `rows` rows of `width` bytes (the *_host entry points' tile uploads).  The copy engine moves a
// 2-D region whose rows are much shorter than the pitch at a fifth of the link rate (measured: 2 790-byte rows of a
// 12 288-byte pitch, 11 GB/s against 51 GB/s for whole rows), so a narrow region of PINNED host memory is pulled by a
// kernel instead (16-byte loads through the mapped pointer); everything else stays a cudaMemcpy2DAsync.
int srf_copy2d_h2d(srf_ctx *ctx, void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t rows, cudaStream_t st);
int srf_note_device_work(srf_ctx *ctx, cudaStream_t st);
int srf_host_enter(srf_ctx *ctx);

static inline int srf_ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t srf_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

#ifdef __CUDACC__
// kernel launch with the programmatic-stream-serialization attribute (see pdl_enter): the launch latency and the ramp of a
// kernel overlap the tail of its predecessor -- what the short, dependent launches of the pyramid levels are made of
template <typename... KArgs, typename... Args>
static inline cudaError_t srf_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

// ---- index maps -------------------------------------------------------------------------------------------
// cv2 BORDER_REFLECT_101 with multi-bounce (index i on an axis of length n)
__host__ __device__ __forceinline__ int srf_reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * (n - 1) - i;
    return i;
}

// ---- pyrDown / pyrUp arithmetic, in OpenCV's float32 op order (see oracle/pyr_oracle.c) -------------------
#ifdef __CUDACC__
// horizontal 5-tap, SIMD-body order: r2*6 + ((r1+r3)*4 + (r0+r4))
__device__ __forceinline__ float dn_h_vec(float r0, float r1, float r2, float r3, float r4)
{
    float a = __fmul_rn(__fadd_rn(r1, r3), 4.f);
    float b = __fadd_rn(r0, r4);
    return __fadd_rn(__fmul_rn(r2, 6.f), __fadd_rn(a, b));
}
// horizontal 5-tap, scalar order (border columns and SIMD tail): ((r2*6 + (r1+r3)*4) + r0) + r4
__device__ __forceinline__ float dn_h_sca(float r0, float r1, float r2, float r3, float r4)
{
    float a = __fmul_rn(r2, 6.f);
    float b = __fmul_rn(__fadd_rn(r1, r3), 4.f);
    return __fadd_rn(__fadd_rn(__fadd_rn(a, b), r0), r4);
}
// vertical 5-tap, SIMD-body order: (((r1+r3)+r2)*4 + ((r0+r4)+(r2+r2))) * 1/256
__device__ __forceinline__ float dn_v_vec(float r0, float r1, float r2, float r3, float r4)
{
    float a = __fmul_rn(__fadd_rn(__fadd_rn(r1, r3), r2), 4.f);
    float b = __fadd_rn(__fadd_rn(r0, r4), __fadd_rn(r2, r2));
    return __fmul_rn(__fadd_rn(a, b), 1.f / 256.f);
}
__device__ __forceinline__ float dn_v_sca(float r0, float r1, float r2, float r3, float r4)
{
    return __fmul_rn(dn_h_sca(r0, r1, r2, r3, r4), 1.f / 256.f);
}
// pyrUp 1-D forms (un-normalised: each pass leaves a factor 8)
__device__ __forceinline__ float up_even(float m, float p, float n) { return __fadd_rn(__fadd_rn(__fmul_rn(p, 6.f), m), n); }
__device__ __forceinline__ float up_even_left(float p, float n) { return __fadd_rn(__fmul_rn(p, 6.f), __fmul_rn(n, 2.f)); }
__device__ __forceinline__ float up_even_right(float m, float p) { return __fadd_rn(m, __fmul_rn(p, 7.f)); }
__device__ __forceinline__ float up_odd(float p, float n) { return __fmul_rn(__fadd_rn(p, n), 4.f); }
__device__ __forceinline__ float up_odd_right(float p) { return __fmul_rn(p, 8.f); }

// horizontal pyrUp sample `tx` (0 .. 2*w-1) of a source row addressed through `at(i)`, i in [0, w)
template <class At>
__device__ __forceinline__ float up_h_sample(int tx, int w, At at)
{
    const int i = tx >> 1;
    const float p = at(i);
    if (w == 1) return __fmul_rn(p, 8.f);
    if (tx & 1) return (i == w - 1) ? up_odd_right(p) : up_odd(p, at(i + 1));
    if (i == 0) return up_even_left(p, at(1));
    if (i == w - 1) return up_even_right(at(i - 1), p);
    return up_even(at(i - 1), p, at(i + 1));
}

// ---- programmatic dependent launch: every kernel of the chain starts with pdl_enter().  launch_dependents lets the NEXT
// kernel of the stream be scheduled while this grid is still running (its CTAs become resident as SMs drain);
// griddepcontrol.wait then holds it until the PREVIOUS grid has completed and flushed its memory, so nothing before the
// wait may touch global data.  Without the launch attribute both instructions are no-ops. ------------------------------
__device__ __forceinline__ void pdl_enter()
{
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
}

// ---- cp.async (LDGSTS) helpers: 16-byte global -> shared copies that need no registers across the latency ------
__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) + mbarrier: one thread moves a whole contiguous row segment ------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk copy: src and dst 16-byte aligned, bytes a multiple of 16; completion is signalled on `bar`
__device__ __forceinline__ void tma_bulk_g2s(void *smem_dst, const void *gsrc, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
#endif

// number of output columns (pixels) of a pyrDown row that OpenCV's 4-lane SIMD body covers: columns
// [1, vec_end) use the vector op order, all others the scalar one (cn = channels of the REFERENCE's
// interleaved array: 3 for image pyramids, 1 for weight pyramids)
__host__ __device__ __forceinline__ int srf_down_vec_end(int w, int ow, int cn)
{
    int width0 = (w - 3) / 2 + 1;
    if (width0 > ow) width0 = ow;
    if (width0 < 1) return 1;
    if (cn == 1) return 1 + 4 * ((width0 - 1) / 4);
    if (cn == 3) return (width0 - 1 > 1) ? width0 - 1 : 1;
    if (cn == 2) return 1 + 2 * ((width0 - 1) / 2);
    if (cn == 4) return width0;
    return 1;
}

// ---- internal entry points shared between translation units ------------------------------------------------
int srf_launch_pyr_down_hwc(srf_ctx *ctx, const float *src, int h, int w, int cn, float *dst, cudaStream_t st);
int srf_launch_pyr_up_hwc(srf_ctx *ctx, const float *src, int h, int w, int cn, float *dst, int dh, int dw,
                          cudaStream_t st);

// fusion of device-resident tiles into a device canvas (srf_fuse.cu) and the IQA pass over device images (srf_iqa.cu)
int srf_fuse_device(srf_ctx *ctx, const srf_tile *tiles, int n_tiles, int C, int levels, const srf_weight_lut *luts, int n_luts, int out_h,
                    int out_w, uint8_t *d_canvas, size_t canvas_pitch, int row_begin, int row_end, cudaStream_t st);
int srf_check_fuse_args(srf_ctx *ctx, const srf_tile *tiles, int n_tiles, int C, int levels, const srf_weight_lut *luts, int n_luts,
                        int out_h, int out_w, const void *canvas, size_t canvas_pitch, int row_begin, int row_end);
void srf_fuse_input_extent(const srf_tile &t, int levels, int out_h, int out_w, int row_begin, int row_end, int *rows, int *cols, int *row0 = nullptr);
// enqueue only: the sums land in slot `slot` (0..63) of ctx->pinned once the stream is synchronised; srf_iqa_collect
// converts them
int srf_iqa_enqueue(srf_ctx *ctx, const uint8_t *a, size_t pa, const uint8_t *b, size_t pb, int h, int w, int C, int scale_a, int scale_b,
                    double data_range, int row_begin, int row_end, cudaStream_t st, int slot, double *d_sums = nullptr, int window = 0);
void srf_iqa_collect(srf_ctx *ctx, int w, int C, int row_begin, int row_end, srf_iqa_sums *out, int slot);
