// srf_comm.cu -- the multi-GPU exchange of the additive PSNR / SSIM sums over NVLink peer memory (SURVEY.md section 8e:
// "partial SSIM/MSE sums are all-reduced"; the reference itself has no collectives -- it is a single-process CPU program).
//
// One process per GPU.  Every rank owns an INBOX in its device memory and opens every peer's inbox through CUDA IPC
// (cudaIpcGetMemHandle / cudaIpcOpenMemHandle; the 64-byte handles travel over whatever the host side already has --
// torch.distributed in the Python mirror).  srf_allreduce_metrics then is ONE single-CTA kernel per rank:
//   1. store my n doubles into slot [my_rank][parity] of EVERY rank's inbox (plain st.global through the peer mapping --
//      NVLink writes), fence, then release-store the exchange number into that slot's flag;
//   2. spin (acquire loads, bounded) until all `world` flags of MY inbox carry this exchange number;
//   3. sum the `world` slots in rank order -- the same order on every rank, so all ranks hold bit-identical totals --
//      into d_sums (in place) and, when asked, into pinned host memory the caller reads after a stream synchronise.
// Two parity slots suffice: a rank can only start exchange k + 2 after every peer has written exchange k + 1, which
// a peer does only after it has finished reading exchange k.
// The kernels of different ranks wait on one another, which is safe only because each runs on its own GPU (never launch
// two ranks' exchanges on one device).  No NCCL call, no host round trip per exchange beyond the final read.
#include "srf_common.cuh"

#include <string.h>

namespace {

struct Inbox {
    double data[SRF_COMM_MAX_RANKS][2][SRF_COMM_MAX_DOUBLES];
    unsigned long long flag[SRF_COMM_MAX_RANKS][2];
};

struct CommPeers {
    Inbox *inbox[SRF_COMM_MAX_RANKS]; // inbox[r]: rank r's inbox as mapped into THIS process (inbox[rank] is the local one)
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(SRF_COMM_MAX_DOUBLES) k_sums_exchange(double *__restrict__ sums, int n, CommPeers peers, int rank, int world,
                                                                        unsigned long long seq, double *__restrict__ host_out, int *__restrict__ status)
{
    pdl_enter(); // the producer of `sums` (k_iqa_reduce) has completed and flushed once this returns
    const int t = threadIdx.x;
    const int par = (int)(seq & 1ull);
    const double mine = (t < n) ? sums[t] : 0.0;
    for (int r = 0; r < world; ++r)
        if (t < n) peers.inbox[r]->data[rank][par][t] = mine;
    __threadfence_system();
    __syncthreads();
    if (t < world) st_release_sys(&peers.inbox[t]->flag[rank][par], seq);
    bool ok = true;
    if (t < world) {
        const unsigned long long *f = &peers.inbox[rank]->flag[t][par];
        unsigned spin = 0;
        while (ld_acquire_sys(f) < seq) {
            if (++spin > (1u << 26)) { // ~ seconds: a peer never arrived -- report instead of hanging the GPU
                ok = false;
                break;
            }
            __nanosleep(32);
        }
    }
    ok = __syncthreads_and(ok);
    if (t == 0 && status) *status = ok ? 0 : 1;
    if (t < n) {
        double tot = 0.0;
        if (ok) {
            const Inbox *me = peers.inbox[rank];
            for (int r = 0; r < world; ++r) tot += __ldcg(&me->data[r][par][t]); // rank order: identical totals everywhere
        } else {
            tot = __longlong_as_double(0x7ff8000000000000ll); // NaN: the caller's results must not look valid
        }
        sums[t] = tot;
        if (host_out) host_out[t] = tot;
    }
}

} // namespace

struct srf_comm {
    int rank = 0, world = 1;
    Inbox *local = nullptr;
    CommPeers peers;
    bool opened[SRF_COMM_MAX_RANKS] = {};
    unsigned long long seq = 0;
    int *d_status = nullptr;
    double *h_out = nullptr, *h_out_dev = nullptr; // pinned host block the kernel writes the totals to (+ its device alias)
};

void srf_comm_free(srf_ctx *ctx)
{
    srf_comm *c = ctx->comm;
    if (!c) return;
    for (int r = 0; r < SRF_COMM_MAX_RANKS; ++r)
        if (c->opened[r]) cudaIpcCloseMemHandle(c->peers.inbox[r]);
    if (c->local) cudaFree(c->local);
    if (c->d_status) cudaFree(c->d_status);
    if (c->h_out) cudaFreeHost(c->h_out);
    delete c;
    ctx->comm = nullptr;
    (void)cudaGetLastError();
}

extern "C" int srf_comm_handle(srf_ctx *ctx, void *handle64)
{
    if (!ctx) return SRF_ERR_ARG;
    SRF_REQUIRE(ctx, handle64, "srf_comm_handle: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == SRF_COMM_HANDLE_BYTES, "IPC handle size");
    SRF_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->comm) ctx->comm = new srf_comm();
    srf_comm *c = ctx->comm;
    for (int r = 0; r < SRF_COMM_MAX_RANKS; ++r) // a re-initialisation: drop the previous group's mappings
        if (c->opened[r]) {
            cudaIpcCloseMemHandle(c->peers.inbox[r]);
            c->opened[r] = false;
        }
    memset(&c->peers, 0, sizeof c->peers);
    if (!c->local) {
        SRF_CHECK_CUDA(ctx, cudaMalloc(&c->local, sizeof(Inbox)));
        SRF_CHECK_CUDA(ctx, cudaMalloc(&c->d_status, sizeof(int)));
        SRF_CHECK_CUDA(ctx, cudaHostAlloc(&c->h_out, SRF_COMM_MAX_DOUBLES * sizeof(double), cudaHostAllocMapped));
        SRF_CHECK_CUDA(ctx, cudaHostGetDevicePointer(&c->h_out_dev, c->h_out, 0));
    }
    // exchange numbers restart at 1 with every (re-)initialisation: stale flags must not satisfy the new group's waits.
    // Peers only write here after srf_comm_init, i.e. after the handle exchange that follows this call on every rank.
    SRF_CHECK_CUDA(ctx, cudaDeviceSynchronize());
    SRF_CHECK_CUDA(ctx, cudaMemset(c->local, 0, sizeof(Inbox)));
    SRF_CHECK_CUDA(ctx, cudaMemset(c->d_status, 0, sizeof(int)));
    SRF_CHECK_CUDA(ctx, cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    SRF_CHECK_CUDA(ctx, cudaIpcGetMemHandle(&h, c->local));
    memcpy(handle64, &h, sizeof h);
    return SRF_OK;
}

extern "C" int srf_comm_init(srf_ctx *ctx, int rank, int world, const void *handles)
{
    if (!ctx) return SRF_ERR_ARG;
    SRF_REQUIRE(ctx, ctx->comm && ctx->comm->local, "srf_comm_init: call srf_comm_handle first");
    SRF_REQUIRE(ctx, world >= 1 && world <= SRF_COMM_MAX_RANKS && rank >= 0 && rank < world, "srf_comm_init: bad rank %d of %d (max %d ranks)", rank,
                world, SRF_COMM_MAX_RANKS);
    SRF_REQUIRE(ctx, handles || world == 1, "srf_comm_init: null handle array");
    SRF_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    srf_comm *c = ctx->comm;
    c->rank = rank;
    c->world = world;
    c->seq = 0;
    for (int r = 0; r < world; ++r) {
        if (r == rank) {
            c->peers.inbox[r] = c->local;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const uint8_t *)handles + (size_t)r * SRF_COMM_HANDLE_BYTES, sizeof h);
        void *p = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            srf_set_error(ctx, "srf_comm_init: cannot map rank %d's inbox (%s) -- no peer access between the GPUs?", r, cudaGetErrorString(e));
            return SRF_ERR_UNSUPPORTED;
        }
        c->peers.inbox[r] = (Inbox *)p;
        c->opened[r] = true;
    }
    return SRF_OK;
}

extern "C" int srf_allreduce_metrics(srf_ctx *ctx, double *d_sums, int n, void *stream)
{
    if (!ctx) return SRF_ERR_ARG;
    SRF_REQUIRE(ctx, d_sums && n >= 1 && n <= SRF_COMM_MAX_DOUBLES, "srf_allreduce_metrics: n must be 1..%d (got %d)", SRF_COMM_MAX_DOUBLES, n);
    SRF_REQUIRE(ctx, ctx->comm && ctx->comm->peers.inbox[ctx->comm->rank], "srf_allreduce_metrics: srf_comm_init has not been called");
    SRF_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    srf_comm *c = ctx->comm;
    c->seq++;
    cudaStream_t st = (cudaStream_t)stream;
    {
        srf_launch_scope scope(ctx, "sums_exchange", 0.0, st);
        srf_launch_pdl(k_sums_exchange, dim3(1), dim3(SRF_COMM_MAX_DOUBLES), 0, st, d_sums, n, c->peers, c->rank, c->world, c->seq, c->h_out_dev,
                       c->d_status);
    }
    SRF_CHECK_CUDA(ctx, cudaGetLastError());
    return srf_note_device_work(ctx, st);
}

extern "C" int srf_comm_result(srf_ctx *ctx, double *h_out, int n, void *stream)
{
    if (!ctx) return SRF_ERR_ARG;
    SRF_REQUIRE(ctx, h_out && n >= 1 && n <= SRF_COMM_MAX_DOUBLES, "srf_comm_result: n must be 1..%d (got %d)", SRF_COMM_MAX_DOUBLES, n);
    SRF_REQUIRE(ctx, ctx->comm && ctx->comm->h_out, "srf_comm_result: srf_comm_init has not been called");
    SRF_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    SRF_CHECK_CUDA(ctx, cudaStreamSynchronize((cudaStream_t)stream));
    memcpy(h_out, ctx->comm->h_out, (size_t)n * sizeof(double));
    for (int i = 0; i < n; ++i)
        if (h_out[i] != h_out[i]) {
            srf_set_error(ctx, "srf_allreduce_metrics: a peer rank never arrived (exchange %llu timed out)", ctx->comm->seq);
            return SRF_ERR_CUDA;
        }
    return SRF_OK;
}
