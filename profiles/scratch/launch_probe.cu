// (1) host time of launching a graph of 8 kernels whose __grid_constant__ parameter is 2 / 10 / 20 KB
// (2) per-block cost of a TMA box whose descriptor comes from kernel-parameter space vs global memory (+ proxy fence)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o launch_probe launch_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <chrono>
#include <vector>
template <int N> struct Big { int x[N / 4]; };
__device__ int g_sink;
template <int N> __global__ void k_big(const __grid_constant__ Big<N> p, int i) { if (threadIdx.x == 0 && p.x[i] == 12345) g_sink = 1; }
template <int N> double graph_launch_us(int nodes)
{
    static Big<N> h; h.x[0] = 1;
    cudaStream_t st; cudaStreamCreate(&st);
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    for (int i = 0; i < nodes; ++i) k_big<N><<<148, 128, 0, st>>>(h, i);
    cudaStreamEndCapture(st, &g);
    cudaGraphInstantiate(&ge, g, 0);
    for (int i = 0; i < 5; ++i) cudaGraphLaunch(ge, st);
    cudaStreamSynchronize(st);
    double tot = 0;
    for (int r = 0; r < 200; ++r) {
        auto a = std::chrono::high_resolution_clock::now();
        cudaGraphLaunch(ge, st);
        auto b = std::chrono::high_resolution_clock::now();
        cudaStreamSynchronize(st);
        tot += std::chrono::duration<double, std::micro>(b - a).count();
    }
    // GPU-side: graph launch to completion with the GPU idle before
    double tot2 = 0;
    for (int r = 0; r < 200; ++r) {
        auto a = std::chrono::high_resolution_clock::now();
        cudaGraphLaunch(ge, st);
        cudaStreamSynchronize(st);
        auto b = std::chrono::high_resolution_clock::now();
        tot2 += std::chrono::duration<double, std::micro>(b - a).count();
    }
    printf("param %5d B x %d nodes: cudaGraphLaunch host %.1f us, launch+sync %.1f us\n", N, nodes, tot / 200, tot2 / 200);
    return tot / 200;
}
// ---- TMA descriptor source
template <int MODE> // 0 param, 1 global + fence.acquire.gpu, 2 global + fence.acquire.sys, 3 global no fence
__global__ void k_tma(const __grid_constant__ CUtensorMap pm, const CUtensorMap *gm, int reps, unsigned *out)
{
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    unsigned b = (unsigned)__cvta_generic_to_shared(&bar), d = (unsigned)__cvta_generic_to_shared(sm);
    const CUtensorMap *m = MODE == 0 ? &pm : gm;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;");
        if (MODE == 1) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(m) : "memory");
        if (MODE == 2) asm volatile("fence.proxy.tensormap::generic.acquire.sys [%0], 128;" ::"l"(m) : "memory");
        unsigned par = 0;
        for (int r = 0; r < reps; ++r) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(72u * 4 * 10 * 3));
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(d),
                         "l"(m), "r"(4 * (int)(blockIdx.x % 8)), "r"((int)(blockIdx.x % 50)), "r"(0), "r"(b) : "memory");
            unsigned ok = 0;
            while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(ok) : "r"(b), "r"(par) : "memory");
            par ^= 1;
        }
        if (blockIdx.x == 0) out[0] = ((const unsigned *)sm)[0];
    }
}
template <int MODE> void tma_time(const CUtensorMap &hm, const CUtensorMap *dm, unsigned *out, const char *name)
{
    cudaFuncSetAttribute(k_tma<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int reps : {1, 4}) {
        k_tma<MODE><<<148 * 64, 32, 16 * 1024>>>(hm, dm, reps, out);
        cudaDeviceSynchronize();
        cudaEventRecord(a);
        k_tma<MODE><<<148 * 64, 32, 16 * 1024>>>(hm, dm, reps, out);
        cudaEventRecord(b);
        cudaError_t e = cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, a, b);
        unsigned first = 0; cudaMemcpy(&first, out, 4, cudaMemcpyDeviceToHost);
        printf("%-28s reps %d: %.1f us for %d blocks (%s) first word %u\n", name, reps, ms * 1e3, 148 * 64, cudaGetErrorString(e), first);
    }
}
int main()
{
    graph_launch_us<2048>(8); graph_launch_us<10240>(8); graph_launch_us<20480>(8); graph_launch_us<20480>(16); graph_launch_us<256>(8); graph_launch_us<256>(24);
    const int W = 256, H = 64, C = 3;
    std::vector<float> h((size_t)W * H * C);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
    float *d; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    CUtensorMap m;
    cuuint64_t dims[3] = {W, H, C}, str[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    cuuint32_t box[3] = {72, 10, 3}, one[3] = {1, 1, 1};
    CUresult r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, str, box, one, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d\n", (int)r);
    CUtensorMap *dm; cudaMalloc(&dm, sizeof m); cudaMemcpy(dm, &m, sizeof m, cudaMemcpyHostToDevice);
    unsigned *out; cudaMalloc(&out, 16);
    tma_time<0>(m, dm, out, "param descriptor");
    tma_time<1>(m, dm, out, "global + fence acquire.gpu");
    tma_time<2>(m, dm, out, "global + fence acquire.sys");
    tma_time<3>(m, dm, out, "global, no fence");
    return 0;
}
