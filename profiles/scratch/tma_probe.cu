// TMA descriptor probe: which (dtype, dims, strides, box) combinations complete?  nvcc -arch=sm_100a -o tma_probe tma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
#include <stdlib.h>
__device__ int g_result[4];
__global__ void k(const __grid_constant__ CUtensorMap m, int x, int y, int z, unsigned bytes)
{
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    unsigned b = (unsigned)__cvta_generic_to_shared(&bar), d = (unsigned)__cvta_generic_to_shared(sm);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes));
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(d),
                     "l"(&m), "r"(x), "r"(y), "r"(z), "r"(b)
                     : "memory");
        int ok = 0;
        for (int spin = 0; spin < (1 << 20) && !ok; ++spin) {
            unsigned o;
            asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p;}" : "=r"(o) : "r"(b) : "memory");
            ok = o;
        }
        g_result[0] = ok;
        g_result[1] = ((const unsigned short *)sm)[0];
        g_result[2] = ((const unsigned short *)sm)[1];
    }
}
int main(int argc, char **argv)
{
    const int only = argc > 1 ? atoi(argv[1]) : -1;
    int idx = -1;
    const int W = 128, H = 128, C = 3;
    std::vector<unsigned short> h((size_t)W * H * C);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned short)(i & 0xFFFF);
    unsigned short *d;
    cudaMalloc(&d, h.size() * 2);
    cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    struct Case { const char *name; CUtensorMapDataType dt; int es; int inner_elems; int box0; int x; };
    Case cases[] = {
        {"u16 box72 x=4", CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, W, 72, 4},
        {"u16 box72 x=-4", CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, W, 72, -4},
        {"u16 box64 x=4", CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, W, 64, 4},
        {"u32 box36 x=2", CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, W / 2, 36, 2},
        {"u32 box36 x=-2", CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, W / 2, 36, -2},
        {"u32 box32 x=2", CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, W / 2, 32, 2},
        {"f32 box36 x=2", CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, W / 2, 36, 2},
        {"u32 box40 x=2", CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, W / 2, 40, 2},
        {"f32 box72 W128", CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, W / 2, 72, 2},
        {"u32 box64 x=0", CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, W / 2, 64, 0},
    };
    for (const Case &cs : cases) {
        if (++idx != only && only >= 0) continue;
        CUtensorMap m;
        cuuint64_t dims[3] = {(cuuint64_t)cs.inner_elems, H, C}, str[2] = {(cuuint64_t)W * 2, (cuuint64_t)W * H * 2};
        cuuint32_t box[3] = {(cuuint32_t)cs.box0, 10, 3}, one[3] = {1, 1, 1};
        CUresult r = cuTensorMapEncodeTiled(&m, cs.dt, 3, d, dims, str, box, one, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        int res[4] = {-1, -1, -1, -1};
        cudaMemcpyToSymbol(g_result, res, sizeof res);
        unsigned bytes = (unsigned)cs.box0 * cs.es * 10 * 3;
        k<<<1, 32, 64 * 1024>>>(m, cs.x, 3, 0, bytes);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpyFromSymbol(res, g_result, sizeof res);
        printf("%-16s encode=%d kernel=%s completed=%d first=%d,%d (bytes %u)\n", cs.name, (int)r, cudaGetErrorString(e), res[0], res[1], res[2], bytes);
        if (e != cudaSuccess) { printf("context lost\n"); return 1; }
    }
    return 0;
}
