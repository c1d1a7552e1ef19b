// srf_fuse.cu -- Laplacian-pyramid tile fusion (replaces BlendingModule.laplacian_fusion,
// /root/reference/blending_module.py:369-506) as planar-float32 kernels:
//
//   k_down_u8       tile uint8 HWC          -> G_1 (planar f32)      integer-exact 5x5 / 256, DP4A, register ring
//   k_down_f32s     G_l / W_l (l >= 1)      -> G_{l+1} / W_{l+1}     cv2.pyrDown op order, TMA-bulk staged row streaming
//   k_down_weight0  border-distance profile -> W_1                   (W_0 is never materialised; runs once per shape)
//   k_up_levels     C_l = pyrUp(C_{l+1}) + (G_l - pyrUp(G_{l+1})) * W_l   l = n-3 .. 1, one streaming pass per level
//   k_up_level_top  the same for l = n-2, with C_{n-1} = G_{n-1} * W_{n-1} formed on the fly (tiny planes)
//   k_final         canvas gather: for every canvas pixel, in tile-list order, C_0 of each covering tile is rebuilt
//                   from (C_1, G_1, the uint8 tile, W_0) and summed; divide by max(sum W_0, 1e-6), clip, truncate;
//                   TMA tensor-map staging, packed FP32 interior path
// Every kernel starts with pdl_enter() and is launched with programmatic stream serialization (srf_launch_pdl).
//
// Nothing canvas-sized is ever held in float32: the reference's two float accumulators (16 B per canvas pixel,
// read-modify-written once per tile) become registers in k_final.
//
// Data layout: every pyramid level is stored planar, one (h_l x pitch_l) float plane per channel, pitch_l = w_l
// rounded up to 4 floats so every row is 16-byte aligned.  Tiles that share (h, w, weight profile) form a group;
// a group shares one weight pyramid and one launch per level (grid.z = tile).
#include "srf_common.cuh"

#include <cuda.h> // CUtensorMap (types only; cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint)

#include <chrono>
#include <stdio.h>
#include <algorithm>
#include <type_traits>
#include <string.h>

namespace {

constexpr int kThreads = 256;
constexpr int DN_BH = 16, DN_BW = 64;               // pyrDown output block
constexpr int DN_IN_ROWS = 2 * DN_BH + 3;           // 35 input rows
constexpr int DN_IN_COLS_F32 = 2 * DN_BW + 8;       // staged f32 cols: [2*ox0-4, 2*ox0+132)
constexpr int UP_BH = 32, UP_BW = 64;               // pyrUp / final output block
constexpr int UP_SRC_ROWS = UP_BH / 2 + 3;          // 19 source rows (odd start + both halos)
constexpr int UP_SRC_COLS = 44;                     // staged source cols (11 float4)

struct LevelGeom {
    int h, w, pitch;  // pitch in floats
    size_t plane;     // floats between channel planes
    int vec_end;      // pyrDown op-order bookkeeping for the level ABOVE (this level as destination)
    int vtail;
};

struct GroupDev {
    int n_levels;            // levels actually built (<= requested)
    int cn;                  // channels (1 or 3)
    int lut_len;
    const float *lut;        // device copy of the weight profile
    LevelGeom lv[SRF_MAX_LEVELS];
    float *W[SRF_MAX_LEVELS]; // W[l], l >= 1 (single plane); W[0] unused
    // G_1 = pyrDown(uint8 tile) is an exact integer / 256 (<= 65280): when the row-outer gather runs (its kernels read it as
    // such) it is stored as that integer in uint16 planes -- half the bytes of the level every big kernel touches.  TileDev::G[1]
    // then points at uint16 data with its own pitch / plane size (in ELEMENTS); every other level stays float32.
    int g1_u16;
    int g1_pitch;
    size_t g1_plane;
};

struct TileDev {
    const uint8_t *src;
    size_t pitch;
    int h, w, y, x;
    float *G[SRF_MAX_LEVELS]; // l >= 1
    float *C[SRF_MAX_LEVELS]; // 1 <= l <= n-2
    int g0[SRF_MAX_LEVELS], g1[SRF_MAX_LEVELS]; // rows of G_l to produce
    int c0[SRF_MAX_LEVELS], c1[SRF_MAX_LEVELS]; // rows of C_l to produce
    int gx1[SRF_MAX_LEVELS], cx1[SRF_MAX_LEVELS]; // columns [0, gx1) of G_l / [0, cx1) of C_l are needed (canvas crop)
};

// ------------------------------------------------------------------------------------------------------------
// cooperative staging helpers
// ------------------------------------------------------------------------------------------------------------
// cols [c0, c0+4) of a float plane row with `w` valid floats (pitch multiple of 4, c0 multiple of 4); zeros outside
__device__ __forceinline__ float4 load_f32_chunk(const float *grow, int c0, int w)
{
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (grow != nullptr && c0 >= 0 && c0 < w) {
        if (c0 + 4 <= w) {
            v = *reinterpret_cast<const float4 *>(grow + c0);
        } else {
            v.x = grow[c0];
            if (c0 + 1 < w) v.y = grow[c0 + 1];
            if (c0 + 2 < w) v.z = grow[c0 + 2];
        }
    }
    return v;
}

__device__ __forceinline__ float lut_weight(const float *__restrict__ lut, int lut_len, int y, int x, int h, int w)
{
    int d = min(min(y, h - 1 - y), min(x, w - 1 - x));
    d = min(d, lut_len - 1);
    return __ldg(lut + d);
}

// ------------------------------------------------------------------------------------------------------------
// k_down_u8: uint8 HWC tile -> G_1 planar.  All sums are exact integers (<= 65280), so OpenCV's float op order
// is irrelevant here; the result is (int) * 2^-8, bit-identical to cv2.pyrDown(float32(tile)).
//
// One CTA = 128 output columns x 32 output rows of one tile, 128*C threads: thread (x, c) owns output column x of
// channel c and streams DOWN the input rows.  Input rows are staged 8 at a time (double buffered, one 16-pixel
// group = C 16-byte loads per staging thread) and DE-INTERLEAVED into per-channel byte planes on the way in, so the
// horizontal 5-tap is two aligned 32-bit LDS + two DP4A per value; the five most recent horizontal sums live in a
// register shift ring and every second row emits one output row (coalesced 128-byte stores per warp).
// grid = (ceil(ow/128), ceil(rows/32), tiles in group)
// ------------------------------------------------------------------------------------------------------------
#ifndef SRF_D8_CTAS
#define SRF_D8_CTAS 3 // resident CTAs per SM the register allocation of k_down_u8<3> is held to (2 at the compiler's own choice: +20 % time)
#endif
constexpr int D8_BW = 128, D8_ROWS = 32, D8_BATCH = 8;
constexpr int D8_GROUPS = 18;           // 16-pixel groups staged per row: pixels [2*ox0-16, 2*ox0+272)
constexpr int D8_PSTRIDE = D8_GROUPS * 16;

__device__ __forceinline__ uint4 load_u8_chunk(const uint8_t *grow, int o, int valid)
{
    uint4 v = make_uint4(0, 0, 0, 0);
    if (o >= 0 && o + 16 <= valid) {
        v = *reinterpret_cast<const uint4 *>(grow + o);
    } else if (o + 16 > 0 && o < valid) {
        uint8_t *b = reinterpret_cast<uint8_t *>(&v);
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int p = o + e;
            b[e] = (p >= 0 && p < valid) ? grow[p] : (uint8_t)0;
        }
    }
    return v;
}

template <int C, bool U16>
__global__ void __launch_bounds__(128 * C, (C == 3) ? SRF_D8_CTAS : 8) k_down_u8(const TileDev *__restrict__ tiles, const int *__restrict__ tile_ids, GroupDev g)
{
    pdl_enter(); // programmatic dependent launch: nothing above this line may touch global data
    __shared__ __align__(16) uint8_t s_pl[2][D8_BATCH][C][D8_PSTRIDE];
    constexpr int NT = 128 * C;

    const TileDev &t = tiles[tile_ids[blockIdx.z]];
    const int h = t.h, w = t.w;
    const LevelGeom L1 = g.lv[1];
    const int ow = L1.w;
    const int oy0 = t.g0[1] + blockIdx.y * D8_ROWS;
    const int ox0 = blockIdx.x * D8_BW;
    if (oy0 >= t.g1[1] || ox0 >= ow || ox0 >= t.gx1[1]) return;
    const int oy1 = min(oy0 + D8_ROWS, t.g1[1]);
    const int in_y0 = 2 * oy0 - 2;
    const int px_g0 = 2 * ox0 - 16;                 // first staged pixel (multiple of 16 -> 16-byte aligned bytes)
    const int n_in_rows = 2 * (oy1 - oy0) + 3;
    const int n_batches = (n_in_rows + D8_BATCH - 1) / D8_BATCH;
    const uint8_t *src = t.src;
    const size_t pitch = t.pitch;
    const int valid = w * C;

    // staging role: task = (row j of the batch, group q); 144 tasks over NT threads
    constexpr int N_TASKS = D8_BATCH * D8_GROUPS;
    constexpr int TPT = (N_TASKS + NT - 1) / NT; // tasks per thread: 1 (C = 3) or 2 (C = 1)
    uint4 pre[TPT][C];
    auto stage_load = [&](int b) {
#pragma unroll
        for (int u = 0; u < TPT; ++u) {
            const int task = threadIdx.x + u * NT;
            if (task >= N_TASKS) continue;
            const int sj = task / D8_GROUPS, sq = task % D8_GROUPS;
            const int gy = srf_reflect101(in_y0 + b * D8_BATCH + sj, h);
            const uint8_t *grow = src + (size_t)gy * pitch;
            const int o = (px_g0 + 16 * sq) * C;
#pragma unroll
            for (int k = 0; k < C; ++k) pre[u][k] = load_u8_chunk(grow, o + 16 * k, valid);
        }
    };
    auto stage_store = [&](int buf) {
#pragma unroll
        for (int u = 0; u < TPT; ++u) {
            const int task = threadIdx.x + u * NT;
            if (task >= N_TASKS) continue;
            const int sj = task / D8_GROUPS, sq = task % D8_GROUPS;
            if (C == 3) {
                const uint32_t *wv = reinterpret_cast<const uint32_t *>(&pre[u][0]);
                uint32_t o0[4], o1[4], o2[4];
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    const uint32_t A = wv[3 * m], B = wv[3 * m + 1], Cw = wv[3 * m + 2];
                    o0[m] = __byte_perm(__byte_perm(A, B, 0x0630), Cw, 0x5210);
                    o1[m] = __byte_perm(__byte_perm(A, B, 0x0741), Cw, 0x6210);
                    o2[m] = __byte_perm(__byte_perm(A, B, 0x0052), Cw, 0x7410);
                }
                *reinterpret_cast<uint4 *>(&s_pl[buf][sj][0][16 * sq]) = make_uint4(o0[0], o0[1], o0[2], o0[3]);
                *reinterpret_cast<uint4 *>(&s_pl[buf][sj][C > 1 ? 1 : 0][16 * sq]) = make_uint4(o1[0], o1[1], o1[2], o1[3]);
                *reinterpret_cast<uint4 *>(&s_pl[buf][sj][C > 2 ? 2 : 0][16 * sq]) = make_uint4(o2[0], o2[1], o2[2], o2[3]);
            } else {
                *reinterpret_cast<uint4 *>(&s_pl[buf][sj][0][16 * sq]) = pre[u][0];
            }
        }
    };

    // compute role: output column x of channel c
    const int x = threadIdx.x % D8_BW, c = threadIdx.x / D8_BW;
    const int ox = ox0 + x;
    const bool x_edge = (2 * ox0 - 2 < 0) || (2 * (ox0 + D8_BW - 1) + 2 >= w); // some tap of this strip is reflected
    const int pb = 14 + 2 * x;                      // planar index of tap 0 (pixel 2*ox - 2)
    const int wo = (pb >> 2) * 4, sh = pb & 3;      // aligned word offset; sh is 0 or 2
    const uint32_t wA = sh ? 0x04010000u : 0x04060401u, wB = sh ? 0x00010406u : 0x00000001u;
    int ti[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        int q = srf_reflect101(2 * min(ox, ow - 1) + k - 2, w) - px_g0;
        ti[k] = max(min(q, D8_PSTRIDE - 1), 0);
    }
    // output rows of a batch leave through a running pointer: row oy0 + 4 b - 2 + (j / 2) at even j of batch b
    const bool col_ok = ox < ow;
    const ptrdiff_t row_stride = U16 ? (ptrdiff_t)g.g1_pitch : (ptrdiff_t)L1.pitch;
    float *po32 = t.G[1] + (size_t)c * L1.plane + ((ptrdiff_t)oy0 - 2) * row_stride + ox;
    unsigned short *po16 = reinterpret_cast<unsigned short *>(t.G[1]) + (size_t)c * g.g1_plane + ((ptrdiff_t)oy0 - 2) * row_stride + ox;
    int oyb = oy0 - 2;

    int h0 = 0, h1 = 0, h2 = 0, h3 = 0, h4 = 0;
    stage_load(0);
    stage_store(0);
    __syncthreads();
    for (int b = 0; b < n_batches; ++b) {
        const int buf = b & 1;
        if (b + 1 < n_batches) stage_load(b + 1);
#pragma unroll
        for (int j = 0; j < D8_BATCH; ++j) {
            const uint8_t *row = &s_pl[buf][j][c][0];
            int hv;
            if (!x_edge) {
                const uint32_t w0 = *reinterpret_cast<const uint32_t *>(row + wo);
                const uint32_t w1 = *reinterpret_cast<const uint32_t *>(row + wo + 4);
                hv = (int)__dp4a(w0, wA, __dp4a(w1, wB, 0u));
            } else {
                hv = (int)row[ti[0]] + (int)row[ti[4]] + 4 * ((int)row[ti[1]] + (int)row[ti[3]]) + 6 * (int)row[ti[2]];
            }
            h0 = h1; h1 = h2; h2 = h3; h3 = h4; h4 = hv;
            if ((j & 1) == 0) { // stream position r = 8 b + j holds input rows r-4 .. r in the ring: output row oy0 + (r - 4) / 2
                const int oy = oyb + (j >> 1);
                if (oy >= oy0 && oy < oy1 && col_ok) {
                    const int v = h0 + h4 + 4 * (h1 + h3) + 6 * h2;
                    if (U16)
                        po16[(j >> 1) * row_stride] = (unsigned short)v; // G_1 * 256, exact
                    else
                        po32[(j >> 1) * row_stride] = __fmul_rn((float)v, 1.f / 256.f);
                }
            }
        }
        oyb += D8_BATCH / 2;
        po16 += (D8_BATCH / 2) * row_stride;
        po32 += (D8_BATCH / 2) * row_stride;
        if (b + 1 < n_batches) stage_store(buf ^ 1);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------------------
// k_down_weight0: pyrDown of the level-0 weight map into W_1 (cv2.pyrDown op order incl. the SIMD-body / scalar
// split), with the source generated on the fly from the border-distance profile (l == 0).  Block form (16 x 64
// outputs per CTA, 2-pass through smem); it runs once per tile shape -- the result is cached across calls -- so it
// is kept simple.  (It still understands a plane source and tile planes; only the profile source is launched.)
// ------------------------------------------------------------------------------------------------------------
template <bool WEIGHT>
__global__ void __launch_bounds__(kThreads) k_down_weight0(const TileDev *__restrict__ tiles, const int *__restrict__ tile_ids, GroupDev g,
                                                       int l /* source level */)
{
    pdl_enter(); // programmatic dependent launch: nothing above this line may touch global data
    __shared__ __align__(16) float s_in[DN_IN_ROWS][DN_IN_COLS_F32];
    __shared__ float s_h[DN_IN_ROWS][DN_BW];

    const LevelGeom Ls = g.lv[l], Ld = g.lv[l + 1];
    const int h = Ls.h, w = Ls.w, ow = Ld.w;
    const float *src = nullptr;
    float *dst;
    int r0, r1, c = 0, cn_ref, xlim = ow;
    if (WEIGHT) {
        if (l > 0) src = g.W[l];
        dst = g.W[l + 1];
        r0 = 0;
        r1 = Ld.h;
        cn_ref = 1;
    } else {
        const int z = blockIdx.z;
        const TileDev &t = tiles[tile_ids[z / g.cn]];
        c = z % g.cn;
        src = t.G[l] + (size_t)c * Ls.plane;
        dst = t.G[l + 1] + (size_t)c * Ld.plane;
        r0 = t.g0[l + 1];
        r1 = t.g1[l + 1];
        cn_ref = g.cn;
        xlim = t.gx1[l + 1];
    }
    const int oy0 = r0 + blockIdx.y * DN_BH;
    const int ox0 = blockIdx.x * DN_BW;
    if (oy0 >= r1 || ox0 >= ow || ox0 >= xlim) return;
    const int oy1 = min(oy0 + DN_BH, r1);
    const int in_y0 = 2 * oy0 - 2;
    const int in_x0 = 2 * ox0 - 4; // multiple of 4
    const int n_in_rows = 2 * (oy1 - oy0) + 3;

    for (int i = threadIdx.x; i < n_in_rows * (DN_IN_COLS_F32 / 4); i += kThreads) {
        const int r = i / (DN_IN_COLS_F32 / 4), ch = i - r * (DN_IN_COLS_F32 / 4);
        const int gy = in_y0 + r;
        const int c0 = in_x0 + ch * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gy >= 0 && gy < h) {
            if (WEIGHT && l == 0) {
                float *e = reinterpret_cast<float *>(&v);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (c0 + k >= 0 && c0 + k < w) e[k] = lut_weight(g.lut, g.lut_len, gy, c0 + k, h, w);
            } else {
                v = load_f32_chunk(src + (size_t)gy * Ls.pitch, c0, w);
            }
        }
        *reinterpret_cast<float4 *>(&s_in[r][ch * 4]) = v;
    }
    __syncthreads();

    const bool x_interior = (ox0 > 0) && (2 * (ox0 + DN_BW) + 2 < w);
    const int vec_end = Ld.vec_end;
    for (int i = threadIdx.x; i < n_in_rows * DN_BW; i += kThreads) {
        const int x = i % DN_BW, r = i / DN_BW;
        const int gy = in_y0 + r;
        const int ox = ox0 + x;
        if (gy < 0 || gy >= h || ox >= ow) continue;
        const float *row = &s_in[r][0];
        float a0, a1, a2, a3, a4;
        if (x_interior) {
            const float2 p0 = *reinterpret_cast<const float2 *>(row + 2 * x + 2);
            const float2 p1 = *reinterpret_cast<const float2 *>(row + 2 * x + 4);
            a0 = p0.x; a1 = p0.y; a2 = p1.x; a3 = p1.y; a4 = row[2 * x + 6];
        } else {
            const int gx = 2 * ox;
            a0 = row[srf_reflect101(gx - 2, w) - in_x0];
            a1 = row[srf_reflect101(gx - 1, w) - in_x0];
            a2 = row[srf_reflect101(gx, w) - in_x0];
            a3 = row[srf_reflect101(gx + 1, w) - in_x0];
            a4 = row[srf_reflect101(gx + 2, w) - in_x0];
        }
        const bool sca = (ox < 1) || (ox >= vec_end);
        s_h[r][x] = sca ? dn_h_sca(a0, a1, a2, a3, a4) : dn_h_vec(a0, a1, a2, a3, a4);
    }
    __syncthreads();

    const int n_out_rows = oy1 - oy0;
    const int vtail = Ld.vtail;
    for (int i = threadIdx.x; i < n_out_rows * DN_BW; i += kThreads) {
        const int x = i % DN_BW, ry = i / DN_BW;
        const int ox = ox0 + x;
        if (ox >= ow) continue;
        const int oy = oy0 + ry;
        float rv[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) rv[k] = s_h[srf_reflect101(2 * oy + k - 2, h) - in_y0][x];
        const bool sca = (ox * cn_ref + c) >= vtail;
        dst[(size_t)oy * Ld.pitch + ox] = sca ? dn_v_sca(rv[0], rv[1], rv[2], rv[3], rv[4]) : dn_v_vec(rv[0], rv[1], rv[2], rv[3], rv[4]);
    }
}

// ------------------------------------------------------------------------------------------------------------
// pyrUp building blocks of k_up_level_top.
// A "source block" is rows [jb, jb + nrows) x cols [ib, ib + UP_SRC_COLS) of a level-(l+1) plane staged in smem.
// ------------------------------------------------------------------------------------------------------------
struct SrcWin {
    int jb, nrows; // first staged source row, count
    int ib;        // first staged source col (multiple of 4)
};

// rows/cols of level l+1 needed to up-sample output rows [ty0, ty1) x cols [tx0, tx1) of level l
__device__ __forceinline__ SrcWin src_window(int ty0, int ty1, int tx0, int h1)
{
    SrcWin s;
    const int jlo = max((ty0 >> 1) - 1, 0);
    const int jhi = min(((ty1 - 1) >> 1) + 1, h1 - 1);
    s.jb = jlo;
    s.nrows = jhi - jlo + 1;
    s.ib = (max((tx0 >> 1) - 1, 0)) & ~3;
    return s;
}

// ------------------------------------------------------------------------------------------------------------
// k_up_level_top: C_l = pyrUp(C_{l+1}) + (G_l - pyrUp(G_{l+1})) * W_l for l = n-2, where C_{l+1} is not stored: it is
// G_{l+1} * W_{l+1} (the coarsest Laplacian level is the Gaussian one).  Block form; the planes at this level are tiny
// (128 x 128 at 7 levels of a 4096 tile).  TOP = false (stored C) is served by the streaming k_up_levels below.
// grid = (ceil(w_l/64), ceil(rows/32), tile * C + c)
// ------------------------------------------------------------------------------------------------------------
template <bool TOP>
__global__ void __launch_bounds__(kThreads) k_up_level_top(const TileDev *__restrict__ tiles, const int *__restrict__ tile_ids, GroupDev g, int l)
{
    pdl_enter(); // programmatic dependent launch: nothing above this line may touch global data
    __shared__ __align__(16) float s_c[UP_SRC_ROWS][UP_SRC_COLS];
    __shared__ __align__(16) float s_g[UP_SRC_ROWS][UP_SRC_COLS];
    __shared__ __align__(16) float s_hc[UP_SRC_ROWS][UP_BW];
    __shared__ __align__(16) float s_hg[UP_SRC_ROWS][UP_BW];

    const int z = blockIdx.z;
    const TileDev &t = tiles[tile_ids[z / g.cn]];
    const int c = z % g.cn;
    const LevelGeom L0 = g.lv[l], L1 = g.lv[l + 1];
    const int ty0 = t.c0[l] + blockIdx.y * UP_BH;
    const int tx0 = blockIdx.x * UP_BW;
    if (ty0 >= t.c1[l] || tx0 >= L0.w || tx0 >= t.cx1[l]) return;
    const int ty1 = min(ty0 + UP_BH, t.c1[l]);
    const int tx1 = min(tx0 + UP_BW, L0.w);
    const SrcWin sw = src_window(ty0, ty1, tx0, L1.h);
    const float *G1 = t.G[l + 1] + (size_t)c * L1.plane;
    const float *C1 = TOP ? nullptr : t.C[l + 1] + (size_t)c * L1.plane;
    const float *W1 = TOP ? g.W[l + 1] : nullptr;

    for (int i = threadIdx.x; i < sw.nrows * (UP_SRC_COLS / 4); i += kThreads) {
        const int r = i / (UP_SRC_COLS / 4), ch = i - r * (UP_SRC_COLS / 4);
        const int gy = sw.jb + r, c0 = sw.ib + ch * 4;
        const float4 gv = load_f32_chunk(G1 + (size_t)gy * L1.pitch, c0, L1.w);
        float4 cv;
        if (TOP) {
            const float4 wv = load_f32_chunk(W1 + (size_t)gy * L1.pitch, c0, L1.w);
            cv = make_float4(__fmul_rn(gv.x, wv.x), __fmul_rn(gv.y, wv.y), __fmul_rn(gv.z, wv.z), __fmul_rn(gv.w, wv.w));
        } else {
            cv = load_f32_chunk(C1 + (size_t)gy * L1.pitch, c0, L1.w);
        }
        *reinterpret_cast<float4 *>(&s_g[r][ch * 4]) = gv;
        *reinterpret_cast<float4 *>(&s_c[r][ch * 4]) = cv;
    }
    __syncthreads();

    for (int i = threadIdx.x; i < sw.nrows * UP_BW; i += kThreads) {
        const int x = i % UP_BW, r = i / UP_BW;
        const int tx = tx0 + x;
        if (tx >= tx1) continue;
        const float *rc = &s_c[r][0] - sw.ib;
        const float *rg = &s_g[r][0] - sw.ib;
        s_hc[r][x] = up_h_sample(tx, L1.w, [&](int ii) { return rc[ii]; });
        s_hg[r][x] = up_h_sample(tx, L1.w, [&](int ii) { return rg[ii]; });
    }
    __syncthreads();

    const float *G0 = t.G[l] + (size_t)c * L0.plane;
    const float *W0 = g.W[l];
    float *C0 = t.C[l] + (size_t)c * L0.plane;
    const int n_rows = ty1 - ty0;
    for (int i = threadIdx.x; i < n_rows * (UP_BW / 4); i += kThreads) {
        const int xg = i % (UP_BW / 4), ry = i / (UP_BW / 4);
        const int tx = tx0 + xg * 4;
        if (tx >= tx1) continue;
        const int ty = ty0 + ry;
        const int j = ty >> 1;
        const int rj = j - sw.jb;
        const int rm = srf_reflect101(j - 1, L1.h) - sw.jb;
        const int rn = min(j + 1, L1.h - 1) - sw.jb;
        const float4 pc = *reinterpret_cast<const float4 *>(&s_hc[rj][xg * 4]);
        const float4 nc = *reinterpret_cast<const float4 *>(&s_hc[rn][xg * 4]);
        const float4 pg = *reinterpret_cast<const float4 *>(&s_hg[rj][xg * 4]);
        const float4 ng = *reinterpret_cast<const float4 *>(&s_hg[rn][xg * 4]);
        float uc[4], ug[4];
        if (ty & 1) {
            uc[0] = up_odd(pc.x, nc.x); uc[1] = up_odd(pc.y, nc.y); uc[2] = up_odd(pc.z, nc.z); uc[3] = up_odd(pc.w, nc.w);
            ug[0] = up_odd(pg.x, ng.x); ug[1] = up_odd(pg.y, ng.y); ug[2] = up_odd(pg.z, ng.z); ug[3] = up_odd(pg.w, ng.w);
        } else {
            const float4 mc = *reinterpret_cast<const float4 *>(&s_hc[rm][xg * 4]);
            const float4 mg = *reinterpret_cast<const float4 *>(&s_hg[rm][xg * 4]);
            uc[0] = up_even(mc.x, pc.x, nc.x); uc[1] = up_even(mc.y, pc.y, nc.y); uc[2] = up_even(mc.z, pc.z, nc.z); uc[3] = up_even(mc.w, pc.w, nc.w);
            ug[0] = up_even(mg.x, pg.x, ng.x); ug[1] = up_even(mg.y, pg.y, ng.y); ug[2] = up_even(mg.z, pg.z, ng.z); ug[3] = up_even(mg.w, pg.w, ng.w);
        }
        const size_t off = (size_t)ty * L0.pitch + tx;
        float4 g0;
        if (l == 1 && g.g1_u16) { // the uint16 level: four values (x 256) in one aligned 8-byte word (tx is a multiple of 4)
            const unsigned short *row = reinterpret_cast<const unsigned short *>(t.G[1]) + (size_t)c * g.g1_plane + (size_t)ty * g.g1_pitch;
            const uint2 v = *reinterpret_cast<const uint2 *>(row + tx);
            g0 = make_float4(__fmul_rn((float)(v.x & 0xFFFFu), 1.f / 256.f), __fmul_rn((float)(v.x >> 16), 1.f / 256.f),
                             __fmul_rn((float)(v.y & 0xFFFFu), 1.f / 256.f), __fmul_rn((float)(v.y >> 16), 1.f / 256.f));
        } else {
            g0 = load_f32_chunk(G0 + (size_t)ty * L0.pitch, tx, L0.w);
        }
        const float4 w0 = load_f32_chunk(W0 + (size_t)ty * L0.pitch, tx, L0.w);
        const float gg[4] = {g0.x, g0.y, g0.z, g0.w}, ww[4] = {w0.x, w0.y, w0.z, w0.w};
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float upc = __fmul_rn(uc[k], 1.f / 64.f), upg = __fmul_rn(ug[k], 1.f / 64.f);
            o[k] = __fadd_rn(upc, __fmul_rn(__fsub_rn(gg[k], upg), ww[k]));
        }
        if (tx + 4 <= L0.w) {
            *reinterpret_cast<float4 *>(C0 + off) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
            for (int k = 0; k < 4 && tx + k < L0.w; ++k) C0[off + k] = o[k];
        }
    }
}


// ------------------------------------------------------------------------------------------------------------
// k_down_f32s: pyrDown of one planar f32 level into the next, for image planes G_l and weight planes W_l (l >= 1).  One CTA = 128 output columns x
// 32 output rows of one plane, 128 threads: thread x owns output column x and streams DOWN the input rows.  Input
// rows arrive 8 at a time through cp.async into a double-buffered smem ring; the horizontal 5-tap reads 5 adjacent
// floats (two LDS.64 + one LDS.32, conflict free), the five most recent horizontal sums live in a register shift
// ring, every second row emits one output row (128-byte coalesced stores per warp).  OpenCV's op-order split (SIMD
// body / scalar border + tail columns) is reproduced per column.
// grid = (ceil(ow/128), ceil(rows/32), planes)
// ------------------------------------------------------------------------------------------------------------
constexpr int DF_BW = 128, DF_ROWS = 32, DF_BATCH = 8;
constexpr int DF_COLS = 2 * DF_BW + 8; // staged floats per row: columns [2*ox0 - 4, 2*ox0 + 260)

template <bool WEIGHT>
__global__ void __launch_bounds__(DF_BW) k_down_f32s(const TileDev *__restrict__ tiles, const int *__restrict__ tile_ids, GroupDev g, int l,
                                                     int rows_per_cta)
{
    pdl_enter(); // programmatic dependent launch: nothing above this line may touch global data
    __shared__ __align__(128) float s_in[2][DF_BATCH][DF_COLS];
    __shared__ __align__(8) uint64_t s_bar[2];

    const LevelGeom Ls = g.lv[l], Ld = g.lv[l + 1];
    const int h = Ls.h, w = Ls.w, ow = Ld.w;
    const float *src;
    float *dst;
    int r0, r1, c = 0, cn_ref, xlim = ow;
    if (WEIGHT) {
        src = g.W[l];
        dst = g.W[l + 1];
        r0 = 0;
        r1 = Ld.h;
        cn_ref = 1;
    } else {
        const int z = blockIdx.z;
        const TileDev &t = tiles[tile_ids[z / g.cn]];
        c = z % g.cn;
        src = t.G[l] + (size_t)c * Ls.plane;
        dst = t.G[l + 1] + (size_t)c * Ld.plane;
        r0 = t.g0[l + 1];
        r1 = t.g1[l + 1];
        cn_ref = g.cn;
        xlim = t.gx1[l + 1];
    }
    const int oy0 = r0 + blockIdx.y * rows_per_cta;
    const int ox0 = blockIdx.x * DF_BW;
    if (oy0 >= r1 || ox0 >= ow || ox0 >= xlim) return;
    const int oy1 = min(oy0 + rows_per_cta, r1);
    const int in_y0 = 2 * oy0 - 2;
    const int in_x0 = 2 * ox0 - 4; // multiple of 4
    const int n_in_rows = 2 * (oy1 - oy0) + 3;
    const int n_batches = (n_in_rows + DF_BATCH - 1) / DF_BATCH;
    const int pitch = Ls.pitch;

    // Interior strips (the whole 1056-byte row segment lies inside the plane row) are staged by TMA bulk copies: ONE
    // thread issues 8 cp.async.bulk per batch, completion on an mbarrier.  Edge strips keep the per-chunk cp.async path
    // below (their out-of-row chunks must be zero-filled).
    const bool use_tma = (in_x0 >= 0) && (in_x0 + DF_COLS <= pitch);
    if (use_tma) {
        if (threadIdx.x == 0) {
            mbar_init(&s_bar[0], 1);
            mbar_init(&s_bar[1], 1);
            mbar_fence_init();
        }
        __syncthreads();
    }
    auto stage_tma = [&](int b) {
        if (threadIdx.x == 0) {
            uint64_t *bar = &s_bar[b & 1];
            mbar_expect_tx(bar, DF_BATCH * DF_COLS * (unsigned)sizeof(float));
#pragma unroll
            for (int j = 0; j < DF_BATCH; ++j) {
                const int gy = srf_reflect101(in_y0 + b * DF_BATCH + j, h);
                tma_bulk_g2s(&s_in[b & 1][j][0], src + (size_t)gy * pitch + in_x0, DF_COLS * (unsigned)sizeof(float), bar);
            }
        }
    };

    // staging: lanes 0..65 of the CTA own one 16-byte chunk column each; the row loop is uniform
    const int sch = threadIdx.x;
    const int sc0 = in_x0 + sch * 4;
    const bool s_on = sch < DF_COLS / 4;
    const bool s_ok = s_on && sc0 >= 0 && sc0 + 4 <= pitch;
    auto stage = [&](int b) {
        float(*buf)[DF_COLS] = s_in[b & 1];
        if (s_on) {
#pragma unroll
            for (int j = 0; j < DF_BATCH; ++j) {
                const int gy = srf_reflect101(in_y0 + b * DF_BATCH + j, h); // uniform
                float *d = &buf[j][sch * 4];
                if (s_ok)
                    cp_async_16(d, src + (size_t)gy * pitch + sc0);
                else
                    *reinterpret_cast<float4 *>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        cp_async_commit();
    };

    const int x = threadIdx.x;
    const int ox = ox0 + x;
    const bool x_edge = (ox0 == 0) || (2 * (ox0 + DF_BW - 1) + 2 >= w); // some tap of this strip is reflected
    int ti[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const int q = srf_reflect101(2 * min(ox, ow - 1) + k - 2, w) - in_x0;
        ti[k] = max(min(q, DF_COLS - 1), 0);
    }
    const bool hsca = (ox < 1) || (ox >= Ld.vec_end);
    const bool vsca = (ox * cn_ref + c) >= Ld.vtail;

    float h0 = 0.f, h1 = 0.f, h2 = 0.f, h3 = 0.f, h4 = 0.f;
    if (use_tma) {
        stage_tma(0);
    } else {
        stage(0);
        cp_async_wait<0>();
        __syncthreads();
    }
    for (int b = 0; b < n_batches; ++b) {
        if (b + 1 < n_batches) {
            if (use_tma)
                stage_tma(b + 1);
            else
                stage(b + 1);
        }
        if (use_tma) mbar_wait(&s_bar[b & 1], (b >> 1) & 1); // batch b has landed
        const float(*buf)[DF_COLS] = s_in[b & 1];
#pragma unroll
        for (int j = 0; j < DF_BATCH; ++j) {
            const float *row = buf[j];
            float a0, a1, a2, a3, a4;
            if (!x_edge) {
                const float2 p0 = *reinterpret_cast<const float2 *>(row + 2 * x + 2);
                const float2 p1 = *reinterpret_cast<const float2 *>(row + 2 * x + 4);
                a0 = p0.x; a1 = p0.y; a2 = p1.x; a3 = p1.y; a4 = row[2 * x + 6];
            } else {
                a0 = row[ti[0]]; a1 = row[ti[1]]; a2 = row[ti[2]]; a3 = row[ti[3]]; a4 = row[ti[4]];
            }
            const float hv = hsca ? dn_h_sca(a0, a1, a2, a3, a4) : dn_h_vec(a0, a1, a2, a3, a4);
            h0 = h1; h1 = h2; h2 = h3; h3 = h4; h4 = hv;
            const int r = b * DF_BATCH + j; // stream position; rows r-4..r are in the ring
            if ((j & 1) == 0 && r >= 4) {
                const int oy = oy0 + ((r - 4) >> 1);
                if (oy < oy1 && ox < ow) dst[(size_t)oy * Ld.pitch + ox] = vsca ? dn_v_sca(h0, h1, h2, h3, h4) : dn_v_vec(h0, h1, h2, h3, h4);
            }
        }
        if (!use_tma) cp_async_wait<0>();
        __syncthreads(); // everyone is done with buffer b & 1 before batch b + 2 is copied over it
    }
}

// ------------------------------------------------------------------------------------------------------------
// k_down_u16: G_1 (uint16 planes, value * 256) -> G_2 (float32).  Everything up to G_2 is EXACT in float32: G_1 * 256 <= 65280,
// a horizontal 5-tap sum <= 16 * 65280 < 2^20 and the full 5 x 5 sum <= 256 * 65280 < 2^24 are integers, so cv2.pyrDown's
// float32 result is that integer * 2^-16 whatever its op order -- the level is computed in integer arithmetic (two DP2A per
// horizontal tap row) from half the bytes.  Same streaming structure as k_down_f32s: thread = output column, rows staged 8
// at a time by TMA bulk copies (cp.async-free edge strips are zero-filled by hand), 5-deep register ring, one output row
// per two input rows.  grid = (ceil(ow/128), ceil(rows/rows_per_cta), planes)
// ------------------------------------------------------------------------------------------------------------
constexpr int D16_COLS = 2 * DF_BW + 16; // staged uint16 per row: columns [2*ox0 - 8, 2*ox0 + 264): 544 bytes, 16-byte aligned start

__global__ void __launch_bounds__(DF_BW) k_down_u16(const TileDev *__restrict__ tiles, const int *__restrict__ tile_ids, GroupDev g, int rows_per_cta)
{
    pdl_enter(); // programmatic dependent launch: nothing above this line may touch global data
    __shared__ __align__(128) unsigned short s_in[2][DF_BATCH][D16_COLS];
    __shared__ __align__(8) uint64_t s_bar[2];

    const LevelGeom Ld = g.lv[2];
    const int h = g.lv[1].h, w = g.lv[1].w, ow = Ld.w;
    const int z = blockIdx.z;
    const TileDev &t = tiles[tile_ids[z / g.cn]];
    const int c = z % g.cn;
    const unsigned short *src = reinterpret_cast<const unsigned short *>(t.G[1]) + (size_t)c * g.g1_plane;
    float *dst = t.G[2] + (size_t)c * Ld.plane;
    const int r0 = t.g0[2], r1 = t.g1[2], xlim = t.gx1[2];
    const int oy0 = r0 + blockIdx.y * rows_per_cta;
    const int ox0 = blockIdx.x * DF_BW;
    if (oy0 >= r1 || ox0 >= ow || ox0 >= xlim) return;
    const int oy1 = min(oy0 + rows_per_cta, r1);
    const int in_y0 = 2 * oy0 - 2;
    const int in_x0 = 2 * ox0 - 8; // multiple of 8 elements = 16 bytes
    const int n_in_rows = 2 * (oy1 - oy0) + 3;
    const int n_batches = (n_in_rows + DF_BATCH - 1) / DF_BATCH;
    const int pitch = g.g1_pitch;

    const bool use_tma = (in_x0 >= 0) && (in_x0 + D16_COLS <= pitch);
    if (use_tma) {
        if (threadIdx.x == 0) {
            mbar_init(&s_bar[0], 1);
            mbar_init(&s_bar[1], 1);
            mbar_fence_init();
        }
        __syncthreads();
    }
    auto stage_tma = [&](int b) {
        if (threadIdx.x == 0) {
            uint64_t *bar = &s_bar[b & 1];
            mbar_expect_tx(bar, DF_BATCH * D16_COLS * (unsigned)sizeof(unsigned short));
#pragma unroll
            for (int j = 0; j < DF_BATCH; ++j) {
                const int gy = srf_reflect101(in_y0 + b * DF_BATCH + j, h);
                tma_bulk_g2s(&s_in[b & 1][j][0], src + (size_t)gy * pitch + in_x0, D16_COLS * (unsigned)sizeof(unsigned short), bar);
            }
        }
    };
    // edge strips: 34 lanes own one 16-byte chunk (8 elements) each; chunks outside the padded row are zero
    const int sch = threadIdx.x;
    const int sc0 = in_x0 + sch * 8;
    const bool s_on = sch < D16_COLS / 8;
    const bool s_ok = s_on && sc0 >= 0 && sc0 + 8 <= pitch;
    auto stage = [&](int b) {
        unsigned short(*buf)[D16_COLS] = s_in[b & 1];
        if (s_on) {
#pragma unroll
            for (int j = 0; j < DF_BATCH; ++j) {
                const int gy = srf_reflect101(in_y0 + b * DF_BATCH + j, h); // uniform
                unsigned short *d = &buf[j][sch * 8];
                if (s_ok)
                    cp_async_16(d, src + (size_t)gy * pitch + sc0);
                else
                    *reinterpret_cast<uint4 *>(d) = make_uint4(0, 0, 0, 0);
            }
        }
        cp_async_commit();
    };

    const int x = threadIdx.x;
    const int ox = ox0 + x;
    const bool x_edge = (ox0 == 0) || (2 * (ox0 + DF_BW - 1) + 2 >= w); // some tap of this strip is reflected
    int ti[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const int q = srf_reflect101(2 * min(ox, ow - 1) + k - 2, w) - in_x0;
        ti[k] = max(min(q, D16_COLS - 1), 0);
    }

    int h0 = 0, h1 = 0, h2 = 0, h3 = 0, h4 = 0;
    if (use_tma) {
        stage_tma(0);
    } else {
        stage(0);
        cp_async_wait<0>();
        __syncthreads();
    }
    for (int b = 0; b < n_batches; ++b) {
        if (b + 1 < n_batches) {
            if (use_tma)
                stage_tma(b + 1);
            else
                stage(b + 1);
        }
        if (use_tma) mbar_wait(&s_bar[b & 1], (b >> 1) & 1); // batch b has landed
        const unsigned short(*buf)[D16_COLS] = s_in[b & 1];
#pragma unroll
        for (int j = 0; j < DF_BATCH; ++j) {
            const unsigned short *row = buf[j];
            int hv;
            if (!x_edge) { // taps at staged columns 2x+6 .. 2x+10: two aligned 32-bit words + one halfword
                const unsigned w0 = *reinterpret_cast<const unsigned *>(row + 2 * x + 6);
                const unsigned w1 = *reinterpret_cast<const unsigned *>(row + 2 * x + 8);
                hv = (int)__dp2a_lo(w0, 0x0401u, __dp2a_lo(w1, 0x0406u, (unsigned)row[2 * x + 10]));
            } else {
                hv = (int)row[ti[0]] + (int)row[ti[4]] + 4 * ((int)row[ti[1]] + (int)row[ti[3]]) + 6 * (int)row[ti[2]];
            }
            h0 = h1; h1 = h2; h2 = h3; h3 = h4; h4 = hv;
            const int r = b * DF_BATCH + j; // stream position; rows r-4..r are in the ring
            if ((j & 1) == 0 && r >= 4) {
                const int oy = oy0 + ((r - 4) >> 1);
                if (oy < oy1 && ox < ow) {
                    const int v = h0 + h4 + 4 * (h1 + h3) + 6 * h2; // <= 256 * 65280 < 2^24: exact in float32
                    dst[(size_t)oy * Ld.pitch + ox] = __fmul_rn((float)v, 1.f / 65536.f);
                }
            }
        }
        if (!use_tma) cp_async_wait<0>();
        __syncthreads(); // everyone is done with buffer b & 1 before batch b + 2 is copied over it
    }
}

// ------------------------------------------------------------------------------------------------------------
// k_up_levels: C_l = pyrUp(C_{l+1}) + (G_l - pyrUp(G_{l+1})) * W_l with C_{l+1} stored (every level but the top one).  One CTA = 256 x 32 outputs of
// one plane, 128 threads: thread p owns output columns 2p, 2p+1 and slides down the rows with a 3-row register ring
// of horizontally up-sampled (C_{l+1}, G_{l+1}) pairs, exactly like k_final; the <= 19 x 136 source windows arrive by
// cp.async; G_l, W_l are read and C_l written as float2 (256 contiguous bytes per warp and row).
// grid = (ceil(w_l/256), ceil(rows/32), planes)
// ------------------------------------------------------------------------------------------------------------
constexpr int UL_W = 256, UL_H = 32, UL_PAIRS = UL_W / 2, UL_STEPS = UL_H / 2;
constexpr int UL_SRC_ROWS = UL_STEPS + 2; // 18
constexpr int UL_SRC_COLS = UL_PAIRS + 8; // 136

struct HPair2 {
    float a, b;
};
__device__ __forceinline__ HPair2 up_pair_even_origin(float s0, float s1, float s2)
{
    HPair2 r;
    r.a = up_even(s0, s1, s2); // column 2i, i = window centre
    r.b = up_odd(s1, s2);      // column 2i + 1
    return r;
}

template <int H, bool G0U16 = false> // H: output rows per CTA (32 amortises the 2-row source halo; 8 keeps a small level or a thin band
                                      // from being a few long-running CTAs on an idle GPU); G0U16: G_l is the uint16 level 1
__global__ void __launch_bounds__(UL_PAIRS) k_up_levels(const TileDev *__restrict__ tiles, const int *__restrict__ tile_ids, GroupDev g, int l)
{
    pdl_enter(); // programmatic dependent launch: nothing above this line may touch global data
    constexpr int STEPS = H / 2, SRC_ROWS = STEPS + 2;
    __shared__ __align__(16) float s_c[SRC_ROWS][UL_SRC_COLS];
    __shared__ __align__(16) float s_g[SRC_ROWS][UL_SRC_COLS];

    const int z = blockIdx.z;
    const TileDev &t = tiles[tile_ids[z / g.cn]];
    const int c = z % g.cn;
    const LevelGeom L0 = g.lv[l], L1 = g.lv[l + 1];
    const int h1 = L1.h, w1 = L1.w;
    const int ty0 = t.c0[l] + blockIdx.y * H;
    const int tx0 = blockIdx.x * UL_W;
    if (ty0 >= t.c1[l] || tx0 >= L0.w || tx0 >= t.cx1[l]) return;
    const int ty1 = min(ty0 + H, t.c1[l]);
    const int ypar = ty0 & 1;
    const int i0 = (tx0 >> 1) - 1;           // window start column for pair 0 (tx0 is even)
    const int ws0 = (ty0 >> 1) - (1 - ypar); // window start row for step 0
    const int sx0 = max(i0, 0) & ~3;
    const int srow0 = max(min(ws0, h1 - 1), 0);
    const int n_srows = max(min(ws0 + STEPS + 1, h1 - 1), 0) - srow0 + 1; // <= 18
    const float *G1 = t.G[l + 1] + (size_t)c * L1.plane;
    const float *C1 = t.C[l + 1] + (size_t)c * L1.plane;

    for (int i = threadIdx.x; i < n_srows * (UL_SRC_COLS / 4); i += UL_PAIRS) {
        const int r = i / (UL_SRC_COLS / 4), ch = i - r * (UL_SRC_COLS / 4);
        const int c0 = sx0 + ch * 4;
        const size_t off = (size_t)(srow0 + r) * L1.pitch + c0;
        if (c0 + 4 <= L1.pitch) {
            cp_async_16(&s_g[r][ch * 4], G1 + off);
            cp_async_16(&s_c[r][ch * 4], C1 + off);
        } else {
            *reinterpret_cast<float4 *>(&s_g[r][ch * 4]) = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4 *>(&s_c[r][ch * 4]) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    cp_async_commit();

    const int p = threadIdx.x;
    const int txa = tx0 + 2 * p, txb = txa + 1;
    const bool ina = txa < L0.w, inb = txb < L0.w;
    const bool hfast = (i0 >= 0) && (i0 + UL_PAIRS + 1 <= w1 - 1) && (w1 > 1);
    const int scol = i0 + p - sx0;
    const float *G0 = t.G[l] + (size_t)c * L0.plane;
    const unsigned short *G0h = reinterpret_cast<const unsigned short *>(t.G[l]) + (size_t)c * g.g1_plane; // G0U16 only
    const float *W0 = g.W[l];
    float *C0 = t.C[l] + (size_t)c * L0.plane;

    auto hrow = [&](int r_logical, HPair2 &oc, HPair2 &og) {
        int r = (r_logical == -1) ? min(1, h1 - 1) : max(min(r_logical, h1 - 1), 0);
        r = max(min(r - srow0, n_srows - 1), 0);
        const float *rc = &s_c[r][0];
        const float *rg = &s_g[r][0];
        if (hfast) {
            oc = up_pair_even_origin(rc[scol], rc[scol + 1], rc[scol + 2]);
            og = up_pair_even_origin(rg[scol], rg[scol + 1], rg[scol + 2]);
        } else {
            oc.a = oc.b = og.a = og.b = 0.f;
            if (ina) {
                oc.a = up_h_sample(txa, w1, [&](int ii) { return rc[ii - sx0]; });
                og.a = up_h_sample(txa, w1, [&](int ii) { return rg[ii - sx0]; });
            }
            if (inb) {
                oc.b = up_h_sample(txb, w1, [&](int ii) { return rc[ii - sx0]; });
                og.b = up_h_sample(txb, w1, [&](int ii) { return rg[ii - sx0]; });
            }
        }
    };

    // software prefetch of the per-output operands (G_l, W_l): one step (two rows) ahead, in registers
    float2 gnext[2], wnext[2];
    unsigned graw[2] = {0u, 0u}; // G0U16: the raw pair stays in flight; it is converted where it is used, not where it is loaded
    auto fetch_gw = [&](int k) {
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const int ty = ty0 + 2 * k + a;
            gnext[a] = make_float2(0.f, 0.f);
            wnext[a] = make_float2(0.f, 0.f);
            if (G0U16) graw[a] = 0u;
            if (ty < ty1 && ina) {
                const size_t off = (size_t)ty * L0.pitch + txa;
                if (G0U16) { // two uint16 (value * 256) in one aligned 32-bit word: txa is even, the plane pitch a multiple of 8
                    graw[a] = __ldg(reinterpret_cast<const unsigned *>(G0h + (size_t)ty * g.g1_pitch + txa));
                    if (inb) {
                        wnext[a] = __ldg(reinterpret_cast<const float2 *>(W0 + off));
                    } else {
                        wnext[a].x = __ldg(W0 + off);
                    }
                } else if (inb) {
                    gnext[a] = __ldg(reinterpret_cast<const float2 *>(G0 + off));
                    wnext[a] = __ldg(reinterpret_cast<const float2 *>(W0 + off));
                } else {
                    gnext[a].x = __ldg(G0 + off);
                    wnext[a].x = __ldg(W0 + off);
                }
            }
        }
    };
    fetch_gw(0); // in flight while the source window lands

    cp_async_wait<0>();
    __syncthreads();

    HPair2 rc0, rg0, rc1, rg1, rc2, rg2;
    hrow(ws0, rc0, rg0);
    hrow(ws0 + 1, rc1, rg1);
#pragma unroll 4
    for (int k = 0; k < STEPS; ++k) {
        hrow(ws0 + k + 2, rc2, rg2);
        float uc[2][2], ug[2][2];
        if (ypar == 0) {
            uc[0][0] = up_even(rc0.a, rc1.a, rc2.a); uc[0][1] = up_even(rc0.b, rc1.b, rc2.b);
            ug[0][0] = up_even(rg0.a, rg1.a, rg2.a); ug[0][1] = up_even(rg0.b, rg1.b, rg2.b);
            uc[1][0] = up_odd(rc1.a, rc2.a); uc[1][1] = up_odd(rc1.b, rc2.b);
            ug[1][0] = up_odd(rg1.a, rg2.a); ug[1][1] = up_odd(rg1.b, rg2.b);
        } else {
            uc[0][0] = up_odd(rc0.a, rc1.a); uc[0][1] = up_odd(rc0.b, rc1.b);
            ug[0][0] = up_odd(rg0.a, rg1.a); ug[0][1] = up_odd(rg0.b, rg1.b);
            uc[1][0] = up_even(rc0.a, rc1.a, rc2.a); uc[1][1] = up_even(rc0.b, rc1.b, rc2.b);
            ug[1][0] = up_even(rg0.a, rg1.a, rg2.a); ug[1][1] = up_even(rg0.b, rg1.b, rg2.b);
        }
        // this step's G_l / W_l were fetched one step ago; fetch the next step's now
        float2 gcur[2] = {gnext[0], gnext[1]}, wcur[2] = {wnext[0], wnext[1]};
        if (G0U16) {
#pragma unroll
            for (int a = 0; a < 2; ++a)
                gcur[a] = make_float2(__fmul_rn((float)(graw[a] & 0xFFFFu), 1.f / 256.f), __fmul_rn((float)(graw[a] >> 16), 1.f / 256.f));
        }
        if (k + 1 < STEPS) fetch_gw(k + 1);
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const int ty = ty0 + 2 * k + a;
            if (ty >= ty1 || !ina) continue;
            const size_t off = (size_t)ty * L0.pitch + txa;
            const float g0a = gcur[a].x, g0b = gcur[a].y, w0a = wcur[a].x, w0b = wcur[a].y;
            const float oa = __fadd_rn(__fmul_rn(uc[a][0], 1.f / 64.f), __fmul_rn(__fsub_rn(g0a, __fmul_rn(ug[a][0], 1.f / 64.f)), w0a));
            const float ob = __fadd_rn(__fmul_rn(uc[a][1], 1.f / 64.f), __fmul_rn(__fsub_rn(g0b, __fmul_rn(ug[a][1], 1.f / 64.f)), w0b));
            if (inb)
                *reinterpret_cast<float2 *>(C0 + off) = make_float2(oa, ob);
            else
                C0[off] = oa;
        }
        rc0 = rc1; rg0 = rg1; rc1 = rc2; rg1 = rg2;
    }
}

// ------------------------------------------------------------------------------------------------------------
// k_final: canvas gather.  One CTA per FB_W x 16 canvas block (128 x 16 by default, 64*C threads, four CTAs per SM):
// thread (p, c) owns canvas columns cx0+2p, cx0+2p+1 of channel c for all 16 rows -- 32 accumulators (16 packed float2)
// in registers; the sum of W_0 lives in smem.  Per covering tile (list order = the reference's accumulation order) the
// C_1 / G_1 source windows (10 rows x (FB_W/2 + 8) columns per channel) and the uint8 tile block arrive by TMA tensor
// copies (one 3-D box per array, one 2-D box for the bytes; cp.async for 1- and 2-level tiles), double buffered across
// tiles, then every thread slides DOWN the rows with a 3-row register ring of horizontally up-sampled values: one new
// source row per canvas row pair, 3 LDS per array for two columns, vertical pass + Laplacian weighting + accumulation
// straight from registers (packed FADD2 / FMUL2 / FFMA2 on the predicate-free interior path).
// Tile origins of either parity are handled by shifting the window (xpar / ypar are uniform per CTA and tile); only even
// parity takes the interior path.
// MODE per tile: 0 = >= 3 levels (C_1 stored), 1 = exactly 2 levels (C_1 = G_1 * W_1 on the fly), 2 = single level
// (C_0 = G_0 * W_0, no pyramid).
// ------------------------------------------------------------------------------------------------------------
#ifndef SRF_FB_W
#define SRF_FB_W 128
#endif
#ifndef SRF_FB_H
#define SRF_FB_H 16
#endif
#ifndef SRF_FB_CTAS
#define SRF_FB_CTAS ((SRF_FB_W == 128) ? 4 : 2)
#endif
constexpr int FB_W = SRF_FB_W, FB_H = SRF_FB_H, FB_PAIRS = FB_W / 2, FB_STEPS = FB_H / 2;
constexpr int FB_CTAS_3 = SRF_FB_CTAS, FB_CTAS_1 = 2 * SRF_FB_CTAS; // resident CTAs per SM the smem / registers allow
constexpr int FB_SRC_ROWS = FB_STEPS + 2;   // 10
constexpr int FB_SRC_COLS = FB_PAIRS + 8;   // 72
constexpr int FB_U8_STRIDE = FB_W * 3 + 16; // payload + up to 15 B misalignment, a multiple of 16

template <int C>
struct alignas(128) FinalStage { // everything the gather reads of ONE covering tile; double buffered; each array is the
                                 // dense box of one TMA tensor copy (128-byte aligned) or is filled by cp.async
    alignas(128) float sc[C][FB_SRC_ROWS][FB_SRC_COLS]; // C_1 window (W_1 window when the tile has exactly two levels)
    alignas(128) float sg[C][FB_SRC_ROWS][FB_SRC_COLS]; // G_1 window
    alignas(128) uint8_t u8[FB_H][FB_U8_STRIDE];        // uint8 tile block
    float nwy[FB_H];                       // -W_0 profile of the block's 16 rows (negated: see final_accumulate_tile)
};
template <int C>
struct alignas(128) FinalSmem {
    FinalStage<C> st[2];
    float wsum[FB_H][FB_W];
};


// per-(CTA, tile) geometry, uniform across the CTA
struct FinalGeo {
    int iy0, iy1, ix0, ix1; // intersection of block and tile, canvas coordinates
    int mode;               // 0: >= 3 levels, 1: exactly 2 (C_1 = G_1 * W_1), 2: single level
    int xpar, ypar, i0, ws0;
    int sx0, srow0, n_srows;
    int u8_a0, u8_chunks;
};
__device__ __forceinline__ FinalGeo final_geo(const TileDev &t, const GroupDev &g, int cy0, int cy1, int cx0, int cx1, int C)
{
    FinalGeo q;
    q.iy0 = max(cy0, t.y); q.iy1 = min(cy1, t.y + t.h);
    q.ix0 = max(cx0, t.x); q.ix1 = min(cx1, t.x + t.w);
    q.mode = (g.n_levels >= 3) ? 0 : (g.n_levels == 2 ? 1 : 2);
    const int h1 = g.lv[1].h;
    const int tyf = cy0 - t.y, txf = cx0 - t.x; // tile coordinates of the block origin (may be negative)
    q.xpar = txf & 1; q.ypar = tyf & 1;
    q.i0 = (txf >> 1) - (1 - q.xpar);  // window start column for pair 0
    q.ws0 = (tyf >> 1) - (1 - q.ypar); // window start row for step 0
    q.sx0 = (q.i0 >> 2) * 4; // floor to a multiple of 4 (may be negative: those columns are staged as zeros)
    q.srow0 = max(min(q.ws0, h1 - 1), 0);
    q.n_srows = max(min(q.ws0 + FB_STEPS + 1, h1 - 1), 0) - q.srow0 + 1; // <= 10
    q.u8_a0 = ((q.ix0 - t.x) * C) & ~15;
    q.u8_chunks = ((q.ix1 - t.x) * C - q.u8_a0 + 15) >> 4;
    return q;
}

// Tiles with >= 3 levels are staged by TMA tensor copies when the launch carries their descriptors (bit ti of
// tma_mask): one 3-D box (136 x 10 x C floats) each for the C_1 and G_1 windows and one 2-D box (800 bytes x 16 rows)
// for the uint8 block; whatever part of a box lies outside its array arrives as zeros, which is what the gather expects
// of out-of-array columns.  Other tiles (1 or 2 levels, no descriptor) keep the per-chunk cp.async path.
constexpr int FINAL_TMA_TILES = 48; // descriptors travel as kernel parameters: 48 x 3 x 128 B = 18 KB of the 32 KB limit
struct FinalTmaps {
    CUtensorMap m[FINAL_TMA_TILES][3]; // [tile][C_1 | G_1 | uint8]
};
__device__ __forceinline__ bool final_tma_ok(const FinalGeo &q, unsigned long long tma_mask, int ti)
{
    return q.mode == 0 && ti < FINAL_TMA_TILES && ((tma_mask >> ti) & 1ull);
}
__device__ __forceinline__ void tma_tensor_3d_g2s(void *smem_dst, const CUtensorMap *map, int x, int y, int z, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(reinterpret_cast<unsigned long long>(map)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_tensor_2d_g2s(void *smem_dst, const CUtensorMap *map, int x, int y, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(reinterpret_cast<unsigned long long>(map)), "r"(x), "r"(y), "r"(smem_u32(bar))
                 : "memory");
}
// mbarrier wait that traps instead of hanging the GPU should a copy never land
__device__ __forceinline__ void mbar_wait_or_trap(uint64_t *bar, unsigned parity)
{
    for (unsigned spin = 0; spin < (1u << 22); ++spin) {
        unsigned ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return;
    }
    if ((threadIdx.x & 31) == 0) printf("[srfuse] mbarrier wait timed out: block (%d,%d,%d) warp %d parity %u\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x >> 5, parity);
    __trap();
}

struct FinalParams {
    const TileDev *tiles;
    const GroupDev *groups;
    const int *tile_group; // group index per tile
    const int4 *rects;     // (y, x, h, w) per tile, compact copy for the coverage test
    int n_tiles;
    int out_h, out_w;
    uint8_t *canvas;
    size_t canvas_pitch;
    int row_begin, row_end;
    unsigned long long tma_mask; // tiles whose FinalTmaps entries are valid
};

struct HPair {
    float a, b; // horizontally up-sampled values (x8, un-normalised) for the thread's two columns
};

// horizontal pass for the thread's column pair from source samples s0,s1,s2 = s[i0+p .. i0+p+2] (interior form)
__device__ __forceinline__ HPair h_pair_fast(float s0, float s1, float s2, int xpar)
{
    HPair r;
    if (xpar == 0) { // columns (2i, 2i+1), i = window centre
        r.a = up_even(s0, s1, s2);
        r.b = up_odd(s1, s2);
    } else {         // columns (2i+1, 2i+2), i = window start
        r.a = up_odd(s0, s1);
        r.b = up_even(s0, s1, s2);
    }
    return r;
}

// All arithmetic of the gather runs in the "x64 domain": pyrUp's two un-normalised passes leave a factor 64 that the
// reference removes right away (x 1/64).  Scaling by a power of two commutes with every IEEE rounding, so
// 64*(upC + (g - upG)*w) == up64C + (64 g - up64G)*w bit for bit; the factor is taken out once, before the divide.
template <int C, bool FULL>
__device__ __forceinline__ void final_accumulate_tile(FinalSmem<C> &S, const FinalStage<C> &B, float2 (&acc)[FB_STEPS][2], const TileDev &t,
                                                      const GroupDev &g, int cy0, int cx0, const FinalGeo &q, int c, int p)
{
    const int iy0 = q.iy0, iy1 = q.iy1, ix0 = q.ix0, ix1 = q.ix1, u8_a0 = q.u8_a0, srow0 = q.srow0, n_srows = q.n_srows, sx0 = q.sx0;
    const int i0 = q.i0, ws0 = q.ws0, xpar = q.xpar, ypar = q.ypar, mode = q.mode;
    const int h1 = g.lv[1].h, w1 = g.lv[1].w;
    const int txf = cx0 - t.x;
    const int txa = txf + 2 * p, txb = txa + 1; // tile columns of the thread's two canvas columns
    const bool ina = FULL || ((cx0 + 2 * p >= ix0) && (cx0 + 2 * p < ix1));
    const bool inb = FULL || ((cx0 + 2 * p + 1 >= ix0) && (cx0 + 2 * p + 1 < ix1));
    const int lmax = g.lut_len - 1;
    const int dxa = min(txa, t.w - 1 - txa), dxb = min(txb, t.w - 1 - txb);
    const int u8_col = txa * C + c - u8_a0; // byte offset of (txa, c) inside a staged row
    const int scol = i0 + p - sx0; // smem column of s0 (>= 0: the staged window starts at floor4(i0), zero-filled left of 0)
    // columns whose horizontal sample needs an edge formula (tx = 0, the last source column, a 1-wide source)
    const bool spa = !FULL && ((txa <= 0) || ((txa >> 1) >= w1 - 1));
    const bool spb = !FULL && ((txb <= 0) || ((txb >> 1) >= w1 - 1));

    // The weight profile is monotone non-decreasing in the border distance (linear, cosine and sigmoid all are, and
    // float32 rounding keeps the order), so lut[min(dy, dx)] == min(lut[dy], lut[dx]) exactly: one lookup per row and
    // two per tile instead of one per pixel.
    // The row values come from a 16-entry table the staging step leaves in smem (B.nwy, negated).
    const float wxa = __ldg(g.lut + max(min(dxa, lmax), 0)), wxb = __ldg(g.lut + max(min(dxb, lmax), 0));
    auto hrow = [&](int r_logical, HPair &oc, HPair &og) {
        int r = (r_logical == -1) ? min(1, h1 - 1) : max(min(r_logical, h1 - 1), 0);
        r -= srow0;
        r = max(min(r, n_srows - 1), 0);
        const float *rc = &B.sc[c][r][0];
        const float *rg = &B.sg[c][r][0];
        if (FULL || mode == 0) {
            // interior formula for everyone (windows that poke outside the source row read staged zeros / padding and
            // only ever feed columns that are outside the tile or patched right below)
            oc = h_pair_fast(rc[scol], rc[scol + 1], rc[scol + 2], xpar);
            og = h_pair_fast(rg[scol], rg[scol + 1], rg[scol + 2], xpar);
            if (!FULL) {
                if (spa && ina) {
                    oc.a = up_h_sample(txa, w1, [&](int ii) { return rc[ii - sx0]; });
                    og.a = up_h_sample(txa, w1, [&](int ii) { return rg[ii - sx0]; });
                }
                if (spb && inb) {
                    oc.b = up_h_sample(txb, w1, [&](int ii) { return rc[ii - sx0]; });
                    og.b = up_h_sample(txb, w1, [&](int ii) { return rg[ii - sx0]; });
                }
            }
        } else { // exactly two levels: C_1 = G_1 * W_1 formed on the fly (sc holds the W_1 window)
            oc.a = oc.b = og.a = og.b = 0.f;
            auto atg = [&](int ii) { return rg[ii - sx0]; };
            auto atc = [&](int ii) { return __fmul_rn(rg[ii - sx0], rc[ii - sx0]); };
            if (ina) {
                oc.a = up_h_sample(txa, w1, atc);
                og.a = up_h_sample(txa, w1, atg);
            }
            if (inb) {
                oc.b = up_h_sample(txb, w1, atc);
                og.b = up_h_sample(txb, w1, atg);
            }
        }
    };

    if (mode == 2) { // single level: C_0 = G_0 * W_0
#pragma unroll
        for (int k = 0; k < FB_STEPS; ++k) {
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                const int cy = cy0 + 2 * k + a;
                if (cy < iy0 || cy >= iy1) continue;
                const float wy = -B.nwy[2 * k + a];
                const uint8_t *px = &B.u8[2 * k + a][u8_col];
                const bool wsum_mine = c == (2 * k + a) % C; // the block's rows are dealt round-robin to the channel threads
                if (ina) {
                    const float w0 = fminf(wy, wxa);
                    acc[k][a].x = __fadd_rn(acc[k][a].x, __fmul_rn((float)((int)px[0] << 6), w0));
                    if (wsum_mine) S.wsum[2 * k + a][2 * p] = __fadd_rn(S.wsum[2 * k + a][2 * p], w0);
                }
                if (inb) {
                    const float w0 = fminf(wy, wxb);
                    acc[k][a].y = __fadd_rn(acc[k][a].y, __fmul_rn((float)((int)px[C] << 6), w0));
                    if (wsum_mine) S.wsum[2 * k + a][2 * p + 1] = __fadd_rn(S.wsum[2 * k + a][2 * p + 1], w0);
                }
            }
        }
        return;
    }

    if (FULL) {
        // Interior block of even parity (the caller checked: block inside the tile, every source window strictly inside
        // the level-1 plane in both axes, xpar == ypar == 0): no predicates, staged source row j is logical row ws0 + j,
        // and the column pair is carried as one packed float2 (FADD2 / FMUL2 / FFMA2 round each lane like the scalar
        // instruction, so the reference's op order is kept).  ptxas contracts a packed multiply into a following packed
        // add even for .rn, so every INEXACT product that feeds an add (6 p, d * w) is formed with scalar FMULs; the exact
        // ones (x 4) may fuse freely.  The pixel stage runs on negated values: nd = fma(g, -64, up64G) = -(64 g - up64G)
        // and nw = -w0, whose product is the reference's (64 g - up64G) * w0.
        const float nwxa = -wxa, nwxb = -wxb;
        const float *rcb = &B.sc[c][0][scol];
        const float *rgb = &B.sg[c][0][scol];
        auto hrow2 = [&](int j, float2 &oc, float2 &og) { // columns (2i, 2i+1), i = window centre: even form, odd form
            const float *rc = rcb + j * FB_SRC_COLS, *rg = rgb + j * FB_SRC_COLS;
            const float c0 = rc[0], c1 = rc[1], c2 = rc[2], g0 = rg[0], g1 = rg[1], g2 = rg[2];
            oc = make_float2(up_even(c0, c1, c2), up_odd(c1, c2));
            og = make_float2(up_even(g0, g1, g2), up_odd(g1, g2));
        };
        auto up_even2 = [](float2 m, float2 p6, float2 n) {
            return __fadd2_rn(__fadd2_rn(make_float2(__fmul_rn(p6.x, 6.f), __fmul_rn(p6.y, 6.f)), m), n);
        };
        auto up_odd2 = [](float2 p4, float2 n) { return __fmul2_rn(__fadd2_rn(p4, n), make_float2(4.f, 4.f)); };
        float2 rc0, rg0, rc1, rg1, rc2, rg2; // ring: source rows k, k+1, k+2
        hrow2(0, rc0, rg0);
        hrow2(1, rc1, rg1);
#pragma unroll
        for (int k = 0; k < FB_STEPS; ++k) {
            hrow2(k + 2, rc2, rg2);
            float2 uc[2], ug[2]; // x64, tile rows (2j, 2j+1), j = ws0 + k + 1: even uses the whole ring, odd its last two rows
            uc[0] = up_even2(rc0, rc1, rc2); ug[0] = up_even2(rg0, rg1, rg2);
            uc[1] = up_odd2(rc1, rc2);       ug[1] = up_odd2(rg1, rg2);
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                const int r = 2 * k + a;
                const float nwy = B.nwy[r];
                const uint8_t *px = &B.u8[r][u8_col];
                const float2 nw = make_float2(fmaxf(nwy, nwxa), fmaxf(nwy, nwxb)); // -min(wy, wx)
                const float2 nd = __ffma2_rn(make_float2(__int2float_rn((int)px[0]), __int2float_rn((int)px[C])), make_float2(-64.f, -64.f), ug[a]);
                const float2 dw = make_float2(__fmul_rn(nd.x, nw.x), __fmul_rn(nd.y, nw.y));
                acc[k][a] = __fadd2_rn(acc[k][a], __fadd2_rn(uc[a], dw));
                if (c == r % C) {
                    float2 *wp = reinterpret_cast<float2 *>(&S.wsum[r][2 * p]);
                    float2 ws = *wp;
                    ws.x = __fadd_rn(ws.x, -nw.x);
                    ws.y = __fadd_rn(ws.y, -nw.y);
                    *wp = ws;
                }
            }
            rc0 = rc1; rg0 = rg1; rc1 = rc2; rg1 = rg2;
        }
        return;
    }

    HPair rc0, rg0, rc1, rg1, rc2, rg2; // ring: logical rows ws0+k, ws0+k+1, ws0+k+2
    hrow(ws0, rc0, rg0);
    hrow(ws0 + 1, rc1, rg1);
#pragma unroll
    for (int k = 0; k < FB_STEPS; ++k) {
        hrow(ws0 + k + 2, rc2, rg2);
        float uc[2][2], ug[2][2]; // x64
        if (ypar == 0) { // tile rows (2j, 2j+1), j = ws0+k+1: even uses (m,p,n), odd uses (p,n)
            uc[0][0] = up_even(rc0.a, rc1.a, rc2.a); uc[0][1] = up_even(rc0.b, rc1.b, rc2.b);
            ug[0][0] = up_even(rg0.a, rg1.a, rg2.a); ug[0][1] = up_even(rg0.b, rg1.b, rg2.b);
            uc[1][0] = up_odd(rc1.a, rc2.a); uc[1][1] = up_odd(rc1.b, rc2.b);
            ug[1][0] = up_odd(rg1.a, rg2.a); ug[1][1] = up_odd(rg1.b, rg2.b);
        } else {         // tile rows (2j+1, 2j+2), j = ws0+k: odd uses ring 0,1; even uses ring 0,1,2
            uc[0][0] = up_odd(rc0.a, rc1.a); uc[0][1] = up_odd(rc0.b, rc1.b);
            ug[0][0] = up_odd(rg0.a, rg1.a); ug[0][1] = up_odd(rg0.b, rg1.b);
            uc[1][0] = up_even(rc0.a, rc1.a, rc2.a); uc[1][1] = up_even(rc0.b, rc1.b, rc2.b);
            ug[1][0] = up_even(rg0.a, rg1.a, rg2.a); ug[1][1] = up_even(rg0.b, rg1.b, rg2.b);
        }
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const int cy = cy0 + 2 * k + a;
            if (FULL || (cy >= iy0 && cy < iy1)) {
                const float wy = -B.nwy[2 * k + a];
                const uint8_t *px = &B.u8[2 * k + a][u8_col];
                const bool wsum_mine = c == (2 * k + a) % C;
                if (ina) {
                    const float w0 = fminf(wy, wxa);
                    // fma(g, 64, -up64G) == 64 g - up64G rounded once: 64 g is exact, so this IS the reference's subtraction
                    const float v = __fadd_rn(uc[a][0], __fmul_rn(__fmaf_rn((float)px[0], 64.f, -ug[a][0]), w0));
                    acc[k][a].x = __fadd_rn(acc[k][a].x, v);
                    if (wsum_mine) S.wsum[2 * k + a][2 * p] = __fadd_rn(S.wsum[2 * k + a][2 * p], w0);
                }
                if (inb) {
                    const float w0 = fminf(wy, wxb);
                    const float v = __fadd_rn(uc[a][1], __fmul_rn(__fmaf_rn((float)px[C], 64.f, -ug[a][1]), w0));
                    acc[k][a].y = __fadd_rn(acc[k][a].y, v);
                    if (wsum_mine) S.wsum[2 * k + a][2 * p + 1] = __fadd_rn(S.wsum[2 * k + a][2 * p + 1], w0);
                }
            }
        }
        rc0 = rc1; rg0 = rg1; rc1 = rc2; rg1 = rg2;
    }
}

constexpr float FINAL_NEAR = 2.5e-4f;
// MUFU.RCP: <= 1 ulp, exact for powers of two
__device__ __forceinline__ float rcp_approx(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// trunc(clip(v, 0, 255)): float -> u8 conversion rounds toward zero and saturates (NaN -> 0)
__device__ __forceinline__ unsigned f2u8_trunc_sat(float v)
{
    unsigned r;
    asm("cvt.rzi.u8.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}

template <int C>
__global__ void __launch_bounds__(FB_PAIRS * C, (C == 3) ? FB_CTAS_3 : FB_CTAS_1) k_final(const __grid_constant__ FinalParams P, const __grid_constant__ FinalTmaps TM)
{
    pdl_enter(); // programmatic dependent launch: nothing above this line may touch global data
    extern __shared__ __align__(128) uint8_t smem_raw[];
    FinalSmem<C> &S = *reinterpret_cast<FinalSmem<C> *>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127); // +128 B are allocated
    constexpr int NT = FB_PAIRS * C;
    __shared__ int s_list[NT];
    __shared__ int s_wcnt[NT / 32];
    __shared__ __align__(8) uint64_t s_bar[2]; // completion of the TMA-staged tile in buffer 0 / 1
    unsigned tma_phase = 0;                    // bit b: parity of the next completion of s_bar[b]
    if (threadIdx.x == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_fence_init();
    }

    const int cy0 = P.row_begin + blockIdx.y * FB_H;
    const int cx0 = blockIdx.x * FB_W;
    const int cy1 = min(cy0 + FB_H, P.row_end);
    const int cx1 = min(cx0 + FB_W, P.out_w);
    const int c = threadIdx.x / FB_PAIRS; // warp-uniform
    const int p = threadIdx.x % FB_PAIRS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);

    float2 acc[FB_STEPS][2];
#pragma unroll
    for (int k = 0; k < FB_STEPS; ++k) acc[k][0] = acc[k][1] = make_float2(0.f, 0.f);
    for (int i = threadIdx.x; i < FB_H * FB_W; i += NT) (&S.wsum[0][0])[i] = 0.f;

    for (int base = 0; base < P.n_tiles; base += NT) {
        // ---- ordered list of the tiles (of this chunk) that touch the block: one tile per thread, ballot-compacted ----
        const int cand = base + threadIdx.x;
        bool hit = false;
        if (cand < P.n_tiles) {
            const int4 rct = P.rects[cand]; // y, x, h, w
            hit = max(cy0, rct.x) < min(cy1, rct.x + rct.z) && max(cx0, rct.y) < min(cx1, rct.y + rct.w);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        __syncthreads(); // previous chunk's list fully consumed
        if (lane == 0) s_wcnt[warp] = __popc(bal);
        __syncthreads();
        int off = 0, total = 0;
#pragma unroll
        for (int k = 0; k < NT / 32; ++k) {
            const int n = s_wcnt[k];
            if (k < warp) off += n;
            total += n;
        }
        if (hit) s_list[off + __popc(bal & ((1u << lane) - 1u))] = cand;
        __syncthreads();

        // stage(li): cp.async the tile's uint8 block and C_1 / G_1 windows into buffer li & 1 (chunks outside the
        // arrays are zero-filled with plain stores; everything else lands asynchronously)
        auto stage = [&](int li) {
            const int ti = s_list[li];
            const TileDev &t = P.tiles[ti];
            const GroupDev &g = P.groups[P.tile_group[ti]];
            const FinalGeo q = final_geo(t, g, cy0, cy1, cx0, cx1, C);
            FinalStage<C> &B = S.st[li & 1];
            if (threadIdx.x < FB_H) { // W_0 row profile of the block (blending_module.py:508-561 through the host LUT), negated
                const int ty = cy0 + (int)threadIdx.x - t.y;
                B.nwy[threadIdx.x] = -__ldg(g.lut + max(min(min(ty, t.h - 1 - ty), g.lut_len - 1), 0));
            }
            if (final_tma_ok(q, P.tma_mask, ti)) {
                if (threadIdx.x == 0) {
                    uint64_t *bar = &s_bar[li & 1];
                    fence_proxy_async(); // the buffer's earlier generic-proxy accesses (ordered by the CTA barrier) before the async writes
                    mbar_expect_tx(bar, (unsigned)(sizeof(B.sc) + sizeof(B.sg) + sizeof(B.u8)));
                    tma_tensor_3d_g2s(&B.sc[0][0][0], &TM.m[ti][0], q.sx0, q.srow0, 0, bar);
                    tma_tensor_3d_g2s(&B.sg[0][0][0], &TM.m[ti][1], q.sx0, q.srow0, 0, bar);
                    tma_tensor_2d_g2s(&B.u8[0][0], &TM.m[ti][2], q.u8_a0 >> 2, cy0 - t.y, bar);
                }
                cp_async_commit(); // an empty group keeps the cp.async group count of the fallback path in step
                return;
            }
            const int pitch = (int)t.pitch;
            constexpr int NW = NT / 32;
            // one warp per row, one lane per 16-byte chunk (+ a second pass for the chunks beyond 32): no div / mod
            for (int r = warp; r < q.iy1 - q.iy0; r += NW) {
                const uint8_t *grow = t.src + (size_t)(q.iy0 - t.y + r) * t.pitch + q.u8_a0;
                uint8_t *drow = &B.u8[q.iy0 - cy0 + r][0];
                for (int ch = lane; ch < q.u8_chunks; ch += 32) {
                    const int o = q.u8_a0 + ch * 16;
                    if (o >= 0 && o + 16 <= pitch)
                        cp_async_16(drow + ch * 16, grow + ch * 16);
                    else
                        *reinterpret_cast<uint4 *>(drow + ch * 16) = make_uint4(0, 0, 0, 0);
                }
            }
            if (q.mode != 2) {
                const LevelGeom L1 = g.lv[1];
                const float *csrc = (q.mode == 1) ? g.W[1] : t.C[1];
                const size_t cplane = (q.mode == 1) ? 0 : L1.plane;
                for (int r = warp; r < q.n_srows; r += NW) {
                    const ptrdiff_t roff = (ptrdiff_t)(q.srow0 + r) * L1.pitch + q.sx0;
                    for (int ch = lane; ch < FB_SRC_COLS / 4; ch += 32) {
                        const int c0 = q.sx0 + ch * 4;
                        const bool ok = c0 >= 0 && c0 + 4 <= L1.pitch;
#pragma unroll
                        for (int cc = 0; cc < C; ++cc) {
                            float *dg = &B.sg[cc][r][ch * 4], *dc = &B.sc[cc][r][ch * 4];
                            if (ok) {
                                cp_async_16(dg, t.G[1] + (size_t)cc * L1.plane + roff + ch * 4);
                                cp_async_16(dc, csrc + (size_t)cc * cplane + roff + ch * 4);
                            } else {
                                *reinterpret_cast<float4 *>(dg) = make_float4(0.f, 0.f, 0.f, 0.f);
                                *reinterpret_cast<float4 *>(dc) = make_float4(0.f, 0.f, 0.f, 0.f);
                            }
                        }
                    }
                }
            }
            cp_async_commit();
        };

        for (int li = -1; li < total; ++li) { // li = -1: prologue, stages tile 0 only (one inlined copy of the staging code)
            if (li + 1 < total) stage(li + 1); // next tile's copies fly while this tile is consumed
            if (li < 0) continue;
            if (li + 1 < total)
                cp_async_wait<1>();
            else
                cp_async_wait<0>();
            const int ti = s_list[li];
            const TileDev &t = P.tiles[ti];
            const GroupDev &g = P.groups[P.tile_group[ti]];
            const FinalGeo q = final_geo(t, g, cy0, cy1, cx0, cx1, C);
            if (final_tma_ok(q, P.tma_mask, ti)) {
                mbar_wait_or_trap(&s_bar[li & 1], (tma_phase >> (li & 1)) & 1u);
                tma_phase ^= 1u << (li & 1);
            }
            __syncthreads(); // tile li has landed for every thread (cp.async groups, TMA bytes, the row-weight table)
            // block entirely inside the tile, every source window strictly inside the level-1 plane (columns and rows) and
            // even parity in both axes (tile origins are 2 * the low-resolution origin in the pipeline): no predicates
            const bool full = (q.mode == 0) && (q.iy0 == cy0) && (q.iy1 == cy0 + FB_H) && (q.ix0 == cx0) && (q.ix1 == cx0 + FB_W) &&
                              (q.i0 >= 0) && (q.i0 + FB_PAIRS + 1 <= g.lv[1].w - 1) && (q.xpar == 0) && (q.ypar == 0) && (q.ws0 >= 0) &&
                              (q.ws0 + FB_STEPS + 1 <= g.lv[1].h - 1);
            if (full)
                final_accumulate_tile<C, true>(S, S.st[li & 1], acc, t, g, cy0, cx0, q, c, p);
            else
                final_accumulate_tile<C, false>(S, S.st[li & 1], acc, t, g, cy0, cx0, q, c, p);
            __syncthreads(); // buffer li & 1 is free for tile li + 2
        }
    }

    // ---- normalise, clip, truncate; stage the block as bytes congruent (mod 16) with the canvas rows ----
    // The reference's out = trunc(clip(RN(acc / wsum), 0, 255)) only needs the correctly rounded quotient where it lies
    // next to an integer: q ~ acc * rcp(wsum) is within 8e-5 of RN(acc / wsum) for q < 256 (MUFU.RCP + FMUL: 2 ulp, plus
    // the half ulp of the exact rounding), so unless q is within FINAL_NEAR of an integer both truncate to the same byte.
    // The IEEE division (~15 instructions per value) runs only for those rare pairs; cvt.rzi.u8 saturates, i.e. clips.
    __syncthreads();
    const int row_bytes = (cx1 - cx0) * C;
    {
        const uintptr_t cbase = reinterpret_cast<uintptr_t>(P.canvas) + (size_t)cy0 * P.canvas_pitch + (size_t)cx0 * C;
        const int mis0 = (int)(cbase & 15), pit16 = (int)(P.canvas_pitch & 15);
        uint8_t *out0 = &S.st[0].u8[0][2 * p * C + c];
        auto quantise = [&](auto whole) { // whole: the block lies entirely inside the band / canvas, no row or column tests
            constexpr bool WHOLE = decltype(whole)::value;
            const bool px_a = WHOLE || cx0 + 2 * p < cx1, px_b = WHOLE || cx0 + 2 * p + 1 < cx1;
#pragma unroll
            for (int k = 0; k < FB_STEPS; ++k) {
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                    const int r = 2 * k + a;
                    if (!WHOLE && cy0 + r >= cy1) continue;
                    const int mis = (mis0 + r * pit16) & 15;
                    const float2 ws = *reinterpret_cast<const float2 *>(&S.wsum[r][2 * p]);
                    const float b0 = fmaxf(ws.x, 1e-6f), b1 = fmaxf(ws.y, 1e-6f);
                    const float a0 = __fmul_rn(acc[k][a].x, 1.f / 64.f), a1 = __fmul_rn(acc[k][a].y, 1.f / 64.f);
                    float q0 = a0 * rcp_approx(b0), q1 = a1 * rcp_approx(b1);
                    // a power-of-two divisor (1 where one tile covers the pixel, 2 / 4 where fully weighted tiles overlap) makes
                    // q exact -- those are also the pixels whose quotient IS an integer, so they must not take the slow path
                    const bool n0 = fabsf(q0 - rintf(q0)) < FINAL_NEAR && (__float_as_uint(b0) & 0x007FFFFFu) != 0u;
                    const bool n1 = fabsf(q1 - rintf(q1)) < FINAL_NEAR && (__float_as_uint(b1) & 0x007FFFFFu) != 0u;
                    if (n0 || n1) {
                        q0 = __fdiv_rn(a0, b0);
                        q1 = __fdiv_rn(a1, b1);
                    }
                    uint8_t *dst = out0 + r * FB_U8_STRIDE + mis;
                    if (px_a) dst[0] = (uint8_t)f2u8_trunc_sat(q0); // clip then truncate (blending_module.py:504-506)
                    if (px_b) dst[C] = (uint8_t)f2u8_trunc_sat(q1);
                }
            }
        };
        if (cy1 - cy0 == FB_H && cx1 - cx0 == FB_W)
            quantise(std::true_type{});
        else
            quantise(std::false_type{});
    }
    __syncthreads();
    const int n_rows = cy1 - cy0;
    constexpr int kChunks = (FB_W * 3 + 15 + 15) / 16; // chunks that can touch a row's payload
    for (int i = threadIdx.x; i < n_rows * kChunks; i += NT) {
        const int r = i / kChunks, ch = i - r * kChunks;
        uint8_t *grow = P.canvas + (size_t)(cy0 + r) * P.canvas_pitch + (size_t)cx0 * C;
        const int mis = (int)(reinterpret_cast<uintptr_t>(grow) & 15);
        const int lo = ch * 16, hi = lo + 16; // byte range inside the staged row
        const int p0 = mis, p1 = mis + row_bytes; // payload range
        if (hi <= p0 || lo >= p1) continue;
        uint8_t *gdst = grow - mis + lo; // 16-byte aligned
        if (lo >= p0 && hi <= p1) {
            *reinterpret_cast<uint4 *>(gdst) = *reinterpret_cast<const uint4 *>(&S.st[0].u8[r][lo]);
        } else {
            for (int e = max(lo, p0); e < min(hi, p1); ++e) gdst[e - lo] = S.st[0].u8[r][e];
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// k_gather: the canvas gather of the regular pipeline, ROWS OUTERMOST.  k_final above walks the covering tiles one after
// the other and keeps 32 accumulators per thread across them (fully unrolled, 18 K instructions: its warps stall on
// instruction fetch); here every covering tile of the block is resident in shared memory at once and the thread slides
// down the 16 rows ONCE: per row pair it takes, tile after tile in list order, one new source row per array into that
// tile's 3-row register ring, forms the two output rows, adds them -- acc and the sum of W_0 are 4 + 4 registers -- then
// normalises and stages the two finished rows.  The loop is short and rolled (every warp of the SM runs the same few
// hundred instructions), there is no barrier inside it, W_0 sums never touch shared memory, and all TMA copies of the
// block are issued up front by one thread on one mbarrier.
// Which tiles cover which block is decided on the HOST: tile edges cut the canvas into cells of blocks that share one
// ordered tile list (GatherCell); a CTA finds its cell by binary search over the few hundred cells of its class and
// needs no coverage test.  Classes: TMAX = 2 (cells with <= 2 tiles; 2 stages of smem, 4 CTAs / SM) and TMAX = 4 (the
// 4-way corners; 2 CTAs / SM).  Preconditions (else srf_fuse_device launches k_final): every tile has >= 3 levels and a
// TMA descriptor, no cell has more than 4 tiles.
// Arithmetic, op order and the guarded-reciprocal epilogue are k_final's (see there): the bytes are identical.
// ------------------------------------------------------------------------------------------------------------
constexpr int GC_MAX_T = 4;
struct GatherCell {       // 48 bytes = three int4 loads
    int bx0, by0, nbx, nby; // block rectangle of the cell (block = FB_W x FB_H canvas pixels, rows counted from row_begin)
    int first, T;           // index of the cell's first block within its class; number of covering tiles
    int tile[GC_MAX_T];     // their indices, in list order
    int pad0, pad1;
};
struct GatherTile { // what the gather needs of a tile (kernel parameter space: no dependent global loads)
    int y, x, h, w;
    int h1, w1; // level-1 plane size
    int lut_len, pad;
    const float *lut;
};
struct GatherTiles {
    GatherTile t[FINAL_TMA_TILES];
};
struct GatherParams {
    const GatherCell *cells; // cells of this class, `first` ascending from 0
    int n_cells;
    int out_w;
    uint8_t *canvas;
    size_t canvas_pitch;
    int row_begin, row_end;
};
// A TMA box must start on a 16-byte boundary of the innermost dimension (any other start faults as an illegal instruction):
// the float32 C_1 window starts at a multiple of 4 columns, the uint16 G_1 window at a multiple of 8 and is up to 4 columns
// wider for it (66 columns needed + up to 7 of slack, rounded to whole 16 bytes)
constexpr int FB_G_COLS = 80;
template <int C>
struct alignas(128) GatherStage { // everything the row-outer gather reads of ONE covering tile: three dense TMA boxes
    alignas(128) float sc[C][FB_SRC_ROWS][FB_SRC_COLS];          // C_1 window
    alignas(128) unsigned short sg[C][FB_SRC_ROWS][FB_G_COLS];   // G_1 window, uint16 (value * 256); starts at a multiple of 8 columns
    alignas(128) uint8_t u8[FB_H][FB_U8_STRIDE];                 // uint8 tile block
    float nwy[FB_H];                                             // -W_0 profile of the block's 16 rows
};
template <int C, int TMAX>
struct alignas(128) GatherSmem {
    GatherStage<C> st[TMAX];
    alignas(16) uint8_t out[FB_H][FB_U8_STRIDE]; // finished rows, bytes congruent (mod 16) with the canvas rows
};
#ifndef SRF_GATHER_UNROLL
#define SRF_GATHER_UNROLL 2
#endif
#ifndef SRF_GATHER_CTAS
#define SRF_GATHER_CTAS 4
#endif
constexpr int GA_UNROLL = SRF_GATHER_UNROLL; // row pairs per loop iteration (2: the ring shift costs moves only every second step)

__host__ __device__ __forceinline__ FinalGeo gather_geo(const GatherTile &t, int cy0, int cy1, int cx0, int cx1, int C)
{
    FinalGeo q;
    q.iy0 = cy0 > t.y ? cy0 : t.y; q.iy1 = cy1 < t.y + t.h ? cy1 : t.y + t.h;
    q.ix0 = cx0 > t.x ? cx0 : t.x; q.ix1 = cx1 < t.x + t.w ? cx1 : t.x + t.w;
    q.mode = 0;
    const int tyf = cy0 - t.y, txf = cx0 - t.x;
    q.xpar = txf & 1; q.ypar = tyf & 1;
    q.i0 = (txf >> 1) - (1 - q.xpar);
    q.ws0 = (tyf >> 1) - (1 - q.ypar);
    q.sx0 = (q.i0 >> 2) * 4;
    const int lo = q.ws0 < t.h1 - 1 ? q.ws0 : t.h1 - 1;
    q.srow0 = lo > 0 ? lo : 0;
    const int hi = q.ws0 + FB_STEPS + 1 < t.h1 - 1 ? q.ws0 + FB_STEPS + 1 : t.h1 - 1;
    q.n_srows = (hi > 0 ? hi : 0) - q.srow0 + 1;
    q.u8_a0 = ((q.ix0 - t.x) * C) & ~15;
    q.u8_chunks = 0;
    return q;
}
// block entirely inside the tile, every source window strictly inside the level-1 plane (columns and rows), even parity in
// both axes: the predicate-free arithmetic applies.  Shared by the kernels and by the host's cell classification.
__host__ __device__ __forceinline__ bool gather_slot_full(const GatherTile &t, const FinalGeo &q, int cy0, int cx0)
{
    return (q.iy0 == cy0) && (q.iy1 == cy0 + FB_H) && (q.ix0 == cx0) && (q.ix1 == cx0 + FB_W) && (q.i0 >= 0) && (q.i0 + FB_PAIRS + 1 <= t.w1 - 1) &&
           (q.xpar == 0) && (q.ypar == 0) && (q.ws0 >= 0) && (q.ws0 + FB_STEPS + 1 <= t.h1 - 1);
}
// classes of cells (one launch each): what the host can prove about ALL blocks of a cell
enum { GC_U1 = 0,  // one tile, interior, W_0 == 1 everywhere in the cell: out = trunc(clip(C_0)), no weights, no division
       GC_F1 = 1,  // one tile, interior
       GC_F2 = 2,  // two tiles, both interior
       GC_F4 = 3,  // three or four tiles, all interior
       GC_G2 = 4,  // anything else with <= 2 tiles: tile edges, canvas edges, odd parity, no tile at all
       GC_G4 = 5,  // anything else with 3 or 4 tiles
       GC_E2 = 6,  // <= 2 tiles, at least one of them with an edge inside the cell (or the canvas edge), even geometry: the
                   // lean kernel with border fix-ups of the staged windows and 0 / 1 coverage masks (k_gather_full<.., EDGE>)
       GC_E4 = 7,  // the same with 3 or 4 tiles
       GC_CLASSES = 8 };
// what GC_E2 / GC_E4 ask of every covering tile: even offsets between block grid and tile (cx0 is a multiple of FB_W, cy0 of
// FB_H from row_begin) and a level-1 plane of exactly half the size, at least 2 x 2 (so the pyrUp border rules are the plain
// "column -1 := column 1" and "column n := column n - 1", with no duplicated or truncated last row / column)
__host__ __device__ __forceinline__ bool gather_slot_edge_ok(const GatherTile &t, int cy0, int cx0)
{
    return (((cx0 - t.x) & 1) == 0) && (((cy0 - t.y) & 1) == 0) && t.w == 2 * t.w1 && t.h == 2 * t.h1 && t.w1 >= 2 && t.h1 >= 2;
}

// All shared-memory traffic of k_gather goes through byte offsets into the extern array itself (never through pointers
// kept in registers across the rolled loop): the compiler then knows the address space and emits LDS / STS with 32-bit
// addresses and immediate offsets instead of generic 64-bit loads.
extern __shared__ __align__(128) uint8_t g_gather_smem[];
__device__ __forceinline__ float gs_f32(int off) { return *reinterpret_cast<const float *>(g_gather_smem + off); }
__device__ __forceinline__ float2 gs_f32x2(int off) { return *reinterpret_cast<const float2 *>(g_gather_smem + off); }
__device__ __forceinline__ float gs_u8f(int off) { return __int2float_rn((int)g_gather_smem[off]); }
__device__ __forceinline__ int gs_u16(int off) { return (int)*reinterpret_cast<const unsigned short *>(g_gather_smem + off); }
// horizontal pyrUp of the uint16 G_1 window for columns (2i, 2i+1), i = window centre, UN-NORMALISED and still x 256: exact
// integers (<= 8 * 65280), so the float op order of the reference does not matter here; the 2^-8 is folded into the pixel
// stage (x 2^-8 commutes with every rounding)
__device__ __forceinline__ float2 g16_pair_even_origin(int g0, int g1, int g2)
{
    return make_float2(__int2float_rn(6 * g1 + g0 + g2), __int2float_rn(4 * (g1 + g2)));
}

// The cell whose block range [first, first + nbx * nby) holds block `bid` (cells are sorted by `first`, which starts at
// 0).  Every warp searches for itself, 32 probes at a time: two dependent loads for up to 1024 cells instead of the ten
// of a binary search -- the search is the first thing a CTA does and nothing can start before it ends.
__device__ __forceinline__ int gather_find_cell(const GatherCell *__restrict__ cells, int n, int bid)
{
    if (n <= 1024) {
        const int lane = (int)(threadIdx.x & 31u);
        const int stride = (n + 31) >> 5; // <= 32
        int idx = lane * stride;
        int f = idx < n ? __ldg(&cells[idx].first) : 0x7fffffff;
        unsigned m = __ballot_sync(0xffffffffu, f <= bid); // ascending: a prefix of ones, lane 0 always set
        const int base = (31 - __clz((int)m)) * stride;
        idx = base + lane;
        f = (lane < stride && idx < n) ? __ldg(&cells[idx].first) : 0x7fffffff;
        m = __ballot_sync(0xffffffffu, f <= bid);
        return base + 31 - __clz((int)m);
    }
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(&cells[mid].first) <= bid)
            lo = mid;
        else
            hi = mid - 1;
    }
    return lo;
}

template <int C, int TMAX>
__global__ void __launch_bounds__(FB_PAIRS *C, (TMAX <= 2) ? SRF_GATHER_CTAS : 2)
    k_gather(const __grid_constant__ GatherParams P, const __grid_constant__ GatherTiles GT, const __grid_constant__ FinalTmaps TM)
{
    pdl_enter(); // programmatic dependent launch: nothing above this line may touch global data
    using Stage = GatherStage<C>;
    constexpr int NT = FB_PAIRS * C;
    constexpr int O_SC = (int)offsetof(Stage, sc), O_SG = (int)offsetof(Stage, sg), O_U8 = (int)offsetof(Stage, u8), O_NWY = (int)offsetof(Stage, nwy);
    constexpr int STAGE = (int)sizeof(Stage), O_OUT = TMAX * STAGE;
    constexpr int ROWB = FB_SRC_COLS * 4, CHB = FB_SRC_ROWS * ROWB;  // bytes per staged C_1 row / channel plane
    constexpr int ROWG = FB_G_COLS * 2, CHG = FB_SRC_ROWS * ROWG;    // the same for the uint16 G_1 window
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ FinalGeo s_geo[TMAX];
    const int base = (int)((128u - (smem_u32(g_gather_smem) & 127u)) & 127u); // the TMA boxes want 128-byte aligned destinations

    // ---- which cell, which block ----
    const int bid = blockIdx.x;
    const int lo = gather_find_cell(P.cells, P.n_cells, bid);
    const int4 ca = __ldg(reinterpret_cast<const int4 *>(&P.cells[lo]));
    const int4 cb = __ldg(reinterpret_cast<const int4 *>(&P.cells[lo]) + 1);
    const int4 cc = __ldg(reinterpret_cast<const int4 *>(&P.cells[lo]) + 2);
    const int T = min(cb.y, TMAX);
    const int tiles_of[GC_MAX_T] = {cb.z, cb.w, cc.x, cc.y};
    const int local = bid - cb.x;
    const int by = ca.y + local / ca.z, bx = ca.x + local % ca.z;
    const int cy0 = P.row_begin + by * FB_H, cx0 = bx * FB_W;
    const int cy1 = min(cy0 + FB_H, P.row_end), cx1 = min(cx0 + FB_W, P.out_w);
    const int c = threadIdx.x / FB_PAIRS; // warp-uniform
    const int p = threadIdx.x % FB_PAIRS;

    // ---- stage every covering tile: one thread, one mbarrier ----
    if (threadIdx.x == 0 && T > 0) {
        mbar_init(&s_bar, 1);
        mbar_fence_init();
        mbar_expect_tx(&s_bar, (unsigned)(T * (sizeof(Stage::sc) + sizeof(Stage::sg) + sizeof(Stage::u8))));
    }
    unsigned full_mask = 0; // bit s: slot s is an interior block of even parity (predicate-free arithmetic)
#pragma unroll
    for (int s = 0; s < TMAX; ++s) {
        if (s >= T) break;
        const int ti = tiles_of[s];
        const GatherTile &gt = GT.t[ti];
        const FinalGeo q = gather_geo(gt, cy0, cy1, cx0, cx1, C);
        const int sb = base + s * STAGE;
        if (threadIdx.x == 0) {
            tma_tensor_3d_g2s(g_gather_smem + sb + O_SC, &TM.m[ti][0], q.sx0, q.srow0, 0, &s_bar);
            tma_tensor_3d_g2s(g_gather_smem + sb + O_SG, &TM.m[ti][1], (q.i0 >> 3) * 8, q.srow0, 0, &s_bar);
            tma_tensor_2d_g2s(g_gather_smem + sb + O_U8, &TM.m[ti][2], q.u8_a0 >> 2, cy0 - gt.y, &s_bar);
            s_geo[s] = q;
        }
        if (threadIdx.x >= 32 && threadIdx.x < 32 + FB_H) { // W_0 row profile of the block (negated), second warp
            const int ty = cy0 + (int)threadIdx.x - 32 - gt.y;
            *reinterpret_cast<float *>(g_gather_smem + sb + O_NWY + 4 * ((int)threadIdx.x - 32)) =
                -__ldg(gt.lut + max(min(min(ty, gt.h - 1 - ty), gt.lut_len - 1), 0));
        }
        full_mask |= gather_slot_full(gt, q, cy0, cx0) ? (1u << s) : 0u;
    }
    __syncthreads(); // barrier initialised, geometry and row profiles visible
    if (T > 0) mbar_wait_or_trap(&s_bar, 0);

    const uintptr_t cbase = reinterpret_cast<uintptr_t>(P.canvas) + (size_t)cy0 * P.canvas_pitch + (size_t)cx0 * C;
    const int mis0 = (int)(cbase & 15), pit16 = (int)(P.canvas_pitch & 15);
    const int out_thr = base + O_OUT + 2 * p * C + c; // this thread's first byte inside a staged output row

    // ---- per-slot state: a 3-row ring of horizontally up-sampled (C_1, G_1) pairs, the column weights, smem cursors ----
    float2 rc0[TMAX], rg0[TMAX], rc1[TMAX], rg1[TMAX]; // .x / .y = the thread's two columns
    float wxa[TMAX], wxb[TMAX];                        // column weights; NEGATED for full slots (see k_final's FULL path)
    int o_src[TMAX], o_g[TMAX], o_u8[TMAX];            // byte offsets: next C_1 / G_1 source row (full slots), the thread's first pixel byte
    unsigned flags[TMAX] = {};                         // generic slots: bit 0 ina, 1 inb, 2 spa, 3 spb

    // `o` / `oh` address the first of the three C_1 / G_1 samples (byte offsets inside the stage's sc / sg arrays).  G values
    // are x 256 (see g16_pair_even_origin) all the way to the pixel stage.
    auto hrow_fast = [&](int sb, int o, int oh, float2 &oc, float2 &og) { // interior form: columns (2i, 2i+1), i = window centre
        const float c0 = gs_f32(sb + O_SC + o), c1 = gs_f32(sb + O_SC + o + 4), c2 = gs_f32(sb + O_SC + o + 8);
        oc = make_float2(up_even(c0, c1, c2), up_odd(c1, c2));
        og = g16_pair_even_origin(gs_u16(sb + O_SG + oh), gs_u16(sb + O_SG + oh + 2), gs_u16(sb + O_SG + oh + 4));
    };
    auto hrow_generic = [&](int s, int r_logical, float2 &oc, float2 &og) {
        const GatherTile &gt = GT.t[tiles_of[s]];
        const FinalGeo &q = s_geo[s];
        const int h1 = gt.h1, w1 = gt.w1;
        int r = (r_logical == -1) ? min(1, h1 - 1) : max(min(r_logical, h1 - 1), 0);
        r = max(min(r - q.srow0, q.n_srows - 1), 0);
        const int sb = base + s * STAGE;
        const int row = c * CHB + r * ROWB; // byte offset of the staged C_1 row inside sc
        const int o = row + 4 * (q.i0 + p - q.sx0);
        const int sxg = (q.i0 >> 3) * 8;    // first staged column of the G_1 window
        const int rowg = sb + O_SG + c * CHG + r * ROWG;
        // integer-valued floats below 2^24: every pyrUp formula is exact on them, so the float forms serve the uint16 level too
        auto g16 = [&](int col) { return (float)gs_u16(rowg + 2 * (col - sxg)); };
        const int i0p = q.i0 + p;
        const HPair hc = h_pair_fast(gs_f32(sb + O_SC + o), gs_f32(sb + O_SC + o + 4), gs_f32(sb + O_SC + o + 8), q.xpar);
        const HPair hg = h_pair_fast(g16(i0p), g16(i0p + 1), g16(i0p + 2), q.xpar);
        oc = make_float2(hc.a, hc.b);
        og = make_float2(hg.a, hg.b);
        if (flags[s] & 12u) { // edge formulas (tile column 0, the last source column, a 1-wide source)
            const int txa = cx0 - gt.x + 2 * p, txb = txa + 1;
            auto atc = [&](int ii) { return gs_f32(sb + O_SC + row + 4 * (ii - q.sx0)); };
            auto atg = [&](int ii) { return g16(ii); };
            if (flags[s] & 4u) {
                oc.x = up_h_sample(txa, w1, atc);
                og.x = up_h_sample(txa, w1, atg);
            }
            if (flags[s] & 8u) {
                oc.y = up_h_sample(txb, w1, atc);
                og.y = up_h_sample(txb, w1, atg);
            }
        }
    };
#pragma unroll
    for (int s = 0; s < TMAX; ++s) {
        if (s >= T) break;
        const GatherTile &gt = GT.t[tiles_of[s]];
        const FinalGeo &q = s_geo[s];
        const int txa = cx0 - gt.x + 2 * p, txb = txa + 1;
        const int lmax = gt.lut_len - 1;
        const float wa = __ldg(gt.lut + max(min(min(txa, gt.w - 1 - txa), lmax), 0)), wb = __ldg(gt.lut + max(min(min(txb, gt.w - 1 - txb), lmax), 0));
        o_u8[s] = base + s * STAGE + O_U8 + txa * C + c - q.u8_a0;
        if (full_mask & (1u << s)) {
            wxa[s] = -wa;
            wxb[s] = -wb;
            flags[s] = 0;
            o_src[s] = c * CHB + 4 * (q.i0 + p - q.sx0); // relative to the stage's sc array; staged row 0 is logical row ws0
            o_g[s] = c * CHG + 2 * (q.i0 + p - (q.i0 >> 3) * 8);
            hrow_fast(base + s * STAGE, o_src[s], o_g[s], rc0[s], rg0[s]);
            hrow_fast(base + s * STAGE, o_src[s] + ROWB, o_g[s] + ROWG, rc1[s], rg1[s]);
            o_src[s] += 2 * ROWB;
            o_g[s] += 2 * ROWG;
        } else {
            wxa[s] = wa;
            wxb[s] = wb;
            const bool ina = (cx0 + 2 * p >= q.ix0) && (cx0 + 2 * p < q.ix1), inb = (cx0 + 2 * p + 1 >= q.ix0) && (cx0 + 2 * p + 1 < q.ix1);
            const bool spa = ina && ((txa <= 0) || ((txa >> 1) >= gt.w1 - 1)), spb = inb && ((txb <= 0) || ((txb >> 1) >= gt.w1 - 1));
            flags[s] = (ina ? 1u : 0u) | (inb ? 2u : 0u) | (spa ? 4u : 0u) | (spb ? 8u : 0u);
            o_src[s] = o_g[s] = 0;
            hrow_generic(s, q.ws0, rc0[s], rg0[s]);
            hrow_generic(s, q.ws0 + 1, rc1[s], rg1[s]);
        }
    }
    auto up_even2 = [](float2 m, float2 p6, float2 n) {
        return __fadd2_rn(__fadd2_rn(make_float2(__fmul_rn(p6.x, 6.f), __fmul_rn(p6.y, 6.f)), m), n);
    };
    auto up_odd2 = [](float2 p4, float2 n) { return __fmul2_rn(__fadd2_rn(p4, n), make_float2(4.f, 4.f)); };

#pragma unroll GA_UNROLL
    for (int k = 0; k < FB_STEPS; ++k) {
        float2 acc[2], ws[2]; // the two finished rows: sum of C_0 (x64) and sum of W_0 over the covering tiles, in list order
        if (T == 0) acc[0] = acc[1] = ws[0] = ws[1] = make_float2(0.f, 0.f);
#pragma unroll
        for (int s = 0; s < TMAX; ++s) {
            if (s >= T) break;
            const int sb = base + s * STAGE;
            float2 v[2], w[2]; // this tile's C_0 (x64) and W_0 for the two rows
            if (full_mask & (1u << s)) {
                // interior block of even parity (k_final's FULL path): no predicates, packed FP32.  ptxas contracts a packed
                // multiply into a following packed add even for .rn, so every INEXACT product that feeds an add (6 p, d * w)
                // is a scalar FMUL.  nd = fma(g, -64, up64G) = -(64 g - up64G), nw = -w0: their product is the reference's.
                float2 rc2, rg2;
                hrow_fast(sb, o_src[s], o_g[s], rc2, rg2);
                o_src[s] += ROWB;
                o_g[s] += ROWG;
                float2 uc[2], ug[2]; // x64 (ug also x256), tile rows (2j, 2j+1): even uses the whole ring, odd its last two rows
                uc[0] = up_even2(rc0[s], rc1[s], rc2); ug[0] = up_even2(rg0[s], rg1[s], rg2);
                uc[1] = up_odd2(rc1[s], rc2);          ug[1] = up_odd2(rg1[s], rg2);
                const float2 nwy2 = gs_f32x2(sb + O_NWY + 8 * k);
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                    const float nwy = a ? nwy2.y : nwy2.x;
                    const int ob = o_u8[s] + (2 * k + a) * FB_U8_STRIDE;
                    const float2 nw = make_float2(fmaxf(nwy, wxa[s]), fmaxf(nwy, wxb[s])); // -min(wy, wx)
                    // nd = -256 (64 g - up64G): integers below 2^24, exact; dw = 256 x the reference's product (the scale commutes
                    // with its rounding); the last FMA takes the 2^-8 out again: fma(dw, 2^-8, uc) == uc + dw / 256 rounded once
                    const float2 nd = __ffma2_rn(make_float2(gs_u8f(ob), gs_u8f(ob + C)), make_float2(-16384.f, -16384.f), ug[a]);
                    const float2 dw = make_float2(__fmul_rn(nd.x, nw.x), __fmul_rn(nd.y, nw.y));
                    v[a] = __ffma2_rn(dw, make_float2(1.f / 256.f, 1.f / 256.f), uc[a]);
                    w[a] = make_float2(-nw.x, -nw.y);
                }
                rc0[s] = rc1[s]; rg0[s] = rg1[s]; rc1[s] = rc2; rg1[s] = rg2;
            } else {
                // generic block: tile edges, canvas edges, odd tile origins (k_final's generic path).  Pixels the tile does not
                // cover contribute +0 to both sums, which changes nothing.
                const FinalGeo &q = s_geo[s];
                float2 rc2, rg2;
                hrow_generic(s, q.ws0 + k + 2, rc2, rg2);
                float2 uc[2], ug[2]; // x64
                if (q.ypar == 0) { // tile rows (2j, 2j+1), j = ws0+k+1: even uses (m,p,n), odd uses (p,n)
                    uc[0] = make_float2(up_even(rc0[s].x, rc1[s].x, rc2.x), up_even(rc0[s].y, rc1[s].y, rc2.y));
                    ug[0] = make_float2(up_even(rg0[s].x, rg1[s].x, rg2.x), up_even(rg0[s].y, rg1[s].y, rg2.y));
                    uc[1] = make_float2(up_odd(rc1[s].x, rc2.x), up_odd(rc1[s].y, rc2.y));
                    ug[1] = make_float2(up_odd(rg1[s].x, rg2.x), up_odd(rg1[s].y, rg2.y));
                } else {           // tile rows (2j+1, 2j+2), j = ws0+k: odd uses ring 0,1; even uses ring 0,1,2
                    uc[0] = make_float2(up_odd(rc0[s].x, rc1[s].x), up_odd(rc0[s].y, rc1[s].y));
                    ug[0] = make_float2(up_odd(rg0[s].x, rg1[s].x), up_odd(rg0[s].y, rg1[s].y));
                    uc[1] = make_float2(up_even(rc0[s].x, rc1[s].x, rc2.x), up_even(rc0[s].y, rc1[s].y, rc2.y));
                    ug[1] = make_float2(up_even(rg0[s].x, rg1[s].x, rg2.x), up_even(rg0[s].y, rg1[s].y, rg2.y));
                }
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                    const int r = 2 * k + a, cy = cy0 + r;
                    const bool row_in = cy >= q.iy0 && cy < q.iy1; // uniform
                    const float wy = -gs_f32(sb + O_NWY + 4 * r);
                    const int ob = o_u8[s] + r * FB_U8_STRIDE;
                    v[a] = w[a] = make_float2(0.f, 0.f);
                    if (row_in && (flags[s] & 1u)) {
                        const float w0 = fminf(wy, wxa[s]);
                        // fma(g, 16384, -256 up64G) == 256 (64 g - up64G), exact; the product carries the x 256 through its rounding
                        v[a].x = __fmaf_rn(__fmul_rn(__fmaf_rn(gs_u8f(ob), 16384.f, -ug[a].x), w0), 1.f / 256.f, uc[a].x);
                        w[a].x = w0;
                    }
                    if (row_in && (flags[s] & 2u)) {
                        const float w0 = fminf(wy, wxb[s]);
                        v[a].y = __fmaf_rn(__fmul_rn(__fmaf_rn(gs_u8f(ob + C), 16384.f, -ug[a].y), w0), 1.f / 256.f, uc[a].y);
                        w[a].y = w0;
                    }
                }
                rc0[s] = rc1[s]; rg0[s] = rg1[s]; rc1[s] = rc2; rg1[s] = rg2;
            }
            if (s == 0) { // the reference's accumulators start at zero: 0 + v differs from v only in the sign of a zero
                acc[0] = v[0]; acc[1] = v[1]; ws[0] = w[0]; ws[1] = w[1];
            } else {
                acc[0] = __fadd2_rn(acc[0], v[0]);
                acc[1] = __fadd2_rn(acc[1], v[1]);
                ws[0].x = __fadd_rn(ws[0].x, w[0].x); ws[0].y = __fadd_rn(ws[0].y, w[0].y);
                ws[1].x = __fadd_rn(ws[1].x, w[1].x); ws[1].y = __fadd_rn(ws[1].y, w[1].y);
            }
        }
        // ---- normalise, clip, truncate the two finished rows (k_final's epilogue: why the guarded reciprocal is exact) ----
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const int r = 2 * k + a;
            const int mis = (mis0 + r * pit16) & 15;
            const float b0 = fmaxf(ws[a].x, 1e-6f), b1 = fmaxf(ws[a].y, 1e-6f);
            const float a0 = __fmul_rn(acc[a].x, 1.f / 64.f), a1 = __fmul_rn(acc[a].y, 1.f / 64.f);
            float q0 = a0 * rcp_approx(b0), q1 = a1 * rcp_approx(b1);
            const bool n0 = fabsf(q0 - rintf(q0)) < FINAL_NEAR && (__float_as_uint(b0) & 0x007FFFFFu) != 0u;
            const bool n1 = fabsf(q1 - rintf(q1)) < FINAL_NEAR && (__float_as_uint(b1) & 0x007FFFFFu) != 0u;
            if (n0 || n1) {
                q0 = __fdiv_rn(a0, b0);
                q1 = __fdiv_rn(a1, b1);
            }
            const int od = out_thr + r * FB_U8_STRIDE + mis; // columns beyond the canvas land in the row's slack, never copied out
            g_gather_smem[od] = (uint8_t)f2u8_trunc_sat(q0);
            g_gather_smem[od + C] = (uint8_t)f2u8_trunc_sat(q1);
        }
    }
    __syncthreads();

    // ---- copy the staged rows out: aligned 16-byte chunks; a row's ragged ends go as 4-byte words where they can ----
    const int n_rows = cy1 - cy0, row_bytes = (cx1 - cx0) * C;
    constexpr int kChunks = (FB_W * 3 + 15 + 15) / 16;
    for (int i = threadIdx.x; i < n_rows * kChunks; i += NT) {
        const int r = i / kChunks, ch = i - r * kChunks;
        uint8_t *grow = P.canvas + (size_t)(cy0 + r) * P.canvas_pitch + (size_t)cx0 * C;
        const int mis = (int)(reinterpret_cast<uintptr_t>(grow) & 15);
        const int lo_b = ch * 16, hi_b = lo_b + 16; // byte range inside the staged row
        const int p0 = mis, p1 = mis + row_bytes;   // payload range
        if (hi_b <= p0 || lo_b >= p1) continue;
        uint8_t *gdst = grow - mis + lo_b; // 16-byte aligned
        const int so = base + O_OUT + r * FB_U8_STRIDE + lo_b;
        if (lo_b >= p0 && hi_b <= p1) {
            *reinterpret_cast<uint4 *>(gdst) = *reinterpret_cast<const uint4 *>(g_gather_smem + so);
        } else {
#pragma unroll
            for (int wd = 0; wd < 4; ++wd) {
                const int a = lo_b + 4 * wd;
                if (a >= p0 && a + 4 <= p1) {
                    *reinterpret_cast<uint32_t *>(gdst + 4 * wd) = *reinterpret_cast<const uint32_t *>(g_gather_smem + so + 4 * wd);
                } else if (a + 4 > p0 && a < p1) {
                    for (int e = max(a, p0); e < min(a + 4, p1); ++e) gdst[e - lo_b] = g_gather_smem[so + e - lo_b];
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// k_gather_full: the lean form of k_gather for cells the host has PROVEN interior for every covering tile (classes GC_U1,
// GC_F1, GC_F2, GC_F4): no coverage predicates, no edge formulas, no per-slot dispatch -- T is a template parameter
// (TMAX = 4 serves T = 3 and 4).  UNIT (class GC_U1): the tile's weight is 1.0 over the whole block, so C_0 needs no weight
// products, the sum of W_0 is exactly 1 and the quotient is acc itself (a power-of-two divisor: see k_final's epilogue) --
// the epilogue is a scale, a truncating conversion and a byte store.
// ------------------------------------------------------------------------------------------------------------
template <int C, int TMAX, bool UNIT, bool EDGE = false>
__global__ void __launch_bounds__(FB_PAIRS *C, (TMAX == 1) ? 6 : ((TMAX == 2) ? 4 : 2))
    k_gather_full(const __grid_constant__ GatherParams P, const __grid_constant__ GatherTiles GT, const __grid_constant__ FinalTmaps TM)
{
    pdl_enter(); // programmatic dependent launch: nothing above this line may touch global data
    using Stage = GatherStage<C>;
    constexpr int NT = FB_PAIRS * C;
    constexpr int O_SC = (int)offsetof(Stage, sc), O_SG = (int)offsetof(Stage, sg), O_U8 = (int)offsetof(Stage, u8), O_NWY = (int)offsetof(Stage, nwy);
    constexpr int STAGE = (int)sizeof(Stage), O_OUT = TMAX * STAGE;
    constexpr int ROWB = FB_SRC_COLS * 4, CHB = FB_SRC_ROWS * ROWB; // bytes per staged C_1 row / channel plane
    constexpr int ROWG = FB_G_COLS * 2, CHG = FB_SRC_ROWS * ROWG;   // the same for the uint16 G_1 window
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ __align__(8) int s_rowoff[FB_H]; // r * FB_U8_STRIDE + misalignment of canvas row r: where a finished row is staged
    __shared__ __align__(8) unsigned s_mrow[EDGE ? TMAX : 1][FB_H]; // EDGE: all ones where the tile covers the block's row, else 0
    const int base = (int)((128u - (smem_u32(g_gather_smem) & 127u)) & 127u);

    const int bid = blockIdx.x;
    const int lo = gather_find_cell(P.cells, P.n_cells, bid);
    const int4 ca = __ldg(reinterpret_cast<const int4 *>(&P.cells[lo]));
    const int4 cb = __ldg(reinterpret_cast<const int4 *>(&P.cells[lo]) + 1);
    const int4 cc = __ldg(reinterpret_cast<const int4 *>(&P.cells[lo]) + 2);
    const int T = (TMAX <= 2 && !EDGE) ? TMAX : min(cb.y, TMAX);
    const int tiles_of[GC_MAX_T] = {cb.z, cb.w, cc.x, cc.y};
    const int local = bid - cb.x;
    const int by = ca.y + local / ca.z, bx = ca.x + local % ca.z;
    const int cy0 = P.row_begin + by * FB_H, cx0 = bx * FB_W; // the block lies inside the canvas and the band (it is interior to a tile)
    const int c = threadIdx.x / FB_PAIRS, p = threadIdx.x % FB_PAIRS;

    const uintptr_t cbase = reinterpret_cast<uintptr_t>(P.canvas) + (size_t)cy0 * P.canvas_pitch + (size_t)cx0 * C;
    const int mis0 = (int)(cbase & 15), pit16 = (int)(P.canvas_pitch & 15);
    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        mbar_fence_init();
        mbar_expect_tx(&s_bar, (unsigned)(T * (sizeof(Stage::sc) + sizeof(Stage::sg) + sizeof(Stage::u8))));
    }
    if (threadIdx.x >= NT - FB_H) { // the CTA's last 16 threads (a 1-channel CTA has only 64)
        const int r = (int)threadIdx.x - (NT - FB_H);
        s_rowoff[r] = r * FB_U8_STRIDE + ((mis0 + r * pit16) & 15);
    }
    int o_src[TMAX], o_g[TMAX], o_u8[TMAX];
    float nwxa[TMAX], nwxb[TMAX];
    unsigned mcola[EDGE ? TMAX : 1], mcolb[EDGE ? TMAX : 1]; // EDGE: all ones / 0 for the thread's two columns
#pragma unroll
    for (int s = 0; s < TMAX; ++s) {
        if (s >= T) break;
        const int ti = tiles_of[s];
        const GatherTile &gt = GT.t[ti];
        // even; >= 2 unless EDGE, where they may be negative (the tile starts inside the block: the TMA boxes then start at
        // negative coordinates and everything outside the planes arrives as zeros) or reach past the tile's far edge
        const int tyf = cy0 - gt.y, txf = cx0 - gt.x;
        const int i0 = (txf >> 1) - 1, ws0 = (tyf >> 1) - 1, sx0 = (i0 >> 2) * 4, u8_a0 = (txf * C) & ~15;
        const int sb = base + s * STAGE;
        if (threadIdx.x == 0) {
            tma_tensor_3d_g2s(g_gather_smem + sb + O_SC, &TM.m[ti][0], sx0, ws0, 0, &s_bar);
            tma_tensor_3d_g2s(g_gather_smem + sb + O_SG, &TM.m[ti][1], (i0 >> 3) * 8, ws0, 0, &s_bar);
            tma_tensor_2d_g2s(g_gather_smem + sb + O_U8, &TM.m[ti][2], u8_a0 >> 2, tyf, &s_bar);
        }
        const int txa = txf + 2 * p;
        o_src[s] = c * CHB + 4 * (i0 + p - sx0); // first C_1 sample, relative to the stage's sc array
        o_g[s] = c * CHG + 2 * (i0 + p - (i0 >> 3) * 8); // first G_1 sample, relative to sg
        o_u8[s] = sb + O_U8 + txa * C + c - u8_a0;
        if (!UNIT) {
            if (threadIdx.x >= 32 && threadIdx.x < 32 + FB_H) { // -W_0 row profile of the block
                const int ty = tyf + (int)threadIdx.x - 32;
                float nwy = -__ldg(gt.lut + max(min(min(ty, gt.h - 1 - ty), gt.lut_len - 1), 0));
                if (EDGE) { // rows the tile does not cover: weight 0 and mask 0
                    const bool in = ty >= 0 && ty < gt.h;
                    nwy = in ? nwy : 0.f;
                    s_mrow[EDGE ? s : 0][(int)threadIdx.x - 32] = in ? 0xFFFFFFFFu : 0u;
                }
                *reinterpret_cast<float *>(g_gather_smem + sb + O_NWY + 4 * ((int)threadIdx.x - 32)) = nwy;
            }
            const int lmax = gt.lut_len - 1, txb = txa + 1;
            nwxa[s] = -__ldg(gt.lut + max(min(min(txa, gt.w - 1 - txa), lmax), 0));
            nwxb[s] = -__ldg(gt.lut + max(min(min(txb, gt.w - 1 - txb), lmax), 0));
            if (EDGE) {
                const bool ina = txa >= 0 && txa < gt.w, inb = txb >= 0 && txb < gt.w;
                nwxa[s] = ina ? nwxa[s] : 0.f;
                nwxb[s] = inb ? nwxb[s] : 0.f;
                mcola[EDGE ? s : 0] = ina ? 0xFFFFFFFFu : 0u;
                mcolb[EDGE ? s : 0] = inb ? 0xFFFFFFFFu : 0u;
            }
        }
    }
    __syncthreads(); // barrier initialised; row profiles and staging offsets visible
    mbar_wait_or_trap(&s_bar, 0);
    if (EDGE) {
        // pyrUp's border rules, applied to the staged level-1 windows in place: source column -1 := column 1 and column
        // w1 := column w1 - 1 (cv2.pyrUp: reflect-101 at the low edge, replicate at the high edge), then the same for rows
        // (after the columns, so the corners come out as the separable rule has them).  Everything else outside the planes
        // is TMA zero fill or plane padding, and only ever lies under pixels whose mask is 0.
        bool fix_cols = false, fix_rows = false;
#pragma unroll
        for (int s = 0; s < TMAX; ++s) {
            if (s >= T) break;
            const GatherTile &gt = GT.t[tiles_of[s]];
            const int i0 = ((cx0 - gt.x) >> 1) - 1, ws0 = ((cy0 - gt.y) >> 1) - 1;
            fix_cols |= (i0 <= -1) || (i0 + FB_PAIRS + 1 >= gt.w1);
            fix_rows |= (ws0 <= -1) || (ws0 + FB_SRC_ROWS - 1 >= gt.h1);
            // (the second halves of the need* tests below only keep the indices inside the windows if a listed tile should
            // ever miss the block altogether -- its masks are all 0 then)
        }
        if (fix_cols) { // block-uniform
#pragma unroll
            for (int s = 0; s < TMAX; ++s) {
                if (s >= T) break;
                const GatherTile &gt = GT.t[tiles_of[s]];
                const int i0 = ((cx0 - gt.x) >> 1) - 1, sx0 = (i0 >> 2) * 4, gx0 = (i0 >> 3) * 8;
                const bool needl = i0 <= -1 && i0 >= -(FB_PAIRS + 1), needr = i0 + FB_PAIRS + 1 >= gt.w1 && i0 <= gt.w1 - 1;
                const int sb = base + s * STAGE;
                for (int i = threadIdx.x; i < C * FB_SRC_ROWS; i += NT) {
                    float *rowc = reinterpret_cast<float *>(g_gather_smem + sb + O_SC + i * ROWB); // planes are contiguous: i = c * rows + r
                    unsigned short *rowg = reinterpret_cast<unsigned short *>(g_gather_smem + sb + O_SG + i * ROWG);
                    if (needl) {
                        rowc[-1 - sx0] = rowc[1 - sx0];
                        rowg[-1 - gx0] = rowg[1 - gx0];
                    }
                    if (needr) {
                        rowc[gt.w1 - sx0] = rowc[gt.w1 - 1 - sx0];
                        rowg[gt.w1 - gx0] = rowg[gt.w1 - 1 - gx0];
                    }
                }
            }
            __syncthreads();
        }
        if (fix_rows) {
#pragma unroll
            for (int s = 0; s < TMAX; ++s) {
                if (s >= T) break;
                const GatherTile &gt = GT.t[tiles_of[s]];
                const int ws0 = ((cy0 - gt.y) >> 1) - 1;
                const bool needt = ws0 <= -1 && ws0 >= -(FB_SRC_ROWS - 2), needb = ws0 + FB_SRC_ROWS - 1 >= gt.h1 && ws0 <= gt.h1 - 1;
                const int sb = base + s * STAGE;
                constexpr int FIX_COLS = FB_G_COLS > FB_SRC_COLS ? FB_G_COLS : FB_SRC_COLS; // the wider of the two windows
                for (int i = threadIdx.x; i < C * FIX_COLS; i += NT) {
                    const int ch = i / FIX_COLS, col = i - ch * FIX_COLS;
                    float *pc = reinterpret_cast<float *>(g_gather_smem + sb + O_SC + ch * CHB) + col;
                    unsigned short *pg = reinterpret_cast<unsigned short *>(g_gather_smem + sb + O_SG + ch * CHG) + col;
                    const bool c_col = col < FB_SRC_COLS, g_col = col < FB_G_COLS;
                    if (needt) {
                        if (c_col) pc[(-1 - ws0) * FB_SRC_COLS] = pc[(1 - ws0) * FB_SRC_COLS];
                        if (g_col) pg[(-1 - ws0) * FB_G_COLS] = pg[(1 - ws0) * FB_G_COLS];
                    }
                    if (needb) {
                        if (c_col) pc[(gt.h1 - ws0) * FB_SRC_COLS] = pc[(gt.h1 - 1 - ws0) * FB_SRC_COLS];
                        if (g_col) pg[(gt.h1 - ws0) * FB_G_COLS] = pg[(gt.h1 - 1 - ws0) * FB_G_COLS];
                    }
                }
            }
            __syncthreads();
        }
    }

    float2 rc0[TMAX], rg0[TMAX], rc1[TMAX], rg1[TMAX];
    auto hrow = [&](int sb, int o, int oh, float2 &oc, float2 &og) { // columns (2i, 2i+1), i = window centre: even form, odd form
        const float c0 = gs_f32(sb + O_SC + o), c1 = gs_f32(sb + O_SC + o + 4), c2 = gs_f32(sb + O_SC + o + 8);
        oc = make_float2(up_even(c0, c1, c2), up_odd(c1, c2));
        og = g16_pair_even_origin(gs_u16(sb + O_SG + oh), gs_u16(sb + O_SG + oh + 2), gs_u16(sb + O_SG + oh + 4)); // x 256, exact
    };
#pragma unroll
    for (int s = 0; s < TMAX; ++s) {
        if (s >= T) break;
        hrow(base + s * STAGE, o_src[s], o_g[s], rc0[s], rg0[s]);
        hrow(base + s * STAGE, o_src[s] + ROWB, o_g[s] + ROWG, rc1[s], rg1[s]);
        o_src[s] += 2 * ROWB;
        o_g[s] += 2 * ROWG;
    }
    auto up_even2 = [](float2 m, float2 p6, float2 n) {
        return __fadd2_rn(__fadd2_rn(make_float2(__fmul_rn(p6.x, 6.f), __fmul_rn(p6.y, 6.f)), m), n);
    };
    auto up_odd2 = [](float2 p4, float2 n) { return __fmul2_rn(__fadd2_rn(p4, n), make_float2(4.f, 4.f)); };
    int out_thr = base + O_OUT + 2 * p * C + c;
    asm volatile("" : "+r"(out_thr)); // keep it in a register: rematerialising it from the thread index costs more than it saves

#pragma unroll GA_UNROLL
    for (int k = 0; k < FB_STEPS; ++k) {
        float2 acc[2], ws[2];
#pragma unroll
        for (int s = 0; s < TMAX; ++s) {
            if (s >= T) break;
            float2 rc2, rg2;
            hrow(base + s * STAGE, o_src[s], o_g[s], rc2, rg2);
            o_src[s] += ROWB;
            o_g[s] += ROWG;
            float2 uc[2], ug[2]; // x64 (ug also x256), tile rows (2j, 2j+1): even uses the whole ring, odd its last two rows
            uc[0] = up_even2(rc0[s], rc1[s], rc2); ug[0] = up_even2(rg0[s], rg1[s], rg2);
            uc[1] = up_odd2(rc1[s], rc2);          ug[1] = up_odd2(rg1[s], rg2);
            float2 v[2], w[2];
            float2 nwy2 = make_float2(0.f, 0.f);
            uint2 mrow2 = make_uint2(0u, 0u);
            if (!UNIT) nwy2 = gs_f32x2(base + s * STAGE + O_NWY + 8 * k);
            if (EDGE) mrow2 = *reinterpret_cast<const uint2 *>(&s_mrow[EDGE ? s : 0][2 * k]);
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                const int ob = o_u8[s] + (2 * k + a) * FB_U8_STRIDE;
                // nd = -256 (64 g - up64G): integers below 2^24, exact -- the reference's subtraction, scaled by a power of two
                const float2 nd = __ffma2_rn(make_float2(gs_u8f(ob), gs_u8f(ob + C)), make_float2(-16384.f, -16384.f), ug[a]);
                if (UNIT) {
                    v[a] = __ffma2_rn(nd, make_float2(-1.f / 256.f, -1.f / 256.f), uc[a]); // (64 g - up64G) * 1.0f is exact: uc + that, rounded once
                } else {
                    const float nwy = a ? nwy2.y : nwy2.x;
                    const float2 nw = make_float2(fmaxf(nwy, nwxa[s]), fmaxf(nwy, nwxb[s])); // -min(wy, wx)
                    // dw = 256 x the reference's product (the scale commutes with its rounding); scalar so it cannot fuse into an add
                    const float2 dw = make_float2(__fmul_rn(nd.x, nw.x), __fmul_rn(nd.y, nw.y));
                    v[a] = __ffma2_rn(dw, make_float2(1.f / 256.f, 1.f / 256.f), uc[a]); // uc + dw / 256, rounded once
                    w[a] = make_float2(-nw.x, -nw.y);
                    if (EDGE) { // pixels the tile does not cover add exactly +0 to both sums (their weight is already 0), whatever
                                // lies under them in the staged windows (plane padding is not zero-filled and need not be finite)
                        const unsigned mr = a ? mrow2.y : mrow2.x;
                        v[a].x = __uint_as_float(__float_as_uint(v[a].x) & mcola[EDGE ? s : 0] & mr);
                        v[a].y = __uint_as_float(__float_as_uint(v[a].y) & mcolb[EDGE ? s : 0] & mr);
                    }
                }
            }
            rc0[s] = rc1[s]; rg0[s] = rg1[s]; rc1[s] = rc2; rg1[s] = rg2;
            if (s == 0) { // the reference's accumulators start at zero: 0 + v differs from v only in the sign of a zero
                acc[0] = v[0]; acc[1] = v[1];
                if (!UNIT) { ws[0] = w[0]; ws[1] = w[1]; }
            } else {
                acc[0] = __fadd2_rn(acc[0], v[0]);
                acc[1] = __fadd2_rn(acc[1], v[1]);
                ws[0].x = __fadd_rn(ws[0].x, w[0].x); ws[0].y = __fadd_rn(ws[0].y, w[0].y);
                ws[1].x = __fadd_rn(ws[1].x, w[1].x); ws[1].y = __fadd_rn(ws[1].y, w[1].y);
            }
        }
        const int2 ro = *reinterpret_cast<const int2 *>(&s_rowoff[2 * k]);
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const int od = out_thr + (a ? ro.y : ro.x);
            float2 q = __fmul2_rn(acc[a], make_float2(1.f / 64.f, 1.f / 64.f)); // exact (power of two)
            if (!UNIT) {
                const float2 a64 = q;
                const float b0 = fmaxf(ws[a].x, 1e-6f), b1 = fmaxf(ws[a].y, 1e-6f);
                q = __fmul2_rn(a64, make_float2(rcp_approx(b0), rcp_approx(b1)));
                const float2 d = __fadd2_rn(q, make_float2(-rintf(q.x), -rintf(q.y)));
                const bool n0 = fabsf(d.x) < FINAL_NEAR && (__float_as_uint(b0) & 0x007FFFFFu) != 0u;
                const bool n1 = fabsf(d.y) < FINAL_NEAR && (__float_as_uint(b1) & 0x007FFFFFu) != 0u;
                if (n0 || n1) {
                    q.x = __fdiv_rn(a64.x, b0);
                    q.y = __fdiv_rn(a64.y, b1);
                }
            }
            g_gather_smem[od] = (uint8_t)f2u8_trunc_sat(q.x);
            g_gather_smem[od + C] = (uint8_t)f2u8_trunc_sat(q.y);
        }
    }
    __syncthreads();

    // ---- copy the finished rows out (unless EDGE the block is whole: 16 rows of FB_W * C payload bytes) ----
    const int n_rows = EDGE ? min(FB_H, P.row_end - cy0) : FB_H;
    const int ROW_BYTES = EDGE ? min(FB_W, P.out_w - cx0) * C : FB_W * C;
    constexpr int kChunks = (FB_W * C + 15 + 15) / 16;
    for (int i = threadIdx.x; i < n_rows * kChunks; i += NT) {
        const int r = i / kChunks, ch = i - r * kChunks;
        uint8_t *grow = P.canvas + (size_t)(cy0 + r) * P.canvas_pitch + (size_t)cx0 * C;
        const int mis = (int)(reinterpret_cast<uintptr_t>(grow) & 15);
        const int lo_b = ch * 16, hi_b = lo_b + 16;
        const int p0 = mis, p1 = mis + ROW_BYTES;
        if (hi_b <= p0 || lo_b >= p1) continue;
        uint8_t *gdst = grow - mis + lo_b; // 16-byte aligned
        const int so = base + O_OUT + r * FB_U8_STRIDE + lo_b;
        if (lo_b >= p0 && hi_b <= p1) {
            *reinterpret_cast<uint4 *>(gdst) = *reinterpret_cast<const uint4 *>(g_gather_smem + so);
        } else {
#pragma unroll
            for (int wd = 0; wd < 4; ++wd) {
                const int a = lo_b + 4 * wd;
                if (a >= p0 && a + 4 <= p1) {
                    *reinterpret_cast<uint32_t *>(gdst + 4 * wd) = *reinterpret_cast<const uint32_t *>(g_gather_smem + so + 4 * wd);
                } else if (a + 4 > p0 && a < p1) {
                    for (int e = max(a, p0); e < min(a + 4, p1); ++e) gdst[e - lo_b] = g_gather_smem[so + e - lo_b];
                }
            }
        }
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (libsrfuse does not link libcuda)
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled srf_tensor_map_encoder()
{
    static PFN_encodeTiled fn = []() -> PFN_encodeTiled {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) {
            (void)cudaGetLastError();
            return nullptr;
        }
        return (PFN_encodeTiled)p;
    }();
    return fn;
}

// k_down_f32s streams DOWN rows_per_cta output rows per CTA.  32 rows amortise the 3-row vertical halo; a small level (or a
// thin band of a big one) would then be a handful of long-running CTAs on a mostly idle GPU, so the rows are cut until
// the grid fills the SMs a few times over -- the level then costs one short CTA lifetime instead of a long one
static int df_rows_per_cta(const srf_ctx *ctx, int ctas_per_row_block, int rows)
{
    int rpc = DF_ROWS;
    while (rpc > DF_BATCH && (long long)ctas_per_row_block * srf_ceil_div(rows, rpc) < 4LL * ctx->sm_count) rpc /= 2;
    return rpc;
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
struct HostGroup {
    int h, w, lut;
    std::vector<int> tiles;
    GroupDev dev;
};

struct RowRange {
    int a, b;
    void merge(int lo, int hi)
    {
        if (lo >= hi) return;
        if (a >= b) {
            a = lo;
            b = hi;
        } else {
            a = std::min(a, lo);
            b = std::max(b, hi);
        }
    }
};

} // namespace

// Which rows (or columns) of every pyramid level does an interval [a0, b0) of level-0 outputs depend on?
//   up pass  : C_{l-1}[r] reads C_l, G_l at r/2 - 1 .. r/2 + 1 (row -1 is the reflection of row 1)
//   down pass: G_l[r] reads G_{l-1} at 2r-2 .. 2r+2 (reflect-101 at both ends)
// needG[0] is what must be present of the INPUT tile.  dim[l] = extent of level l.  Exact (a dependency cone), so
// restricting work or uploads to these ranges cannot change a needed output.
static void srf_propagate_need(int a0, int b0, const int *dim, int n, RowRange *needC, RowRange *needG)
{
    for (int l = 0; l < SRF_MAX_LEVELS; ++l) needC[l] = needG[l] = RowRange{0, 0};
    if (a0 >= b0) return;
    needC[0].merge(a0, b0);
    needG[0].merge(a0, b0);
    for (int l = 1; l < n; ++l) {
        const int h1 = dim[l];
        const int lo = std::max((needC[l - 1].a >> 1) - 1, 0), hi = std::min(((needC[l - 1].b - 1) >> 1) + 2, h1);
        needC[l].merge(lo, std::max(hi, std::min(2, h1))); // row 1 is the reflection of row -1
        needG[l].merge(needC[l].a, needC[l].b);
    }
    for (int l = n - 1; l >= 1; --l) {
        const int hs = dim[l - 1];
        int lo = std::max(2 * needG[l].a - 2, 0), hi = std::min(2 * needG[l].b + 1, hs);
        if (needG[l].a <= 0) hi = std::max(hi, std::min(3, hs));             // reflections of rows -1, -2
        if (2 * needG[l].b + 1 > hs) lo = std::min(lo, std::max(hs - 3, 0)); // reflections past the end
        needG[l - 1].merge(lo, hi);
    }
}

// rows [row0, rows) x columns [0, cols) of an input tile that canvas rows [row_begin, row_end) can depend on
void srf_fuse_input_extent(const srf_tile &t, int levels, int out_h, int out_w, int row_begin, int row_end, int *rows, int *cols, int *row0)
{
    int hs[SRF_MAX_LEVELS], ws[SRF_MAX_LEVELS], n = 1;
    hs[0] = t.h;
    ws[0] = t.w;
    while (n < levels && hs[n - 1] >= 2 && ws[n - 1] >= 2) {
        hs[n] = (hs[n - 1] + 1) / 2;
        ws[n] = (ws[n - 1] + 1) / 2;
        ++n;
    }
    RowRange nc[SRF_MAX_LEVELS], ng[SRF_MAX_LEVELS], cc[SRF_MAX_LEVELS], cg[SRF_MAX_LEVELS];
    const int xa0 = std::max(0, -t.x), xb0 = std::min(out_w, t.x + t.w) - t.x;
    const int a0 = std::max(row_begin, t.y) - t.y, b0 = std::min(std::min(row_end, out_h), t.y + t.h) - t.y;
    const bool live = (xa0 < xb0) && (a0 < b0);
    srf_propagate_need(live ? a0 : 0, live ? b0 : 0, hs, n, nc, ng);
    srf_propagate_need(live ? xa0 : 0, live ? xb0 : 0, ws, n, cc, cg);
    *rows = ng[0].b;
    *cols = cg[0].b;
    if (row0) *row0 = ng[0].a;
}

extern "C" int srf_fuse_input_rows(int tile_h, int tile_w, int tile_y, int tile_x, int levels, int out_h, int out_w, int row_begin, int row_end,
                                   int *row0, int *row1, int *cols)
{
    if (tile_h < 1 || tile_w < 1 || levels < 1 || levels > SRF_MAX_LEVELS || !row0 || !row1 || !cols) return SRF_ERR_ARG;
    srf_tile t;
    memset(&t, 0, sizeof t);
    t.h = tile_h;
    t.w = tile_w;
    t.y = tile_y;
    t.x = tile_x;
    srf_fuse_input_extent(t, levels, out_h, out_w, row_begin, row_end, row1, cols, row0);
    return SRF_OK;
}

// ------------------------------------------------------------------------------------------------------------
// The gather's plan (host only, no device work): the band's blocks cut into cells, each with its ordered tile list and
// its class.  Returns false when some block has more than GC_MAX_T covering tiles (the general kernel takes the call).
// Shared by srf_fuse_device and by the diagnostic entry point srf_gather_plan (tests check the plan on the CPU).
// ------------------------------------------------------------------------------------------------------------
static bool srf_build_gather_cells(const srf_tile *tiles, int n_tiles, const std::vector<GatherTile> &gtiles, const srf_weight_lut *luts, int C,
                                   int out_w, int rb, int re, std::vector<GatherCell> (&cells)[GC_CLASSES], int (&class_blocks)[GC_CLASSES],
                                   double (&class_bytes)[GC_CLASSES])
{
    bool cells_ok = true;
    {
        const bool classify = getenv("SRF_GATHER_ONE_CLASS") == nullptr; // tuning switch: everything through the generic kernel
        const int nbx = srf_ceil_div(out_w, FB_W), nby = srf_ceil_div(re - rb, FB_H);
        std::vector<int> xs = {0, nbx}, ys = {0, nby};
        auto add_edge = [](std::vector<int> &v, long long e, int org, int lim, int blk, int nmax) {
            long long rel = std::min<long long>(std::max<long long>(e, org), lim) - org; // clipped to the band / canvas
            v.push_back((int)std::min<long long>(rel / blk, nmax));
            v.push_back((int)std::min<long long>((rel + blk - 1) / blk, nmax));
        };
        for (int i = 0; i < n_tiles; ++i) {
            // tile edges (the tile list changes there), the 2-pixel margins inside them (the source windows of the first /
            // last interior block must not poke out of the level-1 plane) and the feather width (inside it W_0 == 1).
            // Superfluous cuts are merged away below; a missing cut only sends a cell to the generic class.
            const long long x0 = tiles[i].x, x1 = x0 + tiles[i].w, y0 = tiles[i].y, y1 = y0 + tiles[i].h, F = luts[tiles[i].lut].len - 1;
            for (long long e : {x0, x0 + 2, x0 + F, x1 - F, x1 - 2, x1}) add_edge(xs, e, 0, out_w, FB_W, nbx);
            for (long long e : {y0, y0 + 2, y0 + F, y1 - F, y1 - 2, y1}) add_edge(ys, e, rb, re, FB_H, nby);
        }
        std::sort(xs.begin(), xs.end());
        xs.erase(std::unique(xs.begin(), xs.end()), xs.end());
        std::sort(ys.begin(), ys.end());
        ys.erase(std::unique(ys.begin(), ys.end()), ys.end());
        struct RawCell {
            GatherCell c;
            int cls;
        };
        std::vector<RawCell> raw;
        for (size_t j = 0; j + 1 < ys.size() && cells_ok; ++j)
            for (size_t i = 0; i + 1 < xs.size(); ++i) {
                const int y0 = rb + ys[j] * FB_H, y1 = std::min(rb + ys[j + 1] * FB_H, re);
                const int x0 = xs[i] * FB_W, x1 = std::min(xs[i + 1] * FB_W, out_w);
                RawCell rc;
                memset(&rc, 0, sizeof rc);
                GatherCell &cell = rc.c;
                cell.bx0 = xs[i];
                cell.by0 = ys[j];
                cell.nbx = xs[i + 1] - xs[i];
                cell.nby = ys[j + 1] - ys[j];
                int T = 0;
                for (int t = 0; t < n_tiles; ++t)
                    if (tiles[t].y < y1 && tiles[t].y + tiles[t].h > y0 && tiles[t].x < x1 && tiles[t].x + tiles[t].w > x0) {
                        if (T < GC_MAX_T) cell.tile[T] = t;
                        ++T;
                    }
                if (T > GC_MAX_T) { // more than four tiles over one block: the general kernel takes the whole call
                    cells_ok = false;
                    break;
                }
                cell.T = T;
                // interior for every tile?  The conditions are intervals per axis, so the cell's four corner blocks decide.
                bool all_full = classify && T >= 1, unit = classify && T == 1;
                for (int s = 0; s < T && all_full; ++s) {
                    const GatherTile &gt = gtiles[cell.tile[s]];
                    for (int cy = 0; cy < 2 && all_full; ++cy)
                        for (int cx = 0; cx < 2 && all_full; ++cx) {
                            const int by = cy ? cell.by0 + cell.nby - 1 : cell.by0, bx = cx ? cell.bx0 + cell.nbx - 1 : cell.bx0;
                            const int cy0 = rb + by * FB_H, cx0 = bx * FB_W;
                            const FinalGeo q = gather_geo(gt, cy0, std::min(cy0 + FB_H, re), cx0, std::min(cx0 + FB_W, out_w), C);
                            all_full = gather_slot_full(gt, q, cy0, cx0);
                        }
                    if (unit) { // W_0 == 1.0f over the whole cell: every pixel at least feather-width inside, and the profile ends at 1
                        const srf_weight_lut &L = luts[tiles[cell.tile[s]].lut];
                        const int F = L.len - 1;
                        unit = L.values[F] == 1.0f && (x0 - gt.x >= F) && (gt.x + gt.w - 1 - (x1 - 1) >= F) && (y0 - gt.y >= F) &&
                               (gt.y + gt.h - 1 - (y1 - 1) >= F);
                    }
                }
                bool edge_ok = classify && T >= 1 && getenv("SRF_GATHER_NO_EDGE") == nullptr; // block-independent conditions
                for (int s = 0; s < T && edge_ok; ++s) edge_ok = gather_slot_edge_ok(gtiles[cell.tile[s]], y0, x0);
                rc.cls = all_full ? (T == 1 ? (unit ? GC_U1 : GC_F1) : (T == 2 ? GC_F2 : GC_F4))
                                  : (edge_ok ? (T <= 2 ? GC_E2 : GC_E4) : (T <= 2 ? GC_G2 : GC_G4));
                raw.push_back(rc);
            }
        if (cells_ok) {
            auto same_list = [](const GatherCell &a, const GatherCell &b) {
                return a.T == b.T && a.tile[0] == b.tile[0] && a.tile[1] == b.tile[1] && a.tile[2] == b.tile[2] && a.tile[3] == b.tile[3];
            };
            // merge horizontally (cells arrive row by row, left to right), then vertically
            std::vector<RawCell> rows;
            for (const RawCell &rc : raw) {
                if (rc.c.nbx <= 0 || rc.c.nby <= 0) continue;
                if (!rows.empty()) {
                    RawCell &pr = rows.back();
                    if (pr.cls == rc.cls && pr.c.by0 == rc.c.by0 && pr.c.nby == rc.c.nby && pr.c.bx0 + pr.c.nbx == rc.c.bx0 && same_list(pr.c, rc.c)) {
                        pr.c.nbx += rc.c.nbx;
                        continue;
                    }
                }
                rows.push_back(rc);
            }
            std::vector<RawCell> merged;
            std::vector<char> used(rows.size(), 0);
            for (size_t i = 0; i < rows.size(); ++i) {
                if (used[i]) continue;
                RawCell cur = rows[i];
                for (size_t k = i + 1; k < rows.size(); ++k) {
                    if (used[k]) continue;
                    const RawCell &nx = rows[k];
                    if (nx.c.by0 > cur.c.by0 + cur.c.nby) break; // rows are ordered by by0: nothing further down can touch
                    if (nx.cls == cur.cls && nx.c.by0 == cur.c.by0 + cur.c.nby && nx.c.bx0 == cur.c.bx0 && nx.c.nbx == cur.c.nbx && same_list(nx.c, cur.c)) {
                        cur.c.nby += nx.c.nby;
                        used[k] = 1;
                    }
                }
                merged.push_back(cur);
            }
            for (RawCell &rc : merged) {
                const int cls = rc.cls;
                const int y0 = rb + rc.c.by0 * FB_H, y1 = std::min(rb + (rc.c.by0 + rc.c.nby) * FB_H, re);
                const int x0 = rc.c.bx0 * FB_W, x1 = std::min((rc.c.bx0 + rc.c.nbx) * FB_W, out_w);
                rc.c.first = class_blocks[cls];
                class_blocks[cls] += rc.c.nbx * rc.c.nby;
                class_bytes[cls] += (double)(y1 - y0) * (x1 - x0) * C * (3.0 * rc.c.T + 1.0);
                cells[cls].push_back(rc.c);
            }
        }
    }
    return cells_ok;
}

extern "C" int srf_gather_plan(const srf_tile *tiles, int n_tiles, int channels, int levels, const srf_weight_lut *luts, int n_luts, int out_h,
                               int out_w, int row_begin, int row_end, int *cells_out, int cap, int *n_cells)
{
    if (!tiles || n_tiles < 1 || !luts || n_luts < 1 || !n_cells || (channels != 1 && channels != 3) || levels < 1 || out_h < 1 || out_w < 1 ||
        cap < 0 || (cap > 0 && !cells_out))
        return SRF_ERR_ARG;
    *n_cells = 0;
    const int rb = std::max(row_begin, 0), re = std::min(row_end, out_h);
    if (rb >= re || n_tiles > FINAL_TMA_TILES) return SRF_ERR_UNSUPPORTED;
    std::vector<GatherTile> gtiles(n_tiles);
    for (int i = 0; i < n_tiles; ++i) {
        if (tiles[i].lut < 0 || tiles[i].lut >= n_luts || tiles[i].h < 1 || tiles[i].w < 1 || !luts[tiles[i].lut].values || luts[tiles[i].lut].len < 1)
            return SRF_ERR_ARG;
        int h = tiles[i].h, w = tiles[i].w, n = 1, h1 = 0, w1 = 0;
        while (n < levels && h >= 2 && w >= 2) { // the pyramid of srf_fuse_device: stops when a side reaches 1
            h = (h + 1) / 2;
            w = (w + 1) / 2;
            if (n == 1) {
                h1 = h;
                w1 = w;
            }
            ++n;
        }
        if (n < 3) return SRF_ERR_UNSUPPORTED; // 1- / 2-level tiles take the general kernel
        GatherTile &gt = gtiles[i];
        gt.y = tiles[i].y;
        gt.x = tiles[i].x;
        gt.h = tiles[i].h;
        gt.w = tiles[i].w;
        gt.h1 = h1;
        gt.w1 = w1;
        gt.lut_len = luts[tiles[i].lut].len;
        gt.pad = 0;
        gt.lut = nullptr;
    }
    std::vector<GatherCell> cells[GC_CLASSES];
    int class_blocks[GC_CLASSES] = {};
    double class_bytes[GC_CLASSES] = {};
    if (!srf_build_gather_cells(tiles, n_tiles, gtiles, luts, channels, out_w, rb, re, cells, class_blocks, class_bytes)) return SRF_ERR_UNSUPPORTED;
    int n = 0;
    for (int cls = 0; cls < GC_CLASSES; ++cls)
        for (const GatherCell &c : cells[cls]) {
            if (n < cap) {
                int *o = cells_out + (size_t)n * 12;
                o[0] = cls; o[1] = c.bx0; o[2] = c.by0; o[3] = c.nbx; o[4] = c.nby; o[5] = c.T;
                o[6] = c.tile[0]; o[7] = c.tile[1]; o[8] = c.tile[2]; o[9] = c.tile[3];
                o[10] = c.first; o[11] = 0;
            }
            ++n;
        }
    *n_cells = n;
    return SRF_OK;
}

static uint64_t fuse_signature(const srf_tile *tiles, int n_tiles, int C, int levels, const srf_weight_lut *luts, int n_luts, int out_h,
                               int out_w, const void *canvas, size_t canvas_pitch, int row_begin, int row_end, const void *stream)
{
    uint64_t hsh = 1469598103934665603ull;
    auto mix = [&](const void *ptr, size_t nb) {
        const uint8_t *b = (const uint8_t *)ptr;
        for (size_t i = 0; i < nb; ++i) hsh = (hsh ^ b[i]) * 1099511628211ull;
    };
    for (int i = 0; i < n_tiles; ++i) { // field by field: struct padding is not defined
        mix(&tiles[i].data, sizeof(void *));
        mix(&tiles[i].pitch, sizeof(size_t));
        const int v[5] = {tiles[i].h, tiles[i].w, tiles[i].y, tiles[i].x, tiles[i].lut};
        mix(v, sizeof v);
    }
    for (int k = 0; k < n_luts; ++k) mix(luts[k].values, sizeof(float) * luts[k].len);
    // the tuning switches change which kernels a call plans: part of the signature (a cached plan must not outlive a flip)
    const int switches = (getenv("SRF_GATHER_ONE_CLASS") ? 1 : 0) | (getenv("SRF_GATHER_NO_EDGE") ? 2 : 0) | (getenv("SRF_LEGACY_GATHER") ? 4 : 0) |
                         (getenv("SRF_GATHER_SERIAL") ? 8 : 0);
    const int geo[8] = {n_tiles, C, levels, out_h, out_w, row_begin, row_end, switches};
    mix(geo, sizeof geo);
    mix(&canvas, sizeof canvas);
    mix(&canvas_pitch, sizeof canvas_pitch);
    mix(&stream, sizeof stream);
    return hsh ? hsh : 1;
}

int srf_fuse_device(srf_ctx *ctx, const srf_tile *tiles, int n_tiles, int C, int levels, const srf_weight_lut *luts, int n_luts,
                       int out_h, int out_w, uint8_t *d_canvas, size_t canvas_pitch, int row_begin, int row_end, cudaStream_t st)
{
    // ---- group tiles by (h, w, lut) ----
    std::vector<HostGroup> groups;
    std::vector<int> tile_group(n_tiles);
    for (int i = 0; i < n_tiles; ++i) {
        int gi = -1;
        for (size_t k = 0; k < groups.size(); ++k)
            if (groups[k].h == tiles[i].h && groups[k].w == tiles[i].w && groups[k].lut == tiles[i].lut) gi = (int)k;
        if (gi < 0) {
            HostGroup hg;
            hg.h = tiles[i].h;
            hg.w = tiles[i].w;
            hg.lut = tiles[i].lut;
            memset(&hg.dev, 0, sizeof hg.dev);
            groups.push_back(hg);
            gi = (int)groups.size() - 1;
        }
        groups[gi].tiles.push_back(i);
        tile_group[i] = gi;
    }

    // ---- geometry + workspace plan ----
    size_t lut_floats = 0;
    for (int k = 0; k < n_luts; ++k) lut_floats += srf_align_up((size_t)luts[k].len, 4);
    std::vector<size_t> lut_off(n_luts);
    {
        size_t o = 0;
        for (int k = 0; k < n_luts; ++k) {
            lut_off[k] = o;
            o += srf_align_up((size_t)luts[k].len, 4);
        }
    }
    size_t pyr_floats = 0;
    auto take = [&](size_t n) {
        size_t o = pyr_floats;
        pyr_floats += srf_align_up(n, 32);
        return o;
    };
    std::vector<std::vector<size_t>> w_off(groups.size());
    for (size_t k = 0; k < groups.size(); ++k) {
        GroupDev &g = groups[k].dev;
        g.cn = C;
        g.lut_len = luts[groups[k].lut].len;
        int h = groups[k].h, w = groups[k].w, n = 1;
        g.lv[0] = LevelGeom{h, w, 0, 0, 0, 0};
        while (n < levels && h >= 2 && w >= 2) {
            const int oh = (h + 1) / 2, ow = (w + 1) / 2;
            LevelGeom L;
            L.h = oh;
            L.w = ow;
            L.pitch = (int)srf_align_up((size_t)ow, 4);
            L.plane = srf_align_up((size_t)L.pitch * oh, 32);
            L.vec_end = 0;
            L.vtail = 0;
            g.lv[n] = L;
            h = oh;
            w = ow;
            ++n;
        }
        g.n_levels = n;
        g.g1_u16 = 0;
        g.g1_pitch = n >= 2 ? (int)srf_align_up((size_t)g.lv[1].w, 8) : 0; // uint16 elements: 16-byte rows
        g.g1_plane = n >= 2 ? srf_align_up((size_t)g.g1_pitch * g.lv[1].h, 64) : 0;
    }
    std::vector<TileDev> tdev(n_tiles);
    int rc;
    // ---- cells of the row-outer gather (k_gather / k_gather_full): rectangles of blocks that share one ordered tile list
    // and one class (what can be proven about every block of the cell) ----
    const int rb = std::max(row_begin, 0), re = std::min(row_end, out_h);
    std::vector<GatherCell> cells[GC_CLASSES];
    int class_blocks[GC_CLASSES] = {};
    double class_bytes[GC_CLASSES] = {};
    std::vector<GatherTile> gtiles(n_tiles);
    for (int i = 0; i < n_tiles; ++i) {
        const GroupDev &g = groups[tile_group[i]].dev;
        GatherTile &gt = gtiles[i];
        gt.y = tiles[i].y;
        gt.x = tiles[i].x;
        gt.h = tiles[i].h;
        gt.w = tiles[i].w;
        gt.h1 = g.n_levels >= 2 ? g.lv[1].h : 0;
        gt.w1 = g.n_levels >= 2 ? g.lv[1].w : 0;
        gt.lut_len = g.lut_len;
        gt.pad = 0;
        gt.lut = nullptr; // device pointer: filled in at launch
    }
    bool cells_ok = (rb < re) && n_tiles <= FINAL_TMA_TILES && getenv("SRF_LEGACY_GATHER") == nullptr;
    for (size_t k = 0; k < groups.size() && cells_ok; ++k) cells_ok = groups[k].dev.n_levels >= 3;
    if (cells_ok) cells_ok = srf_build_gather_cells(tiles, n_tiles, gtiles, luts, C, out_w, rb, re, cells, class_blocks, class_bytes);
    // the row-outer gather also needs a TMA descriptor per tile: 16-byte aligned rows, and the driver's encoder
    for (int i = 0; i < n_tiles && cells_ok; ++i) cells_ok = (tiles[i].pitch % 16 == 0) && (((uintptr_t)tiles[i].data & 15) == 0);
    if (cells_ok && srf_tensor_map_encoder() == nullptr) cells_ok = false;
    const bool g1_u16 = cells_ok; // the row-outer gather kernels read G_1 as uint16 (x 256); k_final reads float32
    for (size_t k = 0; k < groups.size(); ++k) groups[k].dev.g1_u16 = g1_u16 ? 1 : 0;
    // ---- workspace plan: weight pyramids first (their offsets key the cross-call cache), then per tile G_l and C_l ----
    for (size_t k = 0; k < groups.size(); ++k) {
        const GroupDev &g = groups[k].dev;
        w_off[k].assign(SRF_MAX_LEVELS, 0);
        for (int l = 1; l < g.n_levels; ++l) w_off[k][l] = take(g.lv[l].plane);
    }
    std::vector<std::vector<size_t>> g_off(n_tiles, std::vector<size_t>(SRF_MAX_LEVELS, 0)), c_off = g_off;
    for (int i = 0; i < n_tiles; ++i) {
        const GroupDev &g = groups[tile_group[i]].dev;
        for (int l = 1; l < g.n_levels; ++l)
            g_off[i][l] = take((l == 1 && g1_u16) ? (g.g1_plane * C + 1) / 2 : g.lv[l].plane * C); // uint16 plane: half the floats
        for (int l = 1; l <= g.n_levels - 2; ++l) c_off[i][l] = take(g.lv[l].plane * C);
    }
    if ((rc = srf_reserve(ctx, ctx->ws_pyr, pyr_floats * sizeof(float) + 256))) return rc;
    size_t n_cells_all = 0, cell_off[GC_CLASSES];
    for (int k = 0; k < GC_CLASSES; ++k) {
        cell_off[k] = n_cells_all;
        if (cells_ok) n_cells_all += cells[k].size();
    }
    // every descriptor and the weight profiles travel as ONE blob: [tiles | groups | tile_group, tile_ids | rects | luts | cells]
    const size_t o_groups = srf_align_up(sizeof(TileDev) * n_tiles, 256);
    const size_t o_ints = o_groups + srf_align_up(sizeof(GroupDev) * groups.size(), 256);
    const size_t o_rects = o_ints + srf_align_up(sizeof(int) * n_tiles * 2, 256);
    const size_t o_luts = o_rects + srf_align_up(sizeof(int4) * n_tiles, 256);
    const size_t o_cells = o_luts + srf_align_up(lut_floats * sizeof(float), 256);
    const size_t desc_bytes = o_cells + srf_align_up(sizeof(GatherCell) * n_cells_all, 256);
    if ((rc = srf_reserve(ctx, ctx->ws_desc, desc_bytes))) return rc;
    float *pyr = (float *)ctx->ws_pyr.ptr;
    float *d_lut = (float *)((uint8_t *)ctx->ws_desc.ptr + o_luts);
    std::vector<uint8_t> blob(desc_bytes, 0);
    for (int k = 0; k < n_luts; ++k) memcpy(blob.data() + o_luts + lut_off[k] * sizeof(float), luts[k].values, sizeof(float) * luts[k].len);

    // ---- row ranges per tile and level (band mode computes only what canvas rows [row_begin,row_end) need) ----
    for (int i = 0; i < n_tiles; ++i) {
        const GroupDev &g = groups[tile_group[i]].dev;
        TileDev &t = tdev[i];
        memset(&t, 0, sizeof t);
        t.src = tiles[i].data;
        t.pitch = tiles[i].pitch;
        t.h = tiles[i].h;
        t.w = tiles[i].w;
        t.y = tiles[i].y;
        t.x = tiles[i].x;
        const int n = g.n_levels;
        for (int l = 1; l < n; ++l) t.G[l] = pyr + g_off[i][l];
        for (int l = 1; l <= n - 2; ++l) t.C[l] = pyr + c_off[i][l];
        RowRange needC[SRF_MAX_LEVELS], needG[SRF_MAX_LEVELS], colC[SRF_MAX_LEVELS], colG[SRF_MAX_LEVELS];
        int hs[SRF_MAX_LEVELS], ws[SRF_MAX_LEVELS];
        for (int l = 0; l < n; ++l) {
            hs[l] = g.lv[l].h;
            ws[l] = g.lv[l].w;
        }
        const int xa0 = std::max(0, -t.x), xb0 = std::min(out_w, t.x + t.w) - t.x;
        const int a0 = std::max(row_begin, t.y) - t.y, b0 = std::min(std::min(row_end, out_h), t.y + t.h) - t.y;
        const bool live = (xa0 < xb0) && (a0 < b0);
        srf_propagate_need(live ? a0 : 0, live ? b0 : 0, hs, n, needC, needG);
        srf_propagate_need(live ? xa0 : 0, live ? xb0 : 0, ws, n, colC, colG);
        for (int l = 1; l < n; ++l) {
            t.g0[l] = needG[l].a;
            t.g1[l] = needG[l].b;
            t.c0[l] = needC[l].a;
            t.c1[l] = needC[l].b;
            t.gx1[l] = colG[l].b;
            t.cx1[l] = colC[l].b;
        }
    }
    for (size_t k = 0; k < groups.size(); ++k) {
        GroupDev &g = groups[k].dev;
        g.lut = d_lut + lut_off[groups[k].lut];
        for (int l = 1; l < g.n_levels; ++l) {
            g.W[l] = pyr + w_off[k][l];
            g.lv[l].vec_end = 0; // set per launch below (depends on the reference's channel count)
        }
    }

    // ---- upload descriptors ----
    uint8_t *dbase = (uint8_t *)ctx->ws_desc.ptr;
    TileDev *d_tiles = (TileDev *)dbase;
    GroupDev *d_groups = (GroupDev *)(dbase + o_groups);
    int *d_tile_group = (int *)(dbase + o_ints);
    int *d_tile_ids = d_tile_group + n_tiles;
    int4 *d_rects = (int4 *)(dbase + o_rects);
    std::vector<int4> rects(n_tiles);
    for (int i = 0; i < n_tiles; ++i) rects[i] = make_int4(tiles[i].y, tiles[i].x, tiles[i].h, tiles[i].w);
    std::vector<int> ids;
    std::vector<int> ids_off(groups.size());
    for (size_t k = 0; k < groups.size(); ++k) {
        ids_off[k] = (int)ids.size();
        ids.insert(ids.end(), groups[k].tiles.begin(), groups[k].tiles.end());
    }
    std::vector<GroupDev> gdev_img(groups.size());
    for (size_t k = 0; k < groups.size(); ++k) {
        gdev_img[k] = groups[k].dev;
        for (int l = 1; l < gdev_img[k].n_levels; ++l) {
            gdev_img[k].lv[l].vec_end = srf_down_vec_end(gdev_img[k].lv[l - 1].w, gdev_img[k].lv[l].w, C);
            gdev_img[k].lv[l].vtail = 4 * ((gdev_img[k].lv[l].w * C) / 4);
        }
    }
    memcpy(blob.data(), tdev.data(), sizeof(TileDev) * n_tiles);
    memcpy(blob.data() + o_groups, gdev_img.data(), sizeof(GroupDev) * groups.size());
    memcpy(blob.data() + o_ints, tile_group.data(), sizeof(int) * n_tiles);
    memcpy(blob.data() + o_ints + sizeof(int) * n_tiles, ids.data(), sizeof(int) * n_tiles);
    memcpy(blob.data() + o_rects, rects.data(), sizeof(int4) * n_tiles);
    if (cells_ok)
        for (int k = 0; k < GC_CLASSES; ++k)
            if (!cells[k].empty()) memcpy(blob.data() + o_cells + sizeof(GatherCell) * cell_off[k], cells[k].data(), sizeof(GatherCell) * cells[k].size());
    {
        const uint64_t sig = fuse_signature(tiles, n_tiles, C, levels, luts, n_luts, out_h, out_w, d_canvas, canvas_pitch, row_begin, row_end, (const void *)st);
        if (ctx->desc_sig != sig) { // identical call as last time: the descriptors are already in ws_desc
            SRF_CHECK_CUDA(ctx, cudaMemcpyAsync(dbase, blob.data(), desc_bytes, cudaMemcpyHostToDevice, st));
            ctx->desc_sig = sig;
        }
    }

    // ---- weight pyramids depend only on (shape, levels, profile): keep them across calls while ws_pyr is unchanged ----
    std::vector<uint64_t> wkeys(groups.size());
    for (size_t k = 0; k < groups.size(); ++k) {
        uint64_t hsh = 1469598103934665603ull;
        auto mix = [&](const void *ptr, size_t nb) {
            const uint8_t *b = (const uint8_t *)ptr;
            for (size_t i = 0; i < nb; ++i) hsh = (hsh ^ b[i]) * 1099511628211ull;
        };
        const int geo[4] = {groups[k].h, groups[k].w, groups[k].dev.n_levels, C};
        mix(geo, sizeof geo);
        mix(luts[groups[k].lut].values, sizeof(float) * luts[groups[k].lut].len);
        mix(w_off[k].data(), sizeof(size_t) * w_off[k].size());
        wkeys[k] = hsh;
    }
    const bool wcache_hit_base = (ctx->wcache_base == ctx->ws_pyr.ptr) && (ctx->wcache_epoch == ctx->pyr_epoch);
    std::vector<bool> w_valid(groups.size(), false);
    for (size_t k = 0; k < groups.size(); ++k)
        w_valid[k] = wcache_hit_base && k < ctx->wcache_keys.size() && ctx->wcache_keys[k] == wkeys[k];
    ctx->wcache_keys = wkeys;
    ctx->wcache_base = ctx->ws_pyr.ptr;
    ctx->wcache_epoch = ctx->pyr_epoch;

    // ---- launches ----
    for (size_t k = 0; k < groups.size(); ++k) {
        const GroupDev &gi = gdev_img[k];
        GroupDev gw = gi; // weight pyramid: the reference's array is single-channel
        for (int l = 1; l < gw.n_levels; ++l) {
            gw.lv[l].vec_end = srf_down_vec_end(gw.lv[l - 1].w, gw.lv[l].w, 1);
            gw.lv[l].vtail = 4 * (gw.lv[l].w / 4);
        }
        const int n = gi.n_levels;
        const int nt = (int)groups[k].tiles.size();
        const int *ids_k = d_tile_ids + ids_off[k];
        // rows to launch = max over the group's tiles
        auto max_rows = [&](int l, bool cc) {
            int m = 0;
            for (int ti : groups[k].tiles) m = std::max(m, cc ? tdev[ti].c1[l] - tdev[ti].c0[l] : tdev[ti].g1[l] - tdev[ti].g0[l]);
            return m;
        };
        // weight pyramid (full planes; tiny next to the image work, and shared by the whole group)
        for (int l = 0; l + 1 < n && !w_valid[k]; ++l) {
            if (l >= 1) { // W_l -> W_{l+1}: the streaming kernel; W_0 -> W_1 reads the profile, not a plane
                const int rpc = df_rows_per_cta(ctx, srf_ceil_div(gw.lv[l + 1].w, DF_BW), gw.lv[l + 1].h);
                dim3 sgrid(srf_ceil_div(gw.lv[l + 1].w, DF_BW), srf_ceil_div(gw.lv[l + 1].h, rpc), 1);
                srf_launch_scope scope(ctx, "down_weight", 4.0 * ((double)gw.lv[l].h * gw.lv[l].w + (double)gw.lv[l + 1].h * gw.lv[l + 1].w), st);
                srf_launch_pdl(k_down_f32s<true>, sgrid, dim3(DF_BW), 0, st, d_tiles, ids_k, gw, l, rpc);
                continue;
            }
            dim3 grid(srf_ceil_div(gw.lv[l + 1].w, DN_BW), srf_ceil_div(gw.lv[l + 1].h, DN_BH), 1);
            srf_launch_scope scope(ctx, "down_weight", 4.0 * ((l ? (double)gw.lv[l].h * gw.lv[l].w : 0.0) + (double)gw.lv[l + 1].h * gw.lv[l + 1].w), st);
            srf_launch_pdl(k_down_weight0<true>, grid, dim3(kThreads), 0, st, d_tiles, ids_k, gw, l);
        }
        // image pyramids, down
        if (n >= 2) {
            const int rows = max_rows(1, false);
            if (rows > 0) {
                dim3 grid(srf_ceil_div(gi.lv[1].w, D8_BW), srf_ceil_div(rows, D8_ROWS), nt);
                double px1 = 0; // level-1 pixels produced
                for (int ti : groups[k].tiles) px1 += (double)(tdev[ti].g1[1] - tdev[ti].g0[1]) * std::min(gi.lv[1].w, tdev[ti].gx1[1]);
                srf_launch_scope scope(ctx, "down_u8", px1 * C * (4.0 /*u8 in*/ + (g1_u16 ? 2.0 : 4.0) /*G_1 out*/), st);
                if (C == 3) {
                    if (g1_u16)
                        srf_launch_pdl(k_down_u8<3, true>, grid, dim3(128 * 3), 0, st, d_tiles, ids_k, gi);
                    else
                        srf_launch_pdl(k_down_u8<3, false>, grid, dim3(128 * 3), 0, st, d_tiles, ids_k, gi);
                } else {
                    if (g1_u16)
                        srf_launch_pdl(k_down_u8<1, true>, grid, dim3(128), 0, st, d_tiles, ids_k, gi);
                    else
                        srf_launch_pdl(k_down_u8<1, false>, grid, dim3(128), 0, st, d_tiles, ids_k, gi);
                }
            }
        }
        for (int l = 1; l + 1 < n; ++l) {
            const int rows = max_rows(l + 1, false);
            if (rows <= 0) continue;
            const int rpc = df_rows_per_cta(ctx, srf_ceil_div(gi.lv[l + 1].w, DF_BW) * nt * C, rows);
            dim3 grid(srf_ceil_div(gi.lv[l + 1].w, DF_BW), srf_ceil_div(rows, rpc), nt * C);
            double pxo = 0;
            for (int ti : groups[k].tiles) pxo += (double)(tdev[ti].g1[l + 1] - tdev[ti].g0[l + 1]) * std::min(gi.lv[l + 1].w, tdev[ti].gx1[l + 1]);
            if (l == 1 && g1_u16) { // the exact integer level: uint16 in (4 x 2 B per output), float32 out
                srf_launch_scope scope(ctx, "down_f32", pxo * C * (8.0 + 4.0), st);
                srf_launch_pdl(k_down_u16, grid, dim3(DF_BW), 0, st, d_tiles, ids_k, gi, rpc);
                continue;
            }
            srf_launch_scope scope(ctx, "down_f32", pxo * C * (16.0 + 4.0), st);
            srf_launch_pdl(k_down_f32s<false>, grid, dim3(DF_BW), 0, st, d_tiles, ids_k, gi, l, rpc);
        }
        // up: l = n-2 .. 1
        for (int l = n - 2; l >= 1; --l) {
            const int rows = max_rows(l, true);
            if (rows <= 0) continue;
            dim3 grid(srf_ceil_div(gi.lv[l].w, UP_BW), srf_ceil_div(rows, UP_BH), nt * C);
            double pxo = 0;
            for (int ti : groups[k].tiles) pxo += (double)(tdev[ti].c1[l] - tdev[ti].c0[l]) * std::min(gi.lv[l].w, tdev[ti].cx1[l]);
            // per output element: G_l + W_l/C read, C_l written, C_{l+1} (or W_{l+1}/C) + G_{l+1} at quarter density
            srf_launch_scope scope(ctx, "up_level", pxo * C * ((l == 1 && g1_u16 ? 2.0 : 4.0) + 4.0 / C + 4.0 + 2.0), st);
            if (l == n - 2) {
                srf_launch_pdl(k_up_level_top<true>, grid, dim3(kThreads), 0, st, d_tiles, ids_k, gi, l);
            } else {
                const int gx = srf_ceil_div(gi.lv[l].w, UL_W);
                const bool g16 = (l == 1) && g1_u16;
                if ((long long)gx * srf_ceil_div(rows, UL_H) * nt * C >= 4LL * ctx->sm_count) {
                    dim3 sgrid(gx, srf_ceil_div(rows, UL_H), nt * C);
                    if (g16)
                        srf_launch_pdl(k_up_levels<UL_H, true>, sgrid, dim3(UL_PAIRS), 0, st, d_tiles, ids_k, gi, l);
                    else
                        srf_launch_pdl(k_up_levels<UL_H, false>, sgrid, dim3(UL_PAIRS), 0, st, d_tiles, ids_k, gi, l);
                } else { // small level / thin band: short CTAs
                    dim3 sgrid(gx, srf_ceil_div(rows, 8), nt * C);
                    if (g16)
                        srf_launch_pdl(k_up_levels<8, true>, sgrid, dim3(UL_PAIRS), 0, st, d_tiles, ids_k, gi, l);
                    else
                        srf_launch_pdl(k_up_levels<8, false>, sgrid, dim3(UL_PAIRS), 0, st, d_tiles, ids_k, gi, l);
                }
            }
        }
    }
    SRF_CHECK_CUDA(ctx, cudaGetLastError());

    // ---- canvas gather ----
    if (rb < re) {
        FinalParams P;
        P.tiles = d_tiles;
        P.groups = d_groups;
        P.tile_group = d_tile_group;
        P.rects = d_rects;
        P.n_tiles = n_tiles;
        P.out_h = out_h;
        P.out_w = out_w;
        P.canvas = d_canvas;
        P.canvas_pitch = canvas_pitch;
        P.row_begin = rb;
        P.row_end = re;
        double cov = 0; // tile pixels inside the band
        for (int i = 0; i < n_tiles; ++i) {
            const int a = std::max(rb, tiles[i].y), b = std::min(re, tiles[i].y + tiles[i].h);
            const int c0 = std::max(0, tiles[i].x), c1 = std::min(out_w, tiles[i].x + tiles[i].w);
            if (a < b && c0 < c1) cov += (double)(b - a) * (c1 - c0);
        }
        // TMA descriptors of the level-1 planes and the uint8 source of every tile with >= 3 levels (kernel parameters)
        static thread_local FinalTmaps TM; // 18 KB: kept off the stack; filled below before every launch
        static thread_local GatherTiles GT;
        P.tma_mask = 0;
        if (PFN_encodeTiled enc = srf_tensor_map_encoder()) {
            for (int i = 0; i < n_tiles && i < FINAL_TMA_TILES; ++i) {
                const GroupDev &g = gdev_img[tile_group[i]];
                const TileDev &t = tdev[i];
                if (g.n_levels < 3 || t.pitch % 16 != 0 || ((uintptr_t)t.src & 15)) continue;
                const LevelGeom &L1 = g.lv[1];
                const cuuint64_t dims3[3] = {(cuuint64_t)L1.pitch, (cuuint64_t)L1.h, (cuuint64_t)C};
                const cuuint64_t str3[2] = {(cuuint64_t)L1.pitch * sizeof(float), (cuuint64_t)L1.plane * sizeof(float)};
                const cuuint32_t box3[3] = {FB_SRC_COLS, FB_SRC_ROWS, (cuuint32_t)C}, one3[3] = {1, 1, 1};
                const cuuint64_t dims16[3] = {(cuuint64_t)g.g1_pitch, (cuuint64_t)L1.h, (cuuint64_t)C}; // the uint16 G_1 planes
                const cuuint32_t box16[3] = {FB_G_COLS, FB_SRC_ROWS, (cuuint32_t)C};
                const cuuint64_t str16[2] = {(cuuint64_t)g.g1_pitch * sizeof(unsigned short), (cuuint64_t)g.g1_plane * sizeof(unsigned short)};
                const cuuint64_t dims2[2] = {(cuuint64_t)(t.pitch / 4), (cuuint64_t)t.h};
                const cuuint64_t str2[1] = {(cuuint64_t)t.pitch};
                const cuuint32_t box2[2] = {FB_U8_STRIDE / 4, FB_H}, one2[2] = {1, 1};
                const bool ok =
                    enc(&TM.m[i][0], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, t.C[1], dims3, str3, box3, one3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS &&
                    (g1_u16 ? enc(&TM.m[i][1], CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, t.G[1], dims16, str16, box16, one3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
                            : enc(&TM.m[i][1], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, t.G[1], dims3, str3, box3, one3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) == CUDA_SUCCESS &&
                    enc(&TM.m[i][2], CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint8_t *>(t.src), dims2, str2, box2, one2,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
                if (ok) P.tma_mask |= 1ull << i;
            }
        }
        const unsigned long long all_mask = n_tiles >= 64 ? ~0ull : ((1ull << n_tiles) - 1ull);
        if (cells_ok && n_tiles <= FINAL_TMA_TILES && (P.tma_mask & all_mask) == all_mask) {
            // ---- row-outer gather: one launch per class of cells ----
            for (int i = 0; i < n_tiles; ++i) {
                GT.t[i] = gtiles[i];
                GT.t[i].lut = groups[tile_group[i]].dev.lut;
            }
            const GatherCell *d_cells = (const GatherCell *)(dbase + o_cells);
            // The classes write disjoint blocks of the canvas, so their launches are independent: the biggest stays on the
            // caller's stream, the others fork onto side streams and join back -- short launches then fill each other's
            // ramps and tails instead of queueing (a band of an 8-rank job is six launches of a few hundred CTAs each).
            // Under stream capture the fork / join events become graph edges.  Serial while per-kernel timing is on.
            static_assert(GC_CLASSES - 1 <= srf_ctx::kForkStreams, "one side stream per class but the biggest");
            int main_cls = 0;
            for (int cls = 1; cls < GC_CLASSES; ++cls)
                if (class_blocks[cls] > class_blocks[main_cls]) main_cls = cls;
            const bool fork = !ctx->profile && getenv("SRF_GATHER_SERIAL") == nullptr;
            const cudaStream_t st_main = st;
            int n_forked = 0;
            if (fork) SRF_CHECK_CUDA(ctx, cudaEventRecord(ctx->ev_fork, st_main));
            for (int cls = 0; cls < GC_CLASSES; ++cls) {
                if (class_blocks[cls] == 0) continue;
                cudaStream_t st = st_main; // shadows the caller's stream for this launch
                if (fork && cls != main_cls) {
                    st = ctx->fork_stream[n_forked];
                    SRF_CHECK_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_fork, 0));
                }
                GatherParams GP;
                GP.cells = d_cells + cell_off[cls];
                GP.n_cells = (int)cells[cls].size();
                GP.out_w = out_w;
                GP.canvas = d_canvas;
                GP.canvas_pitch = canvas_pitch;
                GP.row_begin = rb;
                GP.row_end = re;
                srf_launch_scope scope(ctx, "final_gather", class_bytes[cls], st);
                const dim3 grid(class_blocks[cls]);
#define SRF_GATHER_LAUNCH(KERNEL, SMEM_T, CC)                                                                                              \
    do {                                                                                                                                   \
        const size_t sm = sizeof(GatherSmem<CC, SMEM_T>) + 128;                                                                            \
        SRF_CHECK_CUDA(ctx, cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));                           \
        SRF_CHECK_CUDA(ctx, cudaFuncSetAttribute(KERNEL, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)); \
        srf_launch_pdl(KERNEL, grid, dim3(FB_PAIRS * CC), sm, st, GP, GT, TM);                                                            \
    } while (0)
                if (C == 3) {
                    switch (cls) {
                    case GC_U1: SRF_GATHER_LAUNCH((k_gather_full<3, 1, true>), 1, 3); break;
                    case GC_F1: SRF_GATHER_LAUNCH((k_gather_full<3, 1, false>), 1, 3); break;
                    case GC_F2: SRF_GATHER_LAUNCH((k_gather_full<3, 2, false>), 2, 3); break;
                    case GC_F4: SRF_GATHER_LAUNCH((k_gather_full<3, 4, false>), 4, 3); break;
                    case GC_G2: SRF_GATHER_LAUNCH((k_gather<3, 2>), 2, 3); break;
                    case GC_E2: SRF_GATHER_LAUNCH((k_gather_full<3, 2, false, true>), 2, 3); break;
                    case GC_E4: SRF_GATHER_LAUNCH((k_gather_full<3, 4, false, true>), 4, 3); break;
                    default: SRF_GATHER_LAUNCH((k_gather<3, 4>), 4, 3); break;
                    }
                } else {
                    switch (cls) {
                    case GC_U1: SRF_GATHER_LAUNCH((k_gather_full<1, 1, true>), 1, 1); break;
                    case GC_F1: SRF_GATHER_LAUNCH((k_gather_full<1, 1, false>), 1, 1); break;
                    case GC_F2: SRF_GATHER_LAUNCH((k_gather_full<1, 2, false>), 2, 1); break;
                    case GC_F4: SRF_GATHER_LAUNCH((k_gather_full<1, 4, false>), 4, 1); break;
                    case GC_G2: SRF_GATHER_LAUNCH((k_gather<1, 2>), 2, 1); break;
                    case GC_E2: SRF_GATHER_LAUNCH((k_gather_full<1, 2, false, true>), 2, 1); break;
                    case GC_E4: SRF_GATHER_LAUNCH((k_gather_full<1, 4, false, true>), 4, 1); break;
                    default: SRF_GATHER_LAUNCH((k_gather<1, 4>), 4, 1); break;
                    }
                }
#undef SRF_GATHER_LAUNCH
                if (fork && cls != main_cls) {
                    SRF_CHECK_CUDA(ctx, cudaEventRecord(ctx->ev_join[n_forked], st));
                    ++n_forked;
                }
            }
            for (int i = 0; i < n_forked; ++i) SRF_CHECK_CUDA(ctx, cudaStreamWaitEvent(st_main, ctx->ev_join[i], 0));
            SRF_CHECK_CUDA(ctx, cudaGetLastError());
            return SRF_OK;
        }
        if (g1_u16) { // G_1 was written as uint16 for the row-outer gather: k_final cannot read it
            srf_set_error(ctx, "laplacian_fuse: cuTensorMapEncodeTiled failed for a tile the gather was planned for");
            return SRF_ERR_CUDA;
        }
        // ---- general gather (1- / 2-level tiles, tiles without a TMA descriptor, more than four tiles over a block) ----
        dim3 grid(srf_ceil_div(out_w, FB_W), srf_ceil_div(re - rb, FB_H), 1);
        // per covered tile pixel: u8 in + C_1, G_1 at quarter density; per canvas pixel: u8 out
        srf_launch_scope scope(ctx, "final_gather", cov * C * (1.0 + 2.0) + (double)(re - rb) * out_w * C, st);
        if (C == 3) {
            const size_t sm = sizeof(FinalSmem<3>) + 128;
            SRF_CHECK_CUDA(ctx, cudaFuncSetAttribute(k_final<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            SRF_CHECK_CUDA(ctx, cudaFuncSetAttribute(k_final<3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            srf_launch_pdl(k_final<3>, grid, dim3(FB_PAIRS * 3), sm, st, P, TM);
        } else {
            const size_t sm = sizeof(FinalSmem<1>) + 128;
            SRF_CHECK_CUDA(ctx, cudaFuncSetAttribute(k_final<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            SRF_CHECK_CUDA(ctx, cudaFuncSetAttribute(k_final<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            srf_launch_pdl(k_final<1>, grid, dim3(FB_PAIRS), sm, st, P, TM);
        }
        SRF_CHECK_CUDA(ctx, cudaGetLastError());
    }
    return SRF_OK;
}

int srf_check_fuse_args(srf_ctx *ctx, const srf_tile *tiles, int n_tiles, int C, int levels, const srf_weight_lut *luts, int n_luts,
                           int out_h, int out_w, const void *canvas, size_t canvas_pitch, int row_begin, int row_end)
{
    SRF_REQUIRE(ctx, tiles && n_tiles >= 1, "laplacian_fuse: no tiles");
    SRF_REQUIRE(ctx, C == 1 || C == 3, "laplacian_fuse: channels must be 1 or 3 (got %d)", C);
    SRF_REQUIRE(ctx, levels >= 1 && levels <= SRF_MAX_LEVELS, "laplacian_fuse: levels must be 1..%d (got %d)", SRF_MAX_LEVELS, levels);
    SRF_REQUIRE(ctx, luts && n_luts >= 1, "laplacian_fuse: no weight profiles");
    SRF_REQUIRE(ctx, out_h >= 1 && out_w >= 1 && canvas, "laplacian_fuse: empty canvas");
    SRF_REQUIRE(ctx, canvas_pitch >= (size_t)out_w * C, "laplacian_fuse: canvas pitch < row bytes");
    SRF_REQUIRE(ctx, row_begin >= 0 && row_begin <= row_end && row_end <= out_h, "laplacian_fuse: bad row band [%d,%d) of %d", row_begin,
                row_end, out_h);
    for (int k = 0; k < n_luts; ++k) SRF_REQUIRE(ctx, luts[k].values && luts[k].len >= 1, "laplacian_fuse: weight profile %d empty", k);
    for (int i = 0; i < n_tiles; ++i) {
        const srf_tile &t = tiles[i];
        SRF_REQUIRE(ctx, t.data, "laplacian_fuse: tile %d has no data", i);
        SRF_REQUIRE(ctx, t.h >= 1 && t.w >= 1, "laplacian_fuse: tile %d is empty (%dx%d)", i, t.h, t.w);
        SRF_REQUIRE(ctx, t.pitch >= (size_t)t.w * C, "laplacian_fuse: tile %d pitch < row bytes", i);
        SRF_REQUIRE(ctx, t.y >= 0 && t.x >= 0, "laplacian_fuse: tile %d has a negative position (%d,%d)", i, t.y, t.x);
        SRF_REQUIRE(ctx, t.lut >= 0 && t.lut < n_luts, "laplacian_fuse: tile %d weight profile index %d out of range", i, t.lut);
    }
    return SRF_OK;
}

extern "C" int srf_laplacian_fuse(srf_ctx *ctx, const srf_tile *tiles, int n_tiles, int channels, int levels, const srf_weight_lut *luts,
                                  int n_luts, int out_h, int out_w, uint8_t *d_canvas, size_t canvas_pitch, int row_begin, int row_end,
                                  void *stream)
{
    if (!ctx) return SRF_ERR_ARG;
#ifdef SRF_HOST_TIMING
    static double acc_t[5] = {0, 0, 0, 0, 0};
    static int acc_n = 0;
    auto now = []() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double ht0 = now();
#endif
    int rc = srf_check_fuse_args(ctx, tiles, n_tiles, channels, levels, luts, n_luts, out_h, out_w, d_canvas, canvas_pitch, row_begin, row_end);
    if (rc) return rc;
    for (int i = 0; i < n_tiles; ++i)
        SRF_REQUIRE(ctx, ((uintptr_t)tiles[i].data % 16 == 0) && (tiles[i].pitch % 16 == 0 || tiles[i].h == 1),
                    "srf_laplacian_fuse: device tile %d must be 16-byte aligned with a 16-byte-multiple pitch "
                    "(use srf_laplacian_fuse_host, or copy into an aligned buffer)", i);
    SRF_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    // the stream HANDLE is part of the signature (by value: hashing the address of a local copy made the two call sites
    // disagree, so no call ever counted as a repeat and the graph was never captured)
    const uint64_t sig = fuse_signature(tiles, n_tiles, channels, levels, luts, n_luts, out_h, out_w, d_canvas, canvas_pitch, row_begin, row_end, (const void *)st);
    const bool repeat = !ctx->profile && st != nullptr && sig == ctx->fuse_sig && ctx->desc_sig == sig;
    if (repeat && ctx->fuse_graph && ctx->graph_epoch == ctx->ws_epoch) { // third and later identical calls: replay
#ifdef SRF_HOST_TIMING
        const double ht1 = now();
#endif
        SRF_CHECK_CUDA(ctx, cudaGraphLaunch(ctx->fuse_graph, st));
#ifdef SRF_HOST_TIMING
        const double ht2 = now();
#endif
        ctx->launches += ctx->fuse_graph_launches;
        ctx->graph_replays++;
        ctx->ev_valid = false;
        rc = srf_note_device_work(ctx, st);
#ifdef SRF_HOST_TIMING
        const double ht3 = now();
        acc_t[0] += ht1 - ht0; acc_t[1] += ht2 - ht1; acc_t[2] += ht3 - ht2;
        if (++acc_n % 200 == 0) {
            fprintf(stderr, "[srf timing] replay: checks+signature %.1f us, cudaGraphLaunch %.1f us, note %.1f us (avg of 200)\n", acc_t[0] / 200, acc_t[1] / 200, acc_t[2] / 200);
            acc_t[0] = acc_t[1] = acc_t[2] = 0;
        }
#endif
        return rc;
    }
    if (ctx->fuse_graph) {
        cudaGraphExecDestroy(ctx->fuse_graph);
        ctx->fuse_graph = nullptr;
    }
    if (repeat) { // second identical call: every workspace is sized and the descriptors are resident -> capture it
        const int64_t l0 = ctx->launches;
        const uint64_t epoch = ctx->ws_epoch;
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            rc = srf_fuse_device(ctx, tiles, n_tiles, channels, levels, luts, n_luts, out_h, out_w, d_canvas, canvas_pitch, row_begin, row_end, st);
            const cudaError_t ce = cudaStreamEndCapture(st, &graph);
            if (rc == SRF_OK && ce == cudaSuccess && graph && epoch == ctx->ws_epoch &&
                cudaGraphInstantiate(&ctx->fuse_graph, graph, 0) == cudaSuccess) {
                ctx->fuse_graph_launches = ctx->launches - l0;
                ctx->graph_epoch = epoch;
                cudaGraphDestroy(graph);
                SRF_CHECK_CUDA(ctx, cudaGraphLaunch(ctx->fuse_graph, st));
                ctx->graph_replays++;
                ctx->ev_valid = false;
                return srf_note_device_work(ctx, st);
            }
            if (graph) cudaGraphDestroy(graph);
            ctx->fuse_graph = nullptr;
            ctx->launches = l0;
            (void)cudaGetLastError(); // capture failed: fall through to a plain enqueue
        } else {
            (void)cudaGetLastError();
        }
    }
    ctx->fuse_sig = sig;
    SRF_CHECK_CUDA(ctx, cudaEventRecord(ctx->ev0, st));
    rc = srf_fuse_device(ctx, tiles, n_tiles, channels, levels, luts, n_luts, out_h, out_w, d_canvas, canvas_pitch, row_begin, row_end, st);
    if (rc) return rc;
    SRF_CHECK_CUDA(ctx, cudaEventRecord(ctx->ev1, st));
    ctx->ev_valid = true;
    return srf_note_device_work(ctx, st);
}

extern "C" int srf_laplacian_fuse_host(srf_ctx *ctx, const srf_tile *tiles, int n_tiles, int channels, int levels,
                                       const srf_weight_lut *luts, int n_luts, int out_h, int out_w, uint8_t *h_canvas, size_t canvas_pitch,
                                       int row_begin, int row_end)
{
    if (!ctx) return SRF_ERR_ARG;
    int rc = srf_check_fuse_args(ctx, tiles, n_tiles, channels, levels, luts, n_luts, out_h, out_w, h_canvas, canvas_pitch, row_begin, row_end);
    if (rc) return rc;
    SRF_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (int hrc = srf_host_enter(ctx)) return hrc;
    // device staging: every tile at a 256-byte aligned offset with a 16-byte-multiple pitch, then the canvas band
    std::vector<srf_tile> dt(tiles, tiles + n_tiles);
    std::vector<size_t> off(n_tiles);
    size_t total = 0;
    for (int i = 0; i < n_tiles; ++i) {
        const size_t p = srf_align_up((size_t)tiles[i].w * channels, 16);
        off[i] = total;
        dt[i].pitch = p;
        total += srf_align_up(p * tiles[i].h, 256);
    }
    const size_t cpitch = srf_align_up((size_t)out_w * channels, 256);
    const int band = row_end - row_begin;
    if ((rc = srf_reserve(ctx, ctx->ws_stage_in, total))) return rc;
    if ((rc = srf_reserve(ctx, ctx->ws_stage_out, cpitch * (size_t)std::max(band, 1)))) return rc;
    cudaStream_t st = ctx->own_stream;
    uint8_t *din = (uint8_t *)ctx->ws_stage_in.ptr;
    size_t h2d = 0;
    for (int i = 0; i < n_tiles; ++i) {
        dt[i].data = din + off[i];
        int rows = 0, cols = 0, row0 = 0; // only the part of the tile the requested canvas rows can depend on travels
        srf_fuse_input_extent(tiles[i], levels, out_h, out_w, row_begin, row_end, &rows, &cols, &row0);
        if (rows > row0 && cols > 0) {
            h2d += (size_t)cols * channels * (rows - row0);
            if ((rc = srf_copy2d_h2d(ctx, din + off[i] + (size_t)row0 * dt[i].pitch, dt[i].pitch, tiles[i].data + (size_t)row0 * tiles[i].pitch,
                                     tiles[i].pitch, (size_t)cols * channels, rows - row0, st)))
                return rc;
        }
    }
    ctx->last_h2d = h2d;
    ctx->last_d2h = (size_t)out_w * channels * (size_t)std::max(band, 0);
    SRF_CHECK_CUDA(ctx, cudaEventRecord(ctx->ev0, st));
    // the staged canvas holds only the band: shift the base so canvas row `row_begin` lands on staged row 0
    uint8_t *dcanvas = (uint8_t *)ctx->ws_stage_out.ptr - (size_t)row_begin * cpitch;
    rc = srf_fuse_device(ctx, dt.data(), n_tiles, channels, levels, luts, n_luts, out_h, out_w, dcanvas, cpitch, row_begin, row_end, st);
    if (rc) return rc;
    SRF_CHECK_CUDA(ctx, cudaEventRecord(ctx->ev1, st));
    ctx->ev_valid = true;
    if (band > 0)
        SRF_CHECK_CUDA(ctx, cudaMemcpy2DAsync(h_canvas + (size_t)row_begin * canvas_pitch, canvas_pitch, ctx->ws_stage_out.ptr, cpitch,
                                              (size_t)out_w * channels, band, cudaMemcpyDeviceToHost, st));
    SRF_CHECK_CUDA(ctx, cudaStreamSynchronize(st));
    return SRF_OK;
}
