// srf_pipeline.cu -- the reference pipeline's stage 3 + stage 4 as ONE host-facing call
// (/root/reference/main.py:347 laplacian_fusion followed by :369 evaluate_full_reference): host tiles in, host
// canvas + PSNR/SSIM sums out, with the fused canvas never leaving the device between the two stages.
//
// The call is PCIe-bound (1.8 GB up, 0.6 GB down at 200 MP against ~8 ms of kernels), so it is organised as a
// band pipeline over three streams that keeps both copy engines busy:
//   upload   : for band b, the tiles that first become necessary for it, then the comparison image's rows of band b
//   compute  : fuse band b (srf_fuse_device recomputes exactly the pyramid rows a band depends on), then the
//              PSNR/SSIM pass of band b-1 (its 5-row halo below is now fused)
//   download : canvas rows of band b, as soon as they are fused -- concurrent with the uploads of bands b+1..
// Partial sums are additive (srf_iqa_sums), one pinned slot per band.
#include "srf_common.cuh"

#include <algorithm>

namespace {
constexpr int kMaxBands = 16;

cudaEvent_t take_event(srf_ctx *ctx)
{
    cudaEvent_t e = nullptr;
    if (!ctx->ev_pool.empty()) {
        e = ctx->ev_pool.back();
        ctx->ev_pool.pop_back();
    } else {
        cudaEventCreate(&e);
    }
    return e;
}
} // namespace

extern "C" int srf_fuse_assess_host(srf_ctx *ctx, const srf_tile *tiles, int n_tiles, int channels, int levels,
                                    const srf_weight_lut *luts, int n_luts, int out_h, int out_w, uint8_t *h_canvas, size_t canvas_pitch,
                                    const uint8_t *h_reference, size_t reference_pitch, double data_range, int row_begin, int row_end,
                                    srf_iqa_sums *h_out)
{
    if (!ctx) return SRF_ERR_ARG;
    int rc = srf_check_fuse_args(ctx, tiles, n_tiles, channels, levels, luts, n_luts, out_h, out_w, h_canvas, canvas_pitch, row_begin, row_end);
    if (rc) return rc;
    // rows to FUSE: the requested band plus the SSIM halo (a rank of a band-sharded job needs its neighbours' 5 rows)
    // (the start is rounded down to an even row: tile origins are even in the pipeline, and the gather's predicate-free
    // interior path needs blocks of even parity -- an odd band start would push every block onto the generic path)
    const int F0 = std::max(row_begin - 5, 0) & ~1, F1 = (row_end > row_begin) ? std::min(row_end + 5, out_h) : F0;
    SRF_REQUIRE(ctx, h_reference && h_out, "srf_fuse_assess_host: null pointer");
    SRF_REQUIRE(ctx, reference_pitch >= (size_t)out_w * channels, "srf_fuse_assess_host: reference pitch < row bytes");
    SRF_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (int hrc = srf_host_enter(ctx)) return hrc;

    // ---- device staging layout ----
    std::vector<srf_tile> dt(tiles, tiles + n_tiles);
    std::vector<size_t> off(n_tiles);
    size_t total = 0;
    for (int i = 0; i < n_tiles; ++i) {
        const size_t p = srf_align_up((size_t)tiles[i].w * channels, 16);
        off[i] = total;
        dt[i].pitch = p;
        total += srf_align_up(p * tiles[i].h, 256);
    }
    // packed device canvas / comparison image when the host side is packed too: the band copies are then plain
    // contiguous transfers (the gather's congruent 16-byte stores and the IQA loader accept any pitch)
    const bool packed = canvas_pitch == (size_t)out_w * channels && reference_pitch == (size_t)out_w * channels;
    const size_t cpitch = packed ? (size_t)out_w * channels : srf_align_up((size_t)out_w * channels, 256);
    const size_t canvas_bytes = srf_align_up(cpitch * (size_t)out_h, 256);
    if ((rc = srf_reserve(ctx, ctx->ws_stage_in, total))) return rc;
    if ((rc = srf_reserve(ctx, ctx->ws_stage_out, 2 * canvas_bytes))) return rc; // [fused canvas | comparison image]
    if ((rc = srf_reserve_pinned(ctx, 4096))) return rc;
    cudaStream_t st = ctx->own_stream, up = ctx->copy_stream, dn = ctx->down_stream;
    uint8_t *din = (uint8_t *)ctx->ws_stage_in.ptr;
    uint8_t *dcanvas = (uint8_t *)ctx->ws_stage_out.ptr, *dref = dcanvas + canvas_bytes;
    for (int i = 0; i < n_tiles; ++i) dt[i].data = din + off[i];

    // ---- bands ----
    const int nb = std::max(1, std::min(kMaxBands, (F1 - F0) / 1024));
    const int per = (int)srf_align_up((size_t)std::max(1, (F1 - F0 + nb - 1) / nb), 16); // sub-bands start on even rows too
    auto band_lo = [&](int b) { return std::min(F0 + b * per, F1); };
    auto band_hi = [&](int b) { return std::min(F0 + (b + 1) * per, F1); };
    auto own_lo = [&](int b) { return std::max(band_lo(b), row_begin); }; // the part of a sub-band the caller asked for
    auto own_hi = [&](int b) { return std::min(band_hi(b), row_end); };
    auto touches = [&](int i, int r0, int r1) { return tiles[i].y < r1 && tiles[i].y + tiles[i].h > r0; };

    // reserve every workspace for the full tile set up front (and build the shared weight pyramids): an empty band
    rc = srf_fuse_device(ctx, dt.data(), n_tiles, channels, levels, luts, n_luts, out_h, out_w, dcanvas, cpitch, 0, 0, st);
    if (rc) return rc;

    std::vector<cudaEvent_t> ev_up(nb), ev_fused(nb), ev_ref(nb);
    for (int b = 0; b < nb; ++b) {
        ev_up[b] = take_event(ctx);
        ev_fused[b] = take_event(ctx);
        ev_ref[b] = take_event(ctx);
    }
    size_t h2d = 0;
    // everything that can fail runs inside `run`; whatever it returns, the single exit below drains the three streams
    // (queued copies reference the caller's buffers and the staging workspaces) and hands the events back to the pool
    auto run = [&]() -> int {
    // ---- upload stream: everything is queued now; the copy engine works through it in band order.  Tiles lead, the
    // comparison image's rows trail by one band: fusing (and downloading) a band only needs its tiles, the assessment
    // of band b-1 only needs comparison rows up to band b -- so the tail of the call is one small upload + one small
    // kernel, not a fuse + download.
    std::vector<char> uploaded(n_tiles, 0);
    auto upload_ref = [&](int b) -> int {
        const int r0 = band_lo(b), r1 = band_hi(b);
        if (r1 > r0)
            SRF_CHECK_CUDA(ctx, cudaMemcpy2DAsync(dref + (size_t)r0 * cpitch, cpitch, h_reference + (size_t)r0 * reference_pitch, reference_pitch,
                                                  (size_t)out_w * channels, r1 - r0, cudaMemcpyHostToDevice, up));
        SRF_CHECK_CUDA(ctx, cudaEventRecord(ev_ref[b], up));
        return SRF_OK;
    };
    for (int b = 0; b < nb; ++b) {
        const int r0 = band_lo(b), r1 = band_hi(b);
        for (int i = 0; i < n_tiles; ++i) {
            if (uploaded[i] || !touches(i, r0, r1)) continue;
            uploaded[i] = 1;
            int rows = 0, cols = 0, row0 = 0; // only the part of the tile the canvas can depend on travels (exact dependency cone)
            srf_fuse_input_extent(tiles[i], levels, out_h, out_w, F0, F1, &rows, &cols, &row0);
            if (rows > row0 && cols > 0) {
                h2d += (size_t)cols * channels * (rows - row0);
                if ((rc = srf_copy2d_h2d(ctx, din + off[i] + (size_t)row0 * dt[i].pitch, dt[i].pitch, tiles[i].data + (size_t)row0 * tiles[i].pitch,
                                         tiles[i].pitch, (size_t)cols * channels, rows - row0, up)))
                    return rc;
            }
        }
        SRF_CHECK_CUDA(ctx, cudaEventRecord(ev_up[b], up));
        if (b >= 1 && (rc = upload_ref(b - 1))) return rc;
    }
    if ((rc = upload_ref(nb - 1))) return rc;
    // ---- compute + download ----
    SRF_CHECK_CUDA(ctx, cudaEventRecord(ctx->ev0, st));
    std::vector<srf_tile> sub;
    for (int b = 0; b < nb; ++b) {
        const int r0 = band_lo(b), r1 = band_hi(b);
        SRF_CHECK_CUDA(ctx, cudaStreamWaitEvent(st, ev_up[b], 0));
        sub.clear();
        for (int i = 0; i < n_tiles; ++i)
            if (touches(i, r0, r1)) sub.push_back(dt[i]); // list order is the accumulation order
        if (r1 > r0) {
            if (!sub.empty()) {
                rc = srf_fuse_device(ctx, sub.data(), (int)sub.size(), channels, levels, luts, n_luts, out_h, out_w, dcanvas, cpitch, r0, r1, st);
                if (rc) return rc;
            } else {
                SRF_CHECK_CUDA(ctx, cudaMemset2DAsync(dcanvas + (size_t)r0 * cpitch, cpitch, 0, (size_t)out_w * channels, r1 - r0, st));
            }
        }
        SRF_CHECK_CUDA(ctx, cudaEventRecord(ev_fused[b], st));
        SRF_CHECK_CUDA(ctx, cudaStreamWaitEvent(dn, ev_fused[b], 0));
        if (own_hi(b) > own_lo(b))
            SRF_CHECK_CUDA(ctx, cudaMemcpy2DAsync(h_canvas + (size_t)own_lo(b) * canvas_pitch, canvas_pitch, dcanvas + (size_t)own_lo(b) * cpitch,
                                                  cpitch, (size_t)out_w * channels, own_hi(b) - own_lo(b), cudaMemcpyDeviceToHost, dn));
        if (b >= 1 && own_hi(b - 1) > own_lo(b - 1)) {
            SRF_CHECK_CUDA(ctx, cudaStreamWaitEvent(st, ev_ref[b], 0)); // comparison rows of bands <= b (5-row halo below b-1)
            rc = srf_iqa_enqueue(ctx, dref, cpitch, dcanvas, cpitch, out_h, out_w, channels, 0, 0, data_range, own_lo(b - 1), own_hi(b - 1), st,
                                 b - 1);
            if (rc) return rc;
        }
    }
    if (own_hi(nb - 1) > own_lo(nb - 1)) {
        SRF_CHECK_CUDA(ctx, cudaStreamWaitEvent(st, ev_ref[nb - 1], 0));
        rc = srf_iqa_enqueue(ctx, dref, cpitch, dcanvas, cpitch, out_h, out_w, channels, 0, 0, data_range, own_lo(nb - 1), own_hi(nb - 1), st,
                             nb - 1);
        if (rc) return rc;
    }
    SRF_CHECK_CUDA(ctx, cudaEventRecord(ctx->ev1, st));
    return SRF_OK;
    };
    rc = run();
    const cudaError_t e_up = cudaStreamSynchronize(up), e_st = cudaStreamSynchronize(st), e_dn = cudaStreamSynchronize(dn);
    for (int b = 0; b < nb; ++b) {
        ctx->ev_pool.push_back(ev_up[b]);
        ctx->ev_pool.push_back(ev_fused[b]);
        ctx->ev_pool.push_back(ev_ref[b]);
    }
    ctx->ev_valid = (rc == SRF_OK);
    if (rc) return rc;
    SRF_CHECK_CUDA(ctx, e_up);
    SRF_CHECK_CUDA(ctx, e_st);
    SRF_CHECK_CUDA(ctx, e_dn);

    srf_iqa_sums acc;
    acc.sq_err = 0;
    acc.n_elems = 0;
    acc.ssim_sum = 0.0;
    acc.ssim_count = 0;
    acc.bits_a = acc.bits_b = 0;
    for (int b = 0; b < nb; ++b) {
        if (own_hi(b) <= own_lo(b)) continue;
        srf_iqa_sums s;
        srf_iqa_collect(ctx, out_w, channels, own_lo(b), own_hi(b), &s, b);
        acc.sq_err += s.sq_err;
        acc.n_elems += s.n_elems;
        acc.ssim_sum += s.ssim_sum; // fixed band order: deterministic
        acc.ssim_count += s.ssim_count;
        acc.bits_a |= s.bits_a;
        acc.bits_b |= s.bits_b;
    }
    *h_out = acc;
    ctx->last_h2d = h2d + (size_t)out_w * channels * (size_t)(F1 - F0);
    ctx->last_d2h = (size_t)out_w * channels * (size_t)std::max(row_end - row_begin, 0) + (size_t)nb * 40;
    return SRF_OK;
}
