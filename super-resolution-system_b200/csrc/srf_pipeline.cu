// srf_pipeline.cu -- the reference pipeline's stage 3 + stage 4 as ONE host-facing call
// (/root/reference/main.py:347 laplacian_fusion followed by :369 evaluate_full_reference): host tiles in, host
// canvas + PSNR/SSIM sums out, with the fused canvas never leaving the device between the two stages.
// Copies and kernels overlap on two streams: the comparison image uploads while the pyramids are being built, and
// the canvas downloads while the SSIM kernel runs.
#include "srf_common.cuh"

#include <algorithm>

extern "C" int srf_fuse_assess_host(srf_ctx *ctx, const srf_tile *tiles, int n_tiles, int channels, int levels,
                                    const srf_weight_lut *luts, int n_luts, int out_h, int out_w, uint8_t *h_canvas, size_t canvas_pitch,
                                    const uint8_t *h_reference, size_t reference_pitch, double data_range, srf_iqa_sums *h_out)
{
    if (!ctx) return SRF_ERR_ARG;
    int rc = srf_check_fuse_args(ctx, tiles, n_tiles, channels, levels, luts, n_luts, out_h, out_w, h_canvas, canvas_pitch, 0, out_h);
    if (rc) return rc;
    SRF_REQUIRE(ctx, h_reference && h_out, "srf_fuse_assess_host: null pointer");
    SRF_REQUIRE(ctx, reference_pitch >= (size_t)out_w * channels, "srf_fuse_assess_host: reference pitch < row bytes");
    SRF_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));

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
    const size_t canvas_bytes = cpitch * (size_t)out_h;
    if ((rc = srf_reserve(ctx, ctx->ws_stage_in, total))) return rc;
    if ((rc = srf_reserve(ctx, ctx->ws_stage_out, 2 * canvas_bytes))) return rc; // [fused canvas | comparison image]
    cudaStream_t st = ctx->own_stream, cs = ctx->copy_stream;
    uint8_t *din = (uint8_t *)ctx->ws_stage_in.ptr;
    uint8_t *dcanvas = (uint8_t *)ctx->ws_stage_out.ptr, *dref = dcanvas + canvas_bytes;

    for (int i = 0; i < n_tiles; ++i) {
        dt[i].data = din + off[i];
        SRF_CHECK_CUDA(ctx, cudaMemcpy2DAsync(din + off[i], dt[i].pitch, tiles[i].data, tiles[i].pitch, (size_t)tiles[i].w * channels,
                                              tiles[i].h, cudaMemcpyHostToDevice, st));
    }
    // comparison image: queued behind the tiles on the copy engine, lands while the fusion kernels run
    SRF_CHECK_CUDA(ctx, cudaEventRecord(ctx->ev1, st));
    SRF_CHECK_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->ev1, 0));
    SRF_CHECK_CUDA(ctx, cudaMemcpy2DAsync(dref, cpitch, h_reference, reference_pitch, (size_t)out_w * channels, out_h, cudaMemcpyHostToDevice, cs));
    cudaEvent_t ref_ready = ctx->ev_aux[0], fused = ctx->ev_aux[1], downloaded = ctx->ev_aux[2];
    SRF_CHECK_CUDA(ctx, cudaEventRecord(ref_ready, cs));

    SRF_CHECK_CUDA(ctx, cudaEventRecord(ctx->ev0, st));
    rc = srf_fuse_device(ctx, dt.data(), n_tiles, channels, levels, luts, n_luts, out_h, out_w, dcanvas, cpitch, 0, out_h, st);
    if (rc) return rc;
    SRF_CHECK_CUDA(ctx, cudaEventRecord(fused, st));
    // canvas download on the copy stream, concurrent with the IQA kernel
    SRF_CHECK_CUDA(ctx, cudaStreamWaitEvent(cs, fused, 0));
    SRF_CHECK_CUDA(ctx, cudaMemcpy2DAsync(h_canvas, canvas_pitch, dcanvas, cpitch, (size_t)out_w * channels, out_h, cudaMemcpyDeviceToHost, cs));
    SRF_CHECK_CUDA(ctx, cudaEventRecord(downloaded, cs));
    SRF_CHECK_CUDA(ctx, cudaStreamWaitEvent(st, ref_ready, 0));
    rc = srf_iqa_enqueue(ctx, dref, cpitch, dcanvas, cpitch, out_h, out_w, channels, 0, 0, data_range, 0, out_h, st);
    if (rc) return rc;
    SRF_CHECK_CUDA(ctx, cudaEventRecord(ctx->ev1, st));
    ctx->ev_valid = true;
    SRF_CHECK_CUDA(ctx, cudaStreamSynchronize(st));
    SRF_CHECK_CUDA(ctx, cudaEventSynchronize(downloaded));
    srf_iqa_collect(ctx, out_w, channels, 0, out_h, h_out);
    return SRF_OK;
}
