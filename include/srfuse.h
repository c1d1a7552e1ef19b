/*
 * srfuse.h -- C ABI of libsrfuse.so: the B200 (sm_100a) tile -> Laplacian-fuse -> PSNR/SSIM path.
 *
 * The reference (squellww/super-resolution-system) is pure Python and has no FFI of its own; each entry point
 * below replaces the body of one reference method and is what a ctypes binding inside that method calls
 * (INTEGRATION.md shows the stubs).  Reference line numbers are into /root/reference.
 *
 * Conventions
 *   - plain pointers and sizes only; images are uint8, HWC interleaved (C = 1 or 3), row pitch in BYTES;
 *   - every function returns SRF_OK (0) or a negative srf_status; srf_last_error(ctx) has the text;
 *   - "_host" entry points take HOST buffers, do the host<->device copies themselves and return when the
 *     result is in the caller's buffer; the others take DEVICE pointers, enqueue on `stream`
 *     (a cudaStream_t passed as void*, NULL = the legacy default stream) and do not synchronise;
 *   - the library owns only the workspace inside srf_ctx; callers own every buffer they pass;
 *   - one srf_ctx per device per host thread; calls on one ctx must not overlap.  Several contexts on one device are
 *     fine and independent (own workspaces, own captured graph): a caller that keeps two fusions in flight -- the fuse of
 *     step k + 1 while step k is still being assessed on another stream -- gives each its own context;
 *   - there is NO CPU fallback: without a usable sm_100-class device srf_create fails.
 */
#ifndef SRFUSE_H_
#define SRFUSE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRF_VERSION 100 /* 0.1.0 */
#define SRF_MAX_LEVELS 16

typedef enum srf_status {
    SRF_OK = 0,
    SRF_ERR_ARG = -1,     /* bad argument (the Python mirror raises ValueError) */
    SRF_ERR_CUDA = -2,    /* a CUDA call failed */
    SRF_ERR_NO_DEVICE = -3,
    SRF_ERR_OOM = -4,
    SRF_ERR_UNSUPPORTED = -5
} srf_status;

/* tiling_module.py:40-45 PaddingMode */
typedef enum srf_border {
    SRF_BORDER_MIRROR = 0,    /* 'mirror'    -> cv2.BORDER_REFLECT_101 */
    SRF_BORDER_REPLICATE = 1, /* 'replicate' -> cv2.BORDER_REPLICATE   */
    SRF_BORDER_REFLECT = 2,   /* 'reflect'   -> cv2.BORDER_REFLECT     */
    SRF_BORDER_CONSTANT = 3   /* 'constant'  -> cv2.BORDER_CONSTANT, value 0 */
} srf_border;

/* blending_module.py:52-57 WeightType */
typedef enum srf_weight { SRF_WEIGHT_LINEAR = 0, SRF_WEIGHT_COSINE = 1, SRF_WEIGHT_SIGMOID = 2 } srf_weight;

typedef struct srf_ctx srf_ctx;

/* ---- context ------------------------------------------------------------------------------------------- */
int srf_version(void);
const char *srf_status_string(int status);
const char *srf_last_error(const srf_ctx *ctx);
int srf_create(int device, srf_ctx **out);
void srf_destroy(srf_ctx *ctx);
int srf_sync(srf_ctx *ctx, void *stream);
/* number of kernels this ctx has launched since creation (bench.py's gpu_launches) */
int64_t srf_launch_count(const srf_ctx *ctx);
/* how many srf_laplacian_fuse calls were served by launching the captured CUDA graph of an identical earlier call */
int64_t srf_graph_replays(const srf_ctx *ctx);
/* bytes the last *_host entry point copied host->device / device->host (image data only) */
int srf_last_transfer_bytes(const srf_ctx *ctx, size_t *h2d, size_t *d2h);
/* bytes of device workspace currently held */
size_t srf_workspace_bytes(const srf_ctx *ctx);
/* CUDA-event time (ms) of the device work of the last *_host / srf_laplacian_fuse / srf_psnr_ssim call,
 * valid after that call's stream has been synchronised; <0 if none */
float srf_last_device_ms(srf_ctx *ctx);

/* per-kernel CUDA-event timing (bench.py's roofline leg).  While enabled every kernel launch is bracketed by two
 * events on its stream; srf_profile_read synchronises the device, writes up to `cap` entries in launch order,
 * clears the log and returns the number of entries it held. */
typedef struct srf_profile_entry {
    char name[40];
    float ms;     /* device time of this launch */
    double bytes; /* algorithmic bytes of this launch (compulsory reads + writes), 0 if not tracked */
} srf_profile_entry;
int srf_profile_enable(srf_ctx *ctx, int on);
int srf_profile_read(srf_ctx *ctx, srf_profile_entry *out, int cap);

/* ---- tiling -------------------------------------------------------------------------------------------- */
/* replaces TilingModule._calculate_tile_positions (tiling_module.py:572-608).  Host-only, exact.
 * Writes min(count, cap) quadruples (x, y, w, h), row-major, and returns the count (or <0). */
int srf_tile_positions(int width, int height, int block, int overlap, int *xywh, int cap);
/* replaces TilingModule._calculate_overlap_for_tile (tiling_module.py:610-646).  out = top,bottom,left,right */
int srf_tile_overlaps(int x, int y, int w, int h, int width, int height, int block, int overlap, int *out4);
/* replaces the crop + cv2.copyMakeBorder loop of TilingModule.split_image (tiling_module.py:713-724, 522-570):
 * tile i = src[y:y+h, x:x+w] padded bottom/right to (block, block, C); tiles_out is n*block*block*C bytes */
int srf_tile_extract_pad(srf_ctx *ctx, const uint8_t *d_src, int height, int width, int channels, size_t pitch,
                         const int *xywh, int n, int block, int border, uint8_t *d_tiles_out, void *stream);
int srf_tile_extract_pad_host(srf_ctx *ctx, const uint8_t *h_src, int height, int width, int channels,
                              size_t pitch, const int *xywh, int n, int block, int border, uint8_t *h_tiles_out);

/* replaces TilingModule._apply_padding (tiling_module.py:522-570) in full generality: cv2.copyMakeBorder(image, top,
 * bottom, left, right, mode); dst is (h + top + bottom, w + left + right, channels), packed rows.  Padding may exceed
 * the image size (multi-bounce reflection, as OpenCV). */
int srf_copy_make_border(srf_ctx *ctx, const uint8_t *d_src, int h, int w, int channels, size_t pitch, int top, int bottom,
                         int left, int right, int border, uint8_t *d_dst, void *stream);
int srf_copy_make_border_host(srf_ctx *ctx, const uint8_t *h_src, int h, int w, int channels, size_t pitch, int top,
                              int bottom, int left, int right, int border, uint8_t *h_dst);

/* ---- pyramids (float32, HWC interleaved, packed rows -- the arrays the reference hands to cv2) ---------- */
/* replaces cv2.pyrDown in BlendingModule.build_gaussian_pyramid (blending_module.py:263);
 * dst is ((h+1)/2, (w+1)/2, C) */
int srf_pyr_down_f32(srf_ctx *ctx, const float *d_src, int h, int w, int channels, float *d_dst, void *stream);
/* replaces cv2.pyrUp(dstsize=) in build_laplacian_pyramid / collapse_laplacian_pyramid
 * (blending_module.py:306, 358); dst_h in {2h, 2h-1}, dst_w in {2w, 2w-1} */
int srf_pyr_up_f32(srf_ctx *ctx, const float *d_src, int h, int w, int channels, float *d_dst, int dst_h,
                   int dst_w, void *stream);
int srf_pyr_down_f32_host(srf_ctx *ctx, const float *h_src, int h, int w, int channels, float *h_dst);
int srf_pyr_up_f32_host(srf_ctx *ctx, const float *h_src, int h, int w, int channels, float *h_dst, int dst_h,
                        int dst_w);

/* ---- Laplacian fusion ---------------------------------------------------------------------------------- */
/* one upscaled tile (blending_module.py:96-112 TileInfo: image, x, y) */
typedef struct srf_tile {
    const uint8_t *data; /* (h, w, C) uint8; device pointer (host pointer for the _host entry point) */
    size_t pitch;        /* bytes between rows */
    int h, w;
    int y, x;            /* canvas position of the tile's top-left pixel (may overhang the canvas) */
    int lut;             /* index into the srf_weight_lut array */
} srf_tile;

/* W_0(y,x) = values[min(d, len-1)], d = min distance to the tile border; the 1-D profile of
 * BlendingModule._create_distance_weight_map (blending_module.py:508-561), len = feather_width + 1 */
typedef struct srf_weight_lut {
    const float *values; /* HOST pointer */
    int len;
} srf_weight_lut;

/* helper for C callers: fills feather+1 floats with the reference profile (float64 math, cast to float32) */
int srf_weight_profile(int weight_type, int feather, float *out);

/* replaces BlendingModule.laplacian_fusion (blending_module.py:369-506): per-tile Gaussian/Laplacian pyramids
 * (levels arrays, tile-local borders), per-level multiply by the tile's weight pyramid, per-tile collapse,
 * accumulation in list order, divide by max(sum W_0, 1e-6), clip, TRUNCATING cast to uint8.
 * Only canvas rows [row_begin, row_end) are produced (band sharding; 0, out_h for everything); rows outside
 * are left untouched.  canvas is (out_h, out_w, C) with `canvas_pitch` bytes per row. */
int srf_laplacian_fuse(srf_ctx *ctx, const srf_tile *tiles, int n_tiles, int channels, int levels,
                       const srf_weight_lut *luts, int n_luts, int out_h, int out_w, uint8_t *d_canvas,
                       size_t canvas_pitch, int row_begin, int row_end, void *stream);
int srf_laplacian_fuse_host(srf_ctx *ctx, const srf_tile *tiles, int n_tiles, int channels, int levels,
                            const srf_weight_lut *luts, int n_luts, int out_h, int out_w, uint8_t *h_canvas,
                            size_t canvas_pitch, int row_begin, int row_end);

/* host only, exact: the rows [row0, row1) and columns [0, cols) of a tile (h, w at canvas position y, x) that canvas rows
 * [row_begin, row_end) of an out_h x out_w fusion with `levels` pyramid levels can depend on -- the dependency cone the
 * host entry points upload, exposed so a band-sharded caller can upload each tile row to ONE rank and ship the few halo
 * rows GPU to GPU (distributed.BandedHostFuser).  row0 == row1: the tile does not reach those rows. */
int srf_fuse_input_rows(int tile_h, int tile_w, int tile_y, int tile_x, int levels, int out_h, int out_w, int row_begin,
                        int row_end, int *row0, int *row1, int *cols);

/* diagnostic, host only, no context, no device work: the plan the canvas gather of srf_laplacian_fuse uses for this
 * geometry -- the band's 128 x 16-pixel blocks (rows counted from max(row_begin, 0)) cut into rectangular cells that share
 * one ordered list of covering tiles and one class.  Writes min(*n_cells, cap) records of 12 ints
 * {class, bx0, by0, nbx, nby, T, tile0, tile1, tile2, tile3, first, 0} and the number of cells.  Classes: 0 one tile,
 * interior, unit weight; 1 one tile interior; 2 two tiles interior; 3 three / four tiles interior; 4 / 5 anything else
 * with <= 2 / <= 4 tiles (generic arithmetic); 6 / 7 cells with a tile, canvas or band edge inside them that still run
 * the interior arithmetic (<= 2 / <= 4 tiles).  SRF_ERR_UNSUPPORTED when the call would take the general kernel instead
 * (more than 48 tiles, a tile with fewer than 3 pyramid levels, more than 4 tiles over one block, an empty band).
 * tiles[i].data is not read. */
int srf_gather_plan(const srf_tile *tiles, int n_tiles, int channels, int levels, const srf_weight_lut *luts, int n_luts,
                    int out_h, int out_w, int row_begin, int row_end, int *cells, int cap, int *n_cells);

/* ---- full-reference IQA -------------------------------------------------------------------------------- */
/* partial sums of one image pair over a row band; additive across bands / ranks */
typedef struct srf_iqa_sums {
    int64_t sq_err;     /* sum (a-b)^2 over every element of rows [row_begin,row_end)            (PSNR) */
    int64_t n_elems;    /* elements in that sum                                                         */
    double ssim_sum;    /* sum of the SSIM map over interior pixels (>=5 from every image edge) whose    */
    int64_t ssim_count; /*   row lies in [row_begin,row_end); and their number                          */
    int32_t bits_a;     /* bitwise OR of every byte of a / b over the band: max() <= 1 exactly when bits <= 1 -- */
    int32_t bits_b;     /*   the caller applies the reference's `image.max() <= 1.0` rescale rule (q.a.m:192-193) */
} srf_iqa_sums;

/* ---- seam detection: the window scores of BlendingModule.detect_seams (blending_module.py:765-851) ------------------
 * For a tile placed at (tile_x, tile_y) on the fused result: the whole-patch SSIM (_compute_ssim, 853-903; float64,
 * gray by cv2.COLOR_BGR2GRAY on the stored channel order, tile first) of every window x window patch, stride `stride`,
 * of the region the two share -- out[wy * nx + wx], ny x nx as srf_window_ssim_shape reports.  c1 = (k1 L)^2,
 * c2 = (k2 L)^2.  Thresholding and the merging of neighbouring windows stay on the host. */
int srf_window_ssim_shape(int result_h, int result_w, int tile_h, int tile_w, int tile_y, int tile_x, int window,
                          int stride, int *ny, int *nx);
int srf_window_ssim(srf_ctx *ctx, const uint8_t *d_result, size_t result_pitch, int result_h, int result_w,
                    const uint8_t *d_tile, size_t tile_pitch, int tile_h, int tile_w, int channels, int tile_y, int tile_x,
                    int window, int stride, double c1, double c2, double *d_out, void *stream);
int srf_window_ssim_host(srf_ctx *ctx, const uint8_t *h_result, size_t result_pitch, int result_h, int result_w,
                         const uint8_t *h_tile, size_t tile_pitch, int tile_h, int tile_w, int channels, int tile_y,
                         int tile_x, int window, int stride, double c1, double c2, double *h_out);

/* ---- merge_tiles: feathered weighted average of the processed tiles (tiling_module.py:1074-1135) -------------------
 * out = astype(uint8)(sum_t tile_t * w_t / max(sum_t w_t, 1e-6)), float32, tiles in list order.  w_t is the map of
 * _create_blend_weight (1137-1175): float32 ones multiplied in place by float64 linspace ramps, i.e.
 * w(y, x) = f32(f32(wy[y] * wl[x]) * wr[x]) with float64 products -- the caller passes the three 1-D tables (HOST
 * pointers; all NULL = weight 1, the reference's blending=False) and the full map is never built.  A tile whose stored
 * size (h, w) differs from its canvas size (out_h, out_w) is first resized with cv2.resize's default INTER_LINEAR, as
 * the reference does (1107-1109: shrinking for the padded edge tiles, enlarging when the upscaler returned less). */
typedef struct srf_merge_tile {
    const uint8_t *data; /* (h, w, channels) uint8; host pointer for _host, device pointer otherwise */
    size_t pitch;
    int h, w;         /* stored size */
    int out_h, out_w; /* size on the canvas (TileMetadata.output_h / output_w) */
    int y, x;         /* canvas position: int(global_y * scale), int(global_x * scale) */
    const float *wy;  /* out_h row factors (top and bottom ramps applied, float32) */
    const double *wl; /* out_w left-ramp factors, 1.0 outside the ramp */
    const double *wr; /* out_w right-ramp factors, 1.0 outside the ramp */
} srf_merge_tile;
int srf_merge_tiles(srf_ctx *ctx, const srf_merge_tile *tiles, int n_tiles, int channels, int out_h, int out_w,
                    uint8_t *d_canvas, size_t canvas_pitch, void *stream);
int srf_merge_tiles_host(srf_ctx *ctx, const srf_merge_tile *tiles, int n_tiles, int channels, int out_h, int out_w,
                         uint8_t *h_canvas, size_t canvas_pitch);
/* ---- weighted_average_fusion with custom weight maps (blending_module.py:661-760, the `weights=` argument) -----------
 * out = astype(uint8)(clip(sum_t tile_t * w_t / max(sum_t w_t, 1e-6), 0, 255)), float32 accumulators, tiles in list order,
 * every tile cropped to the canvas.  w_t is an (weight_h, weight_w) plane of float32 or float64 (weight_f64 != 0) at least
 * as large as the tile's part of the canvas (top-left aligned); NumPy's mixed-precision rules are reproduced: a float64 map
 * makes the products and the in-place sums run in float64 before they are stored as float32.  Without custom maps the call
 * is srf_laplacian_fuse with levels = 1.  `srf_weighted_average` takes DEVICE pointers in data / weight and only enqueues on
 * `stream`; `_host` takes host pointers and returns the finished canvas. */
typedef struct srf_wavg_tile {
    const uint8_t *data; /* (h, w, channels) uint8 */
    size_t pitch;
    int h, w;
    int y, x; /* position on the canvas, >= 0 */
    const void *weight; /* (weight_h, weight_w) float32 or float64 */
    size_t weight_pitch; /* bytes */
    int weight_h, weight_w;
    int weight_f64;
    int reserved;
} srf_wavg_tile;
int srf_weighted_average(srf_ctx *ctx, const srf_wavg_tile *tiles, int n_tiles, int channels, int out_h, int out_w,
                         uint8_t *d_canvas, size_t canvas_pitch, void *stream);
int srf_weighted_average_host(srf_ctx *ctx, const srf_wavg_tile *tiles, int n_tiles, int channels, int out_h, int out_w,
                              uint8_t *h_canvas, size_t canvas_pitch);
/* host only, exact: OpenCV's INTER_LINEAR taps for uint8 -- per destination index the source index pair and the two
 * 11-bit coefficients (cvRound(frac * 2048)) */
int srf_resize_linear_taps(int src, int dst, int *idx2, int *coef2);
/* the same per axis: vertical != 0 gives the table of OpenCV's vertical pass (fraction kept past the first / last source
 * row, the two row indices clamped separately -- differs from the horizontal table only when enlarging) */
int srf_resize_linear_taps_axis(int src, int dst, int vertical, int *idx2, int *coef2);

/* ---- bicubic resize: cv2.resize(img, (dst_w, dst_h), interpolation=cv2.INTER_CUBIC) on uint8 ---------------------
 * replaces QualityAssessmentModule.downsample_bicubic (quality_assessment_module.py:226-253) and upsample_bicubic
 * (255-275); _evaluate_downsample_comparison (518-555) applies it to both images at 0.1x / 0.2x / 0.4x before PSNR and
 * the 7x7-window SSIM.  Keys cubic (a = -0.75) at float64 sample positions, rounded half-to-even and saturated --
 * what the opencv-python wheel (Intel IPP) computes; see oracle/resize_oracle.py for the pinning. */
/* host only, exact: for each of the dst indices of one axis the 4 clamped source indices and their float32 weights */
int srf_resize_cubic_taps(int src, int dst, int *idx4, float *w4);
int srf_resize_cubic_u8(srf_ctx *ctx, const uint8_t *d_src, size_t src_pitch, int h, int w, int channels, uint8_t *d_dst,
                        size_t dst_pitch, int dst_h, int dst_w, void *stream);
int srf_resize_cubic_u8_host(srf_ctx *ctx, const uint8_t *h_src, size_t src_pitch, int h, int w, int channels,
                             uint8_t *h_dst, size_t dst_pitch, int dst_h, int dst_w);

/* replaces calculate_psnr + calculate_ssim(multiscale=True) of QualityAssessmentModule
 * (quality_assessment_module.py:277-320, 322-376) in ONE pass over the pair: integer sum of squared
 * differences over all channels, cv2-exact RGB->gray, 11-tap sigma=1.5 Gaussian SSIM (C1 = (0.01 data_range)^2,
 * C2 = (0.03 data_range)^2, population covariance) summed over the interior.
 * window: 0 = that 11x11 Gaussian window (calculate_ssim(multiscale=True), the pipeline's 'ms_ssim');
 *         1 = scikit-image's default 7x7 uniform window with sample covariance, interior = 3 px in
 *             (calculate_ssim(multiscale=False), quality_assessment_module.py:378-384, the 'ssim' key).  This call SYNCHRONISES `stream`
 * before returning (it hands scalars back to the host).  scale_a / scale_b: multiply that image by 255 first
 * (the preprocess rule above), 0 or 1.  a and b are (h, w, C) views of equal size. */
int srf_psnr_ssim(srf_ctx *ctx, const uint8_t *d_a, size_t pitch_a, const uint8_t *d_b, size_t pitch_b, int h,
                  int w, int channels, int scale_a, int scale_b, double data_range, int window, int row_begin,
                  int row_end, srf_iqa_sums *h_out, void *stream);
int srf_psnr_ssim_host(srf_ctx *ctx, const uint8_t *h_a, size_t pitch_a, const uint8_t *h_b, size_t pitch_b,
                       int h, int w, int channels, int scale_a, int scale_b, double data_range, int window,
                       int row_begin, int row_end, srf_iqa_sums *h_out);
/* enqueue-only variant for multi-GPU use: no synchronisation, the band's sums are written to DEVICE memory as five
 * doubles {sq_err, ssim_sum, ssim_count, any byte of a > 1, any byte of b > 1} -- additive, so the caller can
 * ncclAllReduce(sum) them in place and read the totals back with one copy */
int srf_psnr_ssim_enqueue(srf_ctx *ctx, const uint8_t *d_a, size_t pitch_a, const uint8_t *d_b, size_t pitch_b, int h,
                          int w, int channels, int scale_a, int scale_b, double data_range, int window, int row_begin,
                          int row_end, double *d_sums5, void *stream);
/* replaces the PSNR / SSIM part of QualityAssessmentModule.batch_evaluate (quality_assessment_module.py:1350-1371: a serial
 * loop of evaluate_full_reference over the pairs, one result per pair): n_pairs equally sized (h, w, channels) pairs
 * scored by ONE kernel launch (grid.z = pair) and one reduction launch; out[i] holds pair i's sums over the whole image.
 * The preprocess rule (an image whose max() <= 1 is rescaled) is reported through bits_a / bits_b: the caller re-scores
 * such a pair with srf_psnr_ssim(scale_a / scale_b).  srf_psnr_ssim_batch synchronises `stream`; the _enqueue form writes
 * 5 doubles per pair ({sq_err, ssim_sum, ssim_count, any byte of a > 1, of b > 1}, as srf_psnr_ssim_enqueue) to device
 * memory and does not synchronise -- ranks that own different pairs all-gather / all-reduce that array. */
typedef struct srf_image_pair {
    const uint8_t *a; /* reference image ('original'); device pointer, host pointer for _host */
    const uint8_t *b; /* assessed image ('upscaled') */
    size_t pitch_a, pitch_b;
} srf_image_pair;
int srf_psnr_ssim_batch(srf_ctx *ctx, const srf_image_pair *pairs, int n_pairs, int h, int w, int channels, double data_range,
                        int window, srf_iqa_sums *h_out, void *stream);
int srf_psnr_ssim_batch_enqueue(srf_ctx *ctx, const srf_image_pair *pairs, int n_pairs, int h, int w, int channels,
                                double data_range, int window, double *d_sums5n, void *stream);
int srf_psnr_ssim_batch_host(srf_ctx *ctx, const srf_image_pair *pairs, int n_pairs, int h, int w, int channels,
                             double data_range, int window, srf_iqa_sums *h_out);
/* stage 3 + stage 4 of the reference pipeline in one call (main.py:347 laplacian_fusion, :369
 * evaluate_full_reference): host tiles -> fused canvas (host) + PSNR/SSIM sums of the canvas against a host
 * comparison image of the same size, for canvas rows [row_begin, row_end) (0, out_h for everything; a rank of a
 * band-sharded job passes its band -- rows outside are neither written nor scored, and the 5-row SSIM halo is fused
 * internally).  The canvas stays on the device between the two stages; uploads, kernels and the canvas download
 * overlap on three streams, band by band. */
int srf_fuse_assess_host(srf_ctx *ctx, const srf_tile *tiles, int n_tiles, int channels, int levels,
                         const srf_weight_lut *luts, int n_luts, int out_h, int out_w, uint8_t *h_canvas,
                         size_t canvas_pitch, const uint8_t *h_reference, size_t reference_pitch, double data_range,
                         int row_begin, int row_end, srf_iqa_sums *h_out);
/* ---- multi-GPU: exchange of the additive sums over NVLink peer memory (one process per GPU) -----------------------
 * The reference is a single-process CPU program and has no collectives; a band-sharded job sums
 * srf_iqa_sums over its ranks.  Each rank: srf_comm_handle() -> 64-byte CUDA-IPC handle of its inbox; the host side
 * gathers all ranks' handles (any transport: torch.distributed, MPI, a file) and passes the rank-ordered array to
 * srf_comm_init(), which maps the peers' inboxes.  srf_allreduce_metrics() then replaces d_sums[0..n) by the sum over
 * all ranks, in place, with ONE kernel per rank that writes to the peers' inboxes over NVLink and waits for theirs --
 * enqueue only; totals are summed in rank order, so every rank holds bit-identical results.  srf_comm_result()
 * synchronises `stream` and copies the totals of the last exchange from pinned host memory (the kernel wrote them
 * there too): no separate device->host copy.  All ranks must call srf_allreduce_metrics the same number of times, each
 * on its own GPU.  SRF_ERR_UNSUPPORTED from srf_comm_init = no peer access (the caller falls back to its own collective). */
#define SRF_COMM_MAX_RANKS 16
#define SRF_COMM_MAX_DOUBLES 128
#define SRF_COMM_HANDLE_BYTES 64
int srf_comm_handle(srf_ctx *ctx, void *handle64);
int srf_comm_init(srf_ctx *ctx, int rank, int world, const void *handles);
int srf_allreduce_metrics(srf_ctx *ctx, double *d_sums, int n, void *stream);
int srf_comm_result(srf_ctx *ctx, double *h_out, int n, void *stream);

/* final scalars from (all-reduced) sums: psnr = 10 log10(range^2 / (sq_err / n_elems)) (+inf if sq_err == 0),
 * ssim = ssim_sum / ssim_count */
int srf_iqa_finish(const srf_iqa_sums *sums, double data_range, double *psnr, double *ssim);

#ifdef __cplusplus
}
#endif
#endif /* SRFUSE_H_ */
