/* srcnn_b200.h -- the thin extern "C" layer under the C++ drop-in API (include/libsrcnn.h).
 *
 * Plain pointers and sizes only; this is what a foreign-function binding (ctypes, cgo, JNI ...) attaches to,
 * and what ProcessSRCNN itself is written on.  Every entry point names the reference code it replaces
 * (paths relative to the rageworx/libsrcnn root).  All work runs as hand-written sm_100a CUDA kernels on the
 * context's device and stream.  There is NO CPU fallback: without a usable device srcnn_b200_create fails with
 * SRCNN_B200_ENODEV and nothing else can be called.
 *
 * Images are tightly packed row-major interleaved 8-bit RGB (d = 3) or RGBA (d = 4).  Output size of one pass is
 * ow = (unsigned)((float)w * multiply), oh likewise (reference src/libsrcnn.cpp:662-663) -- srcnn_b200_output_dims.
 * `filter` uses the SRCNNFilterType numbering (0 nearest, 1 bilinear, 2 bicubic, 3 Lanczos-3, 4 B-spline).
 */
#ifndef SRCNN_B200_H
#define SRCNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)     /* the library itself is built with -fvisibility=hidden */
#endif

#define SRCNN_B200_ABI_VERSION 1

/* return codes (0 / -1 / -2 / -100 are the reference's own, src/libsrcnn.cpp:952,965,636) */
#define SRCNN_B200_OK        0
#define SRCNN_B200_EINVAL   (-1)    /* NULL input, zero size, d not in {3,4}, bad row range */
#define SRCNN_B200_ESIZE    (-2)    /* scaled size <= 0 */
#define SRCNN_B200_ENODEV   (-20)   /* no CUDA device / not an sm_100 part / context creation failed */
#define SRCNN_B200_ECUDA    (-21)   /* a CUDA call or kernel failed; srcnn_b200_last_error has the text */
#define SRCNN_B200_ENOTHING (-100)  /* nothing produced (step mode, effective factor 1) */

/* which arithmetic the fused convolution kernel uses */
#define SRCNN_B200_CONV_TENSOR 0    /* tcgen05: FP16 operands (RN), FP32 accumulate in TMEM -- the product path   */
#define SRCNN_B200_CONV_EXACT  1    /* CUDA-core FP32 in the reference's own summation order: bit-exact check path */

typedef struct srcnn_b200_ctx srcnn_b200_ctx;

int         srcnn_b200_abi_version(void);
int         srcnn_b200_device_count(void);
/* one context = one GPU + one stream + grow-only device/pinned buffers.  Not thread-safe; use one per thread.
 * Limits: output and source heights <= 65535 rows per pass (rows map to gridDim.y); larger images: row strips. */
int         srcnn_b200_create(int device, srcnn_b200_ctx** out);
void        srcnn_b200_destroy(srcnn_b200_ctx* ctx);
const char* srcnn_b200_last_error(const srcnn_b200_ctx* ctx);
int         srcnn_b200_set_conv_mode(srcnn_b200_ctx* ctx, int mode);
int         srcnn_b200_get_conv_mode(const srcnn_b200_ctx* ctx);

/* size arithmetic.  _output_dims: one pass (src/libsrcnn.cpp:662-663).  _process_dims: the whole ProcessSRCNN call
 * including the step-scale chain (src/libsrcnn.cpp:980-1061); returns the number of passes (0 = nothing produced). */
void        srcnn_b200_output_dims(unsigned w, unsigned h, float multiply, unsigned* ow, unsigned* oh);
int         srcnn_b200_process_dims(unsigned w, unsigned h, float multiply, int stepscale, unsigned* ow, unsigned* oh);
/* 1 if the whole call produces a convolution plane.  Always 1 without step-scale; in the step-scale chain only the pass with
 * cnt + 1 == repeat receives the caller's convbuff (src/libsrcnn.cpp:1024-1028), and the loop breaks before it when the
 * remaining factor is exactly 1 (x8 = x2 x2 x2 then x1): the reference then leaves *convbuff / *convbuffsz untouched. */
int         srcnn_b200_process_has_conv(unsigned w, float multiply, int stepscale);
/* the factors of the step-scale passes that actually run (src/libsrcnn.cpp:997-1008); returns their number */
int         srcnn_b200_step_factors(unsigned w, float multiply, float* factors, int maxn);

/* ---- one pass = reference libsrcnn::doSRCNN (src/libsrcnn.cpp:628-923) --------------------------------------
 * colour-in -> resize (Y: `filter`; Cb/Cr/A: bilinear, or box for nearest) -> conv I/II/III fused -> colour-out.
 * out  : ow*oh*d bytes.   conv : ow*oh bytes or NULL (convolution-output mode, src/libsrcnn.cpp:889-915).
 * Host pointers may be pageable or pinned (srcnn_b200_host_alloc); the call returns after the results landed. */
int srcnn_b200_pass_host(srcnn_b200_ctx* ctx, const uint8_t* src, unsigned w, unsigned h, unsigned d,
                         float multiply, int filter, uint8_t* out, uint8_t* conv);

/* Row-strip shard of the same pass: produce output rows [row0,row1) only, bit-identical to those rows of the
 * full-image result.  `src` is the WHOLE source image; only the source rows the strip needs are uploaded
 * (conv I +-4 and conv III +-2 -> a 6-row halo of the interpolated Y plane, recomputed, never exchanged).
 * out_rows: (row1-row0)*ow*d bytes; conv_rows: (row1-row0)*ow bytes or NULL. */
int srcnn_b200_pass_host_rows(srcnn_b200_ctx* ctx, const uint8_t* src, unsigned w, unsigned h, unsigned d,
                              float multiply, int filter, unsigned row0, unsigned row1,
                              uint8_t* out_rows, uint8_t* conv_rows);

/* Stream-ordered, NON-BLOCKING variant of srcnn_b200_pass_host: uploads, kernels and downloads are only
 * enqueued.  src / out / conv must stay valid until srcnn_b200_wait returns and should be pinned (srcnn_b200_host_alloc);
 * with pageable buffers the copies make the call wait.  Up to two calls are in flight per context: the upload of call k+1
 * overlaps the kernels and the download of call k.  srcnn_b200_wait blocks until every asynchronous pass of the context has
 * landed and returns the first CUDA error, if any.  Any blocking entry point of the same context waits first, too. */
int srcnn_b200_pass_host_async(srcnn_b200_ctx* ctx, const uint8_t* src, unsigned w, unsigned h, unsigned d,
                               float multiply, int filter, uint8_t* out, uint8_t* conv);
int srcnn_b200_wait(srcnn_b200_ctx* ctx);

/* Host-side geometry of a strip (no GPU needed): the rows [*halo_row0,*halo_row1) of the interpolated planes the strip
 * computes (strip +- 6 clipped to the image) and the source rows [*src_row0,*src_row1) it reads / uploads. */
int srcnn_b200_strip_geometry(unsigned w, unsigned h, unsigned d, float multiply, int filter,
                              unsigned row0, unsigned row1, unsigned* halo_row0, unsigned* halo_row1,
                              unsigned* src_row0, unsigned* src_row1);

/* Device-resident variant: d_src (whole image), d_out / d_conv (rows [row0,row1), d_conv may be NULL) are DEVICE
 * pointers on the context's device.  Work is enqueued on the context's stream; call srcnn_b200_sync (or use
 * srcnn_b200_stream) to order against it.  This is the kernel-throughput path the benchmark times. */
int srcnn_b200_pass_device(srcnn_b200_ctx* ctx, const void* d_src, unsigned w, unsigned h, unsigned d,
                           float multiply, int filter, unsigned row0, unsigned row1,
                           void* d_out, void* d_conv);
/* CUDA graphs for srcnn_b200_pass_device (default off; SRCNN_B200_GRAPH=1): a pass repeated with the same geometry and the same
 * pointers is captured on its second call (front end -> fused conv -> colour-out, programmatic-dependent-launch edges included)
 * and replayed afterwards with a single cudaGraphLaunch.  srcnn_b200_graph_replays counts the replays. */
int   srcnn_b200_set_graphs(srcnn_b200_ctx* ctx, int on);
unsigned long long srcnn_b200_graph_replays(const srcnn_b200_ctx* ctx);
int   srcnn_b200_sync(srcnn_b200_ctx* ctx);
void* srcnn_b200_stream(srcnn_b200_ctx* ctx);            /* cudaStream_t */

/* ---- whole call = reference ProcessSRCNN (src/libsrcnn.cpp:943-1064) -----------------------------------------
 * incl. the step-scale chain; 8-bit intermediates of the chain stay in HBM (quantisation preserved).
 * out must hold the srcnn_b200_process_dims size * d; conv (or NULL) ow*oh of the LAST pass -- written only when
 * srcnn_b200_process_has_conv says the reference would produce it. */
int srcnn_b200_process_host(srcnn_b200_ctx* ctx, const uint8_t* src, unsigned w, unsigned h, unsigned d,
                            float multiply, int filter, int stepscale, uint8_t* out, uint8_t* conv);

/* Frame batch (independent frames, same geometry): frames = n*w*h*d bytes, out = n*ow*oh*d, conv = n*ow*oh or
 * NULL.  Uploads, kernels and downloads of consecutive frames overlap on two streams with pinned staging. */
int srcnn_b200_batch_host(srcnn_b200_ctx* ctx, const uint8_t* frames, unsigned n, unsigned w, unsigned h,
                          unsigned d, float multiply, int filter, uint8_t* out, uint8_t* conv);

/* The same batch with DEVICE pointers (frames, out, conv on the context's device), enqueued on the context's stream like
 * srcnn_b200_pass_device.  Small frames are processed in groups whose fused convolution is ONE kernel launch over all
 * frames of the group (a persistent kernel is not filled by a single small frame); large frames run one by one. */
int srcnn_b200_batch_device(srcnn_b200_ctx* ctx, const void* d_frames, unsigned n, unsigned w, unsigned h,
                            unsigned d, float multiply, int filter, void* d_out, void* d_conv);

/* ---- one host call, ALL GPUs of the box (single process; reference: the OpenMP teams inside ProcessSRCNN,
 * src/libsrcnn.cpp:791, 817, 495, spread one call over every host core) ------------------------------------------
 * srcnn_b200_multi owns one context + one worker thread per device.  A pass over one image is cut into row strips
 * [oh*g/G, oh*(g+1)/G), one per GPU, each recomputing its 6-row halo (no exchange, results bit-identical to the
 * single-GPU call); a batch is cut into contiguous blocks of frames.  devices == NULL / ndev <= 0: every visible device.
 * Calls on one handle are serialised.  Images with fewer than SRCNN_B200_MULTI_MIN_ROWS (256) output rows per GPU use
 * fewer GPUs.  Step-scale chains run every pass on all GPUs, with the 8-bit intermediates in pinned host memory. */
typedef struct srcnn_b200_multi srcnn_b200_multi;
int         srcnn_b200_multi_create(const int* devices, int ndev, srcnn_b200_multi** out);
void        srcnn_b200_multi_destroy(srcnn_b200_multi* m);
int         srcnn_b200_multi_device_count(const srcnn_b200_multi* m);
const char* srcnn_b200_multi_last_error(const srcnn_b200_multi* m);
int         srcnn_b200_multi_set_conv_mode(srcnn_b200_multi* m, int mode);
int srcnn_b200_multi_pass_host(srcnn_b200_multi* m, const uint8_t* src, unsigned w, unsigned h, unsigned d,
                               float multiply, int filter, uint8_t* out, uint8_t* conv);
int srcnn_b200_multi_process_host(srcnn_b200_multi* m, const uint8_t* src, unsigned w, unsigned h, unsigned d,
                                  float multiply, int filter, int stepscale, uint8_t* out, uint8_t* conv);
int srcnn_b200_multi_batch_host(srcnn_b200_multi* m, const uint8_t* frames, unsigned n, unsigned w, unsigned h,
                                unsigned d, float multiply, int filter, uint8_t* out, uint8_t* conv);

/* pinned host memory for fast DMA (cudaHostAlloc / cudaFreeHost) */
void* srcnn_b200_host_alloc(size_t bytes);
void  srcnn_b200_host_free(void* p);

/* ---- introspection used by the parity tests and the benchmark ------------------------------------------------ */
/* FP32 planes of the LAST pass, copied to `host` (cap floats): which = 0 interpolated Y, 1 Cb, 2 Cr, 3 A,
 * 4 network output Y'.  Planes 0-3 cover the strip's extended rows [*row0, *row0 + *rows); plane 4 the strip rows. */
int srcnn_b200_get_plane(srcnn_b200_ctx* ctx, int which, float* host, size_t cap,
                         unsigned* row0, unsigned* rows, unsigned* width);
/* The tensor-core path consumes the interpolated Y plane as FP16 and never writes its FP32 form to HBM.  Switch this
 * on (before the pass) to have plane 0 written as well, for srcnn_b200_get_plane(ctx, 0, ...).  Default off.
 *   on = 1: the FP32 Y plane AND the output-size Cb / Cr / A planes, computed exactly like the reference (two roundings per
 *           tap) -- the planes of the bit-exact route;
 *   on = 2: the FP32 Y plane only, written by the PRODUCT route's own front-end kernel (x2: one double FMA per tap) -- what
 *           the 1e-5 gate on the interpolated plane is checked on. */
int srcnn_b200_set_keep_planes(srcnn_b200_ctx* ctx, int on);
/* GPU time (ms, CUDA events on the context's stream) of the LAST pass's kernels:
 * [0] colour-in [1] resize pass 1 [2] resize pass 2 [3] conv-input pack [4] fused conv [5] colour-out [6] total.
 * Only recorded while srcnn_b200_set_timing(ctx, 1). */
int srcnn_b200_set_timing(srcnn_b200_ctx* ctx, int on);
/* host threads (1..64) that move result rows from the pinned staging slots into a pageable destination buffer */
int srcnn_b200_set_copy_threads(srcnn_b200_ctx* ctx, unsigned n);
int srcnn_b200_last_timings(srcnn_b200_ctx* ctx, float* ms7);
/* kernels launched by this context since creation */
unsigned long long srcnn_b200_launch_count(const srcnn_b200_ctx* ctx);
/* frames covered by the last pass (1) or frame group of a batch -- what srcnn_b200_last_timings then describes */
unsigned srcnn_b200_last_group(const srcnn_b200_ctx* ctx);

/* ---- single stages on host data (per-stage parity against the oracle) ---------------------------------------- */
/* reference converImgU8toYCbCr, src/libsrcnn.cpp:233-272; planes = d pointers to w*h floats */
int srcnn_b200_stage_color_in(srcnn_b200_ctx* ctx, const uint8_t* src, unsigned w, unsigned h, unsigned d,
                              float* const* planes);
/* reference FRAWResizeEngine::scale, src/frawscale.cpp:162-286 */
int srcnn_b200_stage_resize(srcnn_b200_ctx* ctx, const float* src, unsigned sw, unsigned sh,
                            unsigned dw, unsigned dh, int filter, float* dst);
/* reference convolution99 x64 -> convolution11 x32 -> convolution55 (src/libsrcnn.cpp:350-529) on a Y plane */
int srcnn_b200_stage_conv(srcnn_b200_ctx* ctx, const float* y, unsigned w, unsigned h, int mode, float* out);
/* reference convertImgF32XtoImgU8 (+ conv plane), src/libsrcnn.cpp:274-308, 889-915 */
int srcnn_b200_stage_color_out(srcnn_b200_ctx* ctx, const float* const* planes, unsigned w, unsigned h,
                               unsigned d, uint8_t* out, uint8_t* conv);
/* reference FRawScaleWeightsTable ctor, src/frawscale.cpp:8-112 (host-side, double).  weights: dst x maxwin.
 * Returns the window size, or -1 if maxwin is too small. */
int srcnn_b200_contrib_table(int filter, unsigned dst, unsigned src, int* left, int* right,
                             double* weights, int maxwin);

/* The compact form of the same table that the fused front-end kernel reads: per destination index `cwin` in-bounds
 * taps starting at cleft, with the exactly-zero weights of the reference's window trimmed away (adding +-0 to the
 * accumulator changes nothing).  *exact29 is a bit set: bit 0 = every weight has <= 29 significant bits (its product
 * with an FP32 sample is then exact in double); bit 1 = "pow2pair": cwin == 2, both weights exact in FP32 and one of each
 * pair 0 or a power of two (bilinear at x2), which lets the kernel resample chroma / alpha in FP32 with one fused
 * multiply-add per sample, bit-identically.  Returns cwin, 0 if the source is narrower than the window, -1 if maxwin is
 * too small. */
int srcnn_b200_contrib_table_compact(int filter, unsigned dst, unsigned src, int* cleft, double* weights, int maxwin,
                                     int* exact29);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* SRCNN_B200_H */
