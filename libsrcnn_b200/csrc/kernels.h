/* Internal: launchers of the sm_100a kernels (one per stage of the reference's doSRCNN, src/libsrcnn.cpp:628-923).
 * Host code (srcnn_ctx.cu) talks to the kernels only through these functions. */
#ifndef SRCNN_B200_KERNELS_H
#define SRCNN_B200_KERNELS_H

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace srcnn {

/* Programmatic dependent launch (PDL): a kernel launched with launch_pdl may start while the previous kernel of the stream is
 * still draining -- its prologue (weight staging, barrier / tensor-memory set-up, index arithmetic) overlaps the predecessor's
 * tail.  Contract: the predecessor executes pdl_launch_dependents() (any time), the dependent executes pdl_wait() before its
 * first access to memory the predecessor writes.  Both are no-ops in a normally launched kernel.  SRCNN_B200_PDL=0 turns the
 * attribute off. */
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

/* per-axis contribution table on the device (built on the host in double, contrib.cpp) */
struct AxisTableDev {
    const int*    left;     /* [dst]  first source index           */
    const int*    count;    /* [dst]  number of taps (right-left+1) */
    const double* w;        /* [dst][win]                           */
    int           win;
    /* dense form (NULL if unavailable): exactly win in-bounds taps from dleft, zero weights where unused */
    const int*    dleft;
    const double* dw;
    /* compact form (NULL if unavailable): cwin taps from cleft, zero-weight taps trimmed (contrib.h) */
    const int*    cleft;
    const double* cw;
    int           cwin;
    const double* cwt;      /* compact weights transposed: [cwin][dst] (16-byte aligned; NULL without the compact form) */
    int           exact29;  /* all weights have <= 29 significant bits: fma == mul, add */
    int           pow2pair; /* cwin == 2, both weights exact in FP32, one of each pair 0 or a power of two (prep.cu) */
    const void*   pack2;    /* cwin == 2 only: [dst] int4 { left + idx(tap A), bits(wA), bits(wB), left + idx(tap B) } with the weights as
                               FP32 (exact for pow2pair tables, rounded otherwise) (color.cu) */
};

/* geometry of the FP16 conv-input plane consumed by the tensor-core kernel (see DESIGN.md "Y16 plane") */
struct Y16Geom {
    int pitch;      /* halves per row, multiple of 64                                  */
    int rows;       /* strip rows + 2*6                                                */
    int bands;      /* ceil(dw / 124) column bands                                     */
};
constexpr int kY16PadL   = 8;    /* plane column = image column + 8                    */
constexpr int kHaloRows  = 6;    /* conv I (4) + conv III (2)                          */
constexpr int kBandCols  = 124;  /* output columns per 128-lane band                   */

Y16Geom y16_geometry(unsigned dw, unsigned strip_rows);

/* colour-in: u8 interleaved -> d planar FP32 planes of `npix` pixels each (plane p at planes + p*plane_stride) */
void launch_color_in(const uint8_t* src, size_t npix, unsigned d, float* planes, size_t plane_stride,
                     cudaStream_t s);

/* one separable pass along rows (vertical) or columns (horizontal) for nplanes planes at once.
 * Plane 0 uses table ty, planes >= 1 use table tc (reference src/libsrcnn.cpp:675-714).
 *  vertical  : in  [in_rows ][width] with in-row 0 = global source row in_row0
 *              out [out_rows][width] with out-row 0 = global destination row out_row0
 *  horizontal: in  [rows][sw] -> out [rows][dw]                                                         */
void launch_resize_vertical(const float* in, size_t in_plane_stride, unsigned width, unsigned in_row0,
                            float* out, size_t out_plane_stride, unsigned out_row0, unsigned out_rows,
                            unsigned nplanes, AxisTableDev ty, AxisTableDev tc, cudaStream_t s);
void launch_resize_horizontal(const float* in, size_t in_plane_stride, unsigned sw,
                              float* out, size_t out_plane_stride, unsigned dw, unsigned rows,
                              unsigned nplanes, AxisTableDev ty, AxisTableDev tc, cudaStream_t s);

/* fused front end (prep.cu): colour-in + vertical + horizontal resampling + conv-input pack in one kernel, for the
 * upscale order (vertical pass first) with compact tables.  src = device pointer at source row src_row0; planes = d FP32
 * planes whose row 0 is destination row e0 (plane 0 only written if write_y32); y16 may be NULL (no conv-input plane).
 * h_* = the host copies of the compact `left` arrays (tile footprints).  Returns 1 if launched, 0 if this geometry is
 * not covered (the caller then runs the separate kernels), -1 if the launch itself failed. */
int launch_prep(const uint8_t* src, unsigned w, unsigned d, unsigned src_row0, unsigned ow, unsigned oh,
                unsigned e0, unsigned e1, float* planes, size_t plane_stride, bool write_y32,
                __half* y16, Y16Geom g, unsigned strip_row0,
                const AxisTableDev& vy, const AxisTableDev& vc, const AxisTableDev& hy, const AxisTableDev& hc,
                const int* h_vyl, const int* h_vcl, const int* h_hyl, const int* h_hcl, cudaStream_t s,
                float* csrc = nullptr, size_t csrc_stride = 0,
                unsigned nframes = 1, size_t src_fstride = 0, size_t y16_fstride = 0, size_t csrc_fstride = 0);
/* csrc (optional, honoured only when the chroma tables are "pow2pair"): Cb / Cr / A are NOT resampled; their colour-converted
 * FP32 samples are stored at source resolution, plane k at csrc + k * csrc_stride, row 0 = source row src_row0, pitch w,
 * for launch_color_out_up to resample.  prep_chroma_handover() tells whether a set of tables takes that route.
 * nframes > 1 (hand-over route only): a group of same-geometry frames in one launch, frame f at src + f * src_fstride (bytes),
 * y16 + f * y16_fstride (halves), csrc + f * csrc_fstride (floats). */
/* approx_ok (tensor-core conv mode only): any pair of two-tap chroma tables qualifies -- colour-out then resamples with one FP32
 * fused multiply-add per tap pair, which is the correctly rounded exact sum for the power-of-two tables (x2) and within ~1e-7
 * relative of the reference's double accumulation otherwise (other magnifications; the gate there is the final 1 LSB / 50 dB) */
bool prep_chroma_handover(const AxisTableDev& vc, const AxisTableDev& hc, bool approx_ok);
/* Second-generation front end for the hand-over route (prep.cu: prep2_kernel): same outputs as launch_prep with csrc given,
 * y32 (optional) = the FP32 interpolated Y plane, row 0 = destination row e0; src_rows = source rows present from src_row0 on.  fused_mac: one double FMA per tap (product
 * mode, Y within ~1e-16 relative of the reference before its rounding to FP32) instead of the reference's separately rounded
 * multiply and add (bit-exact).  pairs_ok = prep2_pairs_ok(host copy of hy.cleft).  Returns 1 launched, 0 not covered, -1 the
 * launch failed. */
bool prep2_pairs_ok(const int* h_hyl, unsigned ow);
int launch_prep2(const uint8_t* src, unsigned w, unsigned d, unsigned src_row0, unsigned src_rows, unsigned ow, unsigned oh,
                 unsigned e0, unsigned e1, float* y32, bool fused_mac, __half* y16, Y16Geom g, unsigned strip_row0,
                 const AxisTableDev& vy, const AxisTableDev& vc, const AxisTableDev& hy, const AxisTableDev& hc,
                 const int* h_vyl, const int* h_vcl, const int* h_hyl, const int* h_hcl, bool pairs_ok, cudaStream_t s,
                 float* csrc, size_t csrc_stride,
                 unsigned nframes = 1, size_t src_fstride = 0, size_t y16_fstride = 0, size_t csrc_fstride = 0);
/* host-side applicability test of launch_prep for a set of tables */
bool prep_supported(const AxisTableDev& vy, const AxisTableDev& vc, const AxisTableDev& hy, const AxisTableDev& hc);

/* conv-input pack: FP32 Y window (rows [win_row0, win_row0+win_rows) of a dw x dh plane) -> replicate-padded FP16
 * plane covering global rows strip_row0-6 .. strip_row1+6 (clamped) and columns -8 .. pitch-9 (clamped) */
void launch_pack_y16(const float* y, unsigned dw, unsigned dh, unsigned win_row0,
                     __half* y16, Y16Geom g, unsigned strip_row0, cudaStream_t s);

/* fused conv I+II+III, bit-exact FP32 CUDA-core path (validation mode).
 * y: window rows [win_row0, ...) of the dw x dh interpolated plane; out: rows [row0,row1) x dw */
int launch_conv_exact(const float* y, unsigned dw, unsigned dh, unsigned win_row0, unsigned win_rows,
                       float* out, unsigned row0, unsigned row1, cudaStream_t s);
int  conv_exact_init();    /* current device: uploads the FP32 parameter tables to __constant__ memory, opts the kernel in to its
                              dynamic shared memory size */

/* fused conv I+II+III, tcgen05 / TMEM / TMA path (product mode) -- conv_tc.cu */
struct ConvTcState;                         /* per-context: device weights image, tensor maps */
int  conv_tc_create(ConvTcState** st);      /* 0 ok */
void conv_tc_destroy(ConvTcState* st);
/* optional fused store stage: colour-out of the strip rows straight from the conv III registers (the reference's
 * memcpy + convertImgF32XtoImgU8 + conv plane, src/libsrcnn.cpp:855-861, 889-915).  All pointers address row `row0`. */
struct ConvTcColor {
    const float* cb = nullptr; const float* cr = nullptr; const float* al = nullptr;   /* resized chroma / alpha planes */
    uint8_t* rgb = nullptr;      /* interleaved u8 result (d = 3 or 4), NULL = no fused colour-out                    */
    uint8_t* conv8 = nullptr;    /* u8 conv plane or NULL                                                           */
    unsigned d = 3;
};
/* out (FP32 Y', rows [row0,row1) x dw) may be NULL when col.rgb is given */
/* nframes > 1: a group of same-geometry frames in ONE launch (the persistent CTAs then amortise pipeline fill / drain over
 * all of them -- what small frames need): frame f's conv-input plane follows frame f-1's (g.rows rows each), its Y' lies at
 * out + f * out_fstride */
int  launch_conv_tc(ConvTcState* st, const __half* y16, Y16Geom g, unsigned dw, unsigned dh,
                    float* out, unsigned row0, unsigned row1, int num_sms, cudaStream_t s,
                    const ConvTcColor* col = nullptr, unsigned nframes = 1, size_t out_fstride = 0);
const char* conv_tc_error(const ConvTcState* st);

/* colour-out: Y' + Cb/Cr/(A) -> interleaved u8 (+ optional u8 conv plane); all arrays are `npix` long views */
/* colour-out with the x2 chroma / alpha resampling folded in (color.cu): csrc = source-resolution FP32 Cb, Cr (, A) planes
 * written by launch_prep (plane k at csrc + k * cstride, row 0 = source row src_row0, pitch w); yp / out / conv = rows
 * [row0,row1) of the strip.  Returns false (nothing launched) if the geometry is not covered. */
bool launch_color_out_up(const float* yp, const float* csrc, size_t cstride, unsigned w, unsigned src_row0,
                         const AxisTableDev& vc, const AxisTableDev& hc, unsigned ow, unsigned row0, unsigned row1, unsigned d,
                         uint8_t* out, uint8_t* conv, cudaStream_t s, int num_sms,
                         unsigned nframes = 1, size_t yp_fstride = 0, size_t csrc_fstride = 0, size_t out_fstride = 0,
                         size_t conv_fstride = 0);
/* host-side: the source columns any CTA of that kernel reads fit its shared-memory row (h_hcl = compact lefts of the table) */
bool color_out_up_quads_ok(const int* h_hcl, unsigned ow);
void launch_color_out(const float* yprime, const float* cb, const float* cr, const float* a, size_t npix,
                      unsigned d, uint8_t* out, uint8_t* conv, cudaStream_t s);

} // namespace srcnn
#endif
