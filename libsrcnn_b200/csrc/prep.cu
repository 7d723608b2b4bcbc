/* Fused front end of one pass: colour-in + both resampling passes + the conv-input pack, one kernel, one trip to HBM.
 *
 *   u8 RGB(A) source rows  ->  replicate-padded FP16 Y plane (conv input)  +  Cb, Cr, (A) as FP32 planes: at SOURCE size
 *                              on the x2 route (colour-out resamples them, color.cu), at output size otherwise
 *                              (+ the FP32 Y plane when asked for: bit-exact conv mode / plane getters)
 *
 * = reference converImgU8toYCbCr (src/libsrcnn.cpp:233-272) followed by FRAWResizeEngine::scale in its UPSCALE order,
 *   vertical pass then horizontal pass with an FP32 intermediate (src/frawscale.cpp:238-278, 288-385), for the plane /
 *   filter mapping of src/libsrcnn.cpp:675-714 (plane 0 = configured filter, others bilinear or box).
 *
 * A CTA owns a TH x TW tile of the output.  The source footprint of the tile (from the compact contribution tables) is
 * colour-converted once into shared memory, the vertical pass runs shared -> shared, the horizontal pass shared ->
 * registers -> global.  The 4-byte planes of the unfused sequence (source planes, intermediate plane, FP32 Y) never
 * reach HBM: at 1920x1080 -> 3840x2160 RGB the front end moves 6 + 17 + 17 MB (x2 route; 6 + 66 + 17 MB with output-size
 * chroma planes) instead of 6 + 25 + 25 + 50 + 50 + 100 + 33 + 17 MB.
 *
 * Arithmetic is that of the separate kernels (resize.cu, color.cu) and therefore bit-exact against the oracle:
 * unfused FP32 colour conversion in the reference's association order; per destination sample a double accumulator
 * from 0.0, taps ascending, separately rounded multiply and add, one rounding to float after EACH pass.  Two
 * identities cut the FP64 work without changing a bit: taps whose weight is exactly 0 are left out (compact tables,
 * contrib.h), and where every weight of a table has <= 29 significant bits (bilinear at x2: 0.25 / 0.75) the product
 * with an FP32 sample is exact, so one fused multiply-add gives the same double as multiply-then-add.  Shared memory
 * holds the planes as doubles: the (exact) float -> double conversion is paid once per element instead of per tap.
 *
 * Third identity (chroma mode 2, "pow2pair" tables: two taps, both weights exact in FP32 and one of them 0 or a power
 * of two -- bilinear at x2: 0.25 / 0.75, 1 / 0 at the borders): the product with the power-of-two weight is exact in
 * FP32, so fmaf(wA, vA, wB * vB) rounds the EXACT value wA*vA + wB*vB once to FP32.  The reference's double
 * accumulation of the same two products is exact as well (samples derived from 8-bit data span < 40 bits), and its
 * cast to float is that same single rounding.  Cb / Cr / A then need no FP64 at all: they are kept interleaved as
 * FP32 in shared memory and both resampling passes run once for all of them.
 */
#include "kernels.h"

#include <algorithm>
#include <atomic>
#include <climits>
#include <cstdlib>

namespace srcnn {

namespace {

#ifndef SRCNN_PREP_TH
#define SRCNN_PREP_TH 32
#endif
constexpr int TH = SRCNN_PREP_TH, TW = 128, NT = 256, NW = NT / 32;
constexpr int RPT = TH / (NT / TW);        /* destination rows per thread in the horizontal pass */

/* development switch (A/B timing and parity of the two chroma paths) */
inline bool getenv_flag(const char* name) { const char* e = std::getenv(name); return e && *e && *e != '0'; }

struct PrepParams {
    const uint8_t* src;         /* source row src_row0 */
    unsigned w, src_row0;
    unsigned src_rows;          /* source rows available from src_row0 on (prep2_kernel: bounds of its L2 prefetch) */
    unsigned ow, oh;
    unsigned e0, e1;            /* destination rows computed */
    float* planes; size_t plane_stride;   /* plane p row 0 = destination row e0 */
    int write_y32;
    __half* y16; int pitch, y16_rows; int strip_row0;
    const int *vyl, *vcl, *hyl, *hcl;     /* compact lefts   */
    const double *vyw, *vcw, *hyw, *hcw;  /* compact weights */
    const double *hywt, *hcwt;            /* horizontal weights transposed [tap][ow] (NULL: not available) */
    int SR, SC;                 /* shared-memory capacity of the source footprint (rows, columns) */
    /* chroma hand-over at SOURCE resolution (chroma mode 2 only, may be NULL): instead of resampling Cb / Cr / A here, the
     * colour-converted FP32 samples of the footprint are written to csrc[plane][(row - src_row0) * w + col] and the x2
     * colour-out kernel (color.cu: color_out_up_kernel) resamples them on the fly -- 4 x less chroma traffic */
    float* csrc; size_t csrc_stride;
    /* frame groups (gridDim.z = frames of identical geometry): byte / element distance between consecutive frames */
    size_t src_fstride, y16_fstride, csrc_fstride;
};

__device__ __forceinline__ void ycbcr_of(float r, float g, float b, float& y, float& cb, float& cr)
{
    /* same expression trees as color.cu (reference src/libsrcnn.cpp:245-270) */
    y  = __fadd_rn(__fadd_rn(__fmul_rn(0.299f, r), __fmul_rn(0.587f, g)), __fmul_rn(0.114f, b));
    cb = __fadd_rn(__fsub_rn(__fsub_rn(128.f, __fmul_rn(0.1687f, r)), __fmul_rn(0.3313f, g)), __fmul_rn(0.5f, b));
    cr = __fsub_rn(__fsub_rn(__fadd_rn(128.f, __fmul_rn(0.5f, r)), __fmul_rn(0.4187f, g)), __fmul_rn(0.0813f, b));
}

/* acc + w * v the way the reference does it (two roundings), or fused when the product is known to be exact */
template <bool FMA>
__device__ __forceinline__ double mac(double acc, double w, double v)
{
    return FMA ? __fma_rn(w, v, acc) : __dadd_rn(acc, __dmul_rn(w, v));
}

/* vertical pass of one plane, shared -> shared: warp = one destination row at a time, lanes = footprint columns,
 * three 32-column chunks in flight (independent accumulator chains hide the FP64 latency).  vl / vw: this tile's
 * slice of the contribution table, staged in shared memory (left already relative to the footprint). */
template <int WIN, bool FMA, int SC>
__device__ __forceinline__ void vpass(const double* __restrict__ S, double* __restrict__ T, int nc, int ny,
                                      const int* __restrict__ vl, const double* __restrict__ vw, int warp, int lane)
{
    #pragma unroll 1
    for (int y = warp; y < ny; y += NW) {
        double wv[WIN];
        #pragma unroll
        for (int i = 0; i < WIN; i++) wv[i] = vw[y * WIN + i];
        const double* s = S + vl[y] * SC;
        double* t = T + y * SC;
        #pragma unroll 1
        for (int cb = 0; cb < nc; cb += 96) {
            const int c0 = cb + lane, c1 = c0 + 32, c2 = c0 + 64;
            const int k0 = min(c0, nc - 1), k1 = min(c1, nc - 1), k2 = min(c2, nc - 1);
            double a0 = 0.0, a1 = 0.0, a2 = 0.0;
            #pragma unroll
            for (int i = 0; i < WIN; i++) {
                a0 = mac<FMA>(a0, wv[i], s[i * SC + k0]);
                a1 = mac<FMA>(a1, wv[i], s[i * SC + k1]);
                a2 = mac<FMA>(a2, wv[i], s[i * SC + k2]);
            }
            /* the reference's FP32 intermediate plane */
            if (c0 < nc) t[c0] = (double)__double2float_rn(a0);
            if (c1 < nc) t[c1] = (double)__double2float_rn(a1);
            if (c2 < nc) t[c2] = (double)__double2float_rn(a2);
        }
    }
}

/* horizontal pass of one plane for this thread's column: NR rows from T (stride SC) -> f(r, value) */
template <int WIN, bool FMA, int SC, typename F>
__device__ __forceinline__ void hpass(const double* __restrict__ t, const double (&wv)[WIN], int nrows, F&& emit)
{
    if (nrows == RPT) {
        #pragma unroll 8
        for (int r = 0; r < RPT; r++) {
            double acc = 0.0;
            #pragma unroll
            for (int i = 0; i < WIN; i++) acc = mac<FMA>(acc, wv[i], t[r * SC + i]);
            emit(r, __double2float_rn(acc));
        }
    } else {
        #pragma unroll 1
        for (int r = 0; r < nrows; r++) {
            double acc = 0.0;
            #pragma unroll
            for (int i = 0; i < WIN; i++) acc = mac<FMA>(acc, wv[i], t[r * SC + i]);
            emit(r, __double2float_rn(acc));
        }
    }
}


/* ---- chroma mode 2: two-tap FP32 resampling of NC interleaved planes (see the header comment) ---- */
__device__ __forceinline__ bool pow2_or_zero(float w) { return (__float_as_uint(w) & 0x7FFFFFu) == 0u; }
/* value = fmaf(wA, vA, wB * vB) with B the tap whose weight is 0 or a power of two */
__device__ __forceinline__ float mix2(float wA, float vA, float wB, float vB) { return __fmaf_rn(wA, vA, __fmul_rn(wB, vB)); }

template <int NC> struct CVec;
template <> struct CVec<2> { float2 v; };
template <> struct CVec<3> { float v[3]; };

template <int NC>
__device__ __forceinline__ void cmix(const float* __restrict__ a, const float* __restrict__ b, float wA, float wB, float (&o)[NC])
{
    if (NC == 2) {
        const float2 va = *reinterpret_cast<const float2*>(a), vb = *reinterpret_cast<const float2*>(b);
        o[0] = mix2(wA, va.x, wB, vb.x); o[1] = mix2(wA, va.y, wB, vb.y);
    } else {
        #pragma unroll
        for (int k = 0; k < NC; k++) o[k] = mix2(wA, a[k], wB, b[k]);
    }
}

/* vertical pass, shared -> shared: warp = destination row, lanes = footprint columns */
template <int NC, int SC>
__device__ __forceinline__ void vpass_c32(const float* __restrict__ S, float* __restrict__ T, int nc, int ny,
                                          const int* __restrict__ vl, const double* __restrict__ vw, int warp, int lane)
{
    #pragma unroll 1
    for (int y = warp; y < ny; y += NW) {
        const float w0 = (float)vw[2 * y], w1 = (float)vw[2 * y + 1];
        const bool sw = !pow2_or_zero(w1);                     /* tap 1 is the general one -> A = tap 1, B = tap 0 */
        const float wA = sw ? w1 : w0, wB = sw ? w0 : w1;
        const float* sA = S + (size_t)(vl[y] + (sw ? 1 : 0)) * (SC * NC);
        const float* sB = S + (size_t)(vl[y] + (sw ? 0 : 1)) * (SC * NC);
        float* t = T + (size_t)y * (SC * NC);
        #pragma unroll 3
        for (int c = lane; c < nc; c += 32) {
            float o[NC];
            cmix<NC>(sA + c * NC, sB + c * NC, wA, wB, o);
            if (NC == 2) *reinterpret_cast<float2*>(t + c * NC) = make_float2(o[0], o[1]);
            else { for (int k = 0; k < NC; k++) t[c * NC + k] = o[k]; }
        }
    }
}

template <int D, int WY, int WC, int CM, int SC>
__global__ void __launch_bounds__(NT)
prep_kernel(const PrepParams p)
{
    pdl_launch_dependents();                          /* the conv kernel may set itself up while this grid drains */
    /* frame group (gridDim.z > 1): this CTA's frame */
    const uint8_t* const f_src = p.src + (size_t)blockIdx.z * p.src_fstride;
    __half* const f_y16 = p.y16 ? p.y16 + (size_t)blockIdx.z * p.y16_fstride : nullptr;
    float* const f_csrc = p.csrc ? p.csrc + (size_t)blockIdx.z * p.csrc_fstride : nullptr;
    constexpr bool CFMA = (CM == 1);
    constexpr bool C32 = (CM == 2);                   /* chroma / alpha as interleaved FP32 (pow2pair tables) */
    constexpr int NC = D - 1;
    extern __shared__ double sm[];
    double* S = sm;                                   /* [D][SR][SC]  colour-converted source footprint (C32: plane 0 only) */
    double* T = sm + (C32 ? 1 : D) * p.SR * SC;       /* [TH][SC]     one plane after the vertical pass (C32: NC floats / element) */
    float* Sc = reinterpret_cast<float*>(T + (C32 && NC == 3 ? 2 : 1) * TH * SC);   /* C32: [SR][SC][NC] FP32 Cb, Cr, (A) */
    __shared__ int s_lim[4];
    __shared__ int s_vl[2][TH];                       /* this tile's rows of the vertical tables */
    __shared__ double s_vwy[TH * WY], s_vwc[TH * WC];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned x0 = blockIdx.x * TW, gy0 = p.e0 + blockIdx.y * TH;
    const int ny = (int)min((unsigned)TH, p.e1 - gy0), nx = (int)min((unsigned)TW, p.ow - x0);
    const int SR = p.SR;

    /* this thread's column of the horizontal tables (global loads overlap the footprint set-up) */
    const int x = tid & (TW - 1), rbeg = (tid / TW) * RPT;
    const unsigned gx = x0 + x;
    const bool col_ok = x < nx;
    double wy[WY], wc[WC];
    int hly = 0, hlc = 0;
    if (col_ok) {
        hly = __ldg(p.hyl + gx); hlc = __ldg(p.hcl + gx);
        /* tap-major copies of the tables when present: adjacent columns then read adjacent words (see prep2_kernel) */
        #pragma unroll
        for (int i = 0; i < WY; i++) wy[i] = p.hywt ? __ldg(p.hywt + (size_t)i * p.ow + gx) : __ldg(p.hyw + gx * WY + i);
        #pragma unroll
        for (int i = 0; i < WC; i++) wc[i] = p.hcwt ? __ldg(p.hcwt + (size_t)i * p.ow + gx) : __ldg(p.hcw + gx * WC + i);
    }

    if (tid == 0) { s_lim[0] = INT_MAX; s_lim[1] = -1; s_lim[2] = INT_MAX; s_lim[3] = -1; }
    __syncthreads();
    if (tid < ny) {
        const int a = __ldg(p.vyl + gy0 + tid), b = __ldg(p.vcl + gy0 + tid);
        s_vl[0][tid] = a; s_vl[1][tid] = b;
        atomicMin(&s_lim[0], min(a, b)); atomicMax(&s_lim[1], max(a + WY, b + WC));
    }
    if (tid >= TH && tid < TH + nx) {                 /* (a different warp than the row entries above) */
        const int a = __ldg(p.hyl + x0 + tid - TH), b = __ldg(p.hcl + x0 + tid - TH);
        atomicMin(&s_lim[2], min(a, b)); atomicMax(&s_lim[3], max(a + WY, b + WC));
    }
    for (int i = tid; i < ny * WY; i += NT) s_vwy[i] = __ldg(p.vyw + gy0 * WY + i);
    for (int i = tid; i < ny * WC; i += NT) s_vwc[i] = __ldg(p.vcw + gy0 * WC + i);
    __syncthreads();
    const int rlo = s_lim[0], nr = s_lim[1] - rlo, clo = s_lim[2], nc = s_lim[3] - clo;
    if (tid < ny) { s_vl[0][tid] -= rlo; s_vl[1][tid] -= rlo; }      /* made visible by the barrier after colour-in */

    /* 1. colour-in of the footprint: warp = source row, lanes = source columns, three chunks in flight */
    {
        const int pl = SR * SC;
        #pragma unroll 1
        for (int r = warp; r < nr; r += NW) {
            const uint8_t* row = f_src + ((size_t)(rlo + r - (int)p.src_row0) * p.w + (size_t)clo) * D;
            double* s = S + r * SC;
            #pragma unroll 1
            for (int cb = 0; cb < nc; cb += 96) {
                uint32_t v[3][D];
                #pragma unroll
                for (int k = 0; k < 3; k++) {
                    const uint8_t* px = row + min(cb + lane + 32 * k, nc - 1) * D;
                    #pragma unroll
                    for (int j = 0; j < D; j++) v[k][j] = __ldg(px + j);
                }
                #pragma unroll
                for (int k = 0; k < 3; k++) {
                    const int c = cb + lane + 32 * k;
                    float y, cb_, cr;
                    ycbcr_of((float)v[k][0], (float)v[k][1], (float)v[k][2], y, cb_, cr);
                    if (c < nc) {
                        s[c] = (double)y;
                        if (C32 && f_csrc) {
                            /* neighbouring tiles' footprints overlap: they store identical values (benign) */
                            float* q = f_csrc + (size_t)(rlo + r - (int)p.src_row0) * p.w + (size_t)(clo + c);
                            q[0] = cb_; q[p.csrc_stride] = cr;
                            if (NC == 3) q[2 * p.csrc_stride] = (float)v[k][D - 1];
                        } else if (C32) {
                            float* q = Sc + ((size_t)r * SC + c) * NC;
                            if (NC == 2) *reinterpret_cast<float2*>(q) = make_float2(cb_, cr);
                            else { q[0] = cb_; q[1] = cr; q[2] = (float)v[k][D - 1]; }
                        } else {
                            s[pl + c] = (double)cb_; s[2 * pl + c] = (double)cr;
                            if (D == 4) s[3 * pl + c] = (double)(float)v[k][D - 1];
                        }
                    }
                }
            }
        }
    }
    __syncthreads();

    const int nrows = max(0, min(ny - rbeg, RPT));
    const int ybase = (int)gy0 - p.strip_row0 + kHaloRows;       /* conv-input plane row of tile row 0 */

    /* 2. plane 0 (Y): configured filter */
    vpass<WY, false, SC>(S, T, nc, ny, s_vl[0], s_vwy, warp, lane);
    __syncthreads();
    if (col_ok) {
        const double* t = T + rbeg * SC + (hly - clo);
        float* o32 = p.planes + (size_t)(gy0 + rbeg - p.e0) * p.ow + gx;
        /* conv input: FP16 (RN), plane column = image column + 8 */
        __half* o16 = f_y16 ? f_y16 + (size_t)(ybase + rbeg) * p.pitch + gx + kY16PadL : nullptr;
        const unsigned ow = p.ow; const int pitch = p.pitch;
        if (o16 && !p.write_y32)
            hpass<WY, false, SC>(t, wy, nrows, [&](int r, float v) { o16[r * pitch] = __float2half_rn(v); });
        else
            hpass<WY, false, SC>(t, wy, nrows, [&](int r, float v) {
                if (p.write_y32) o32[(size_t)r * ow] = v;
                if (o16) o16[r * pitch] = __float2half_rn(v);
            });
    }
    __syncthreads();            /* the tile's conv-input values are visible to the whole CTA; T is free again */

    /* 2b. replicate padding of the conv-input plane around the image (only border tiles do any work):
     * columns 0..7 repeat image column 0, columns ow+8..pitch-1 repeat column ow-1, the rows above / below the image
     * repeat row 0 / oh-1 -- conv I clamps Y coordinates (src/libsrcnn.cpp:362-392) */
    if (f_y16) {
        const bool first_col = (x0 == 0), last_col = (x0 + nx == p.ow);
        if (first_col || last_col) {
            const int c_hi = (int)p.ow + kY16PadL, npad_r = last_col ? p.pitch - c_hi : 0, npad_l = first_col ? kY16PadL : 0;
            const int per_row = npad_l + npad_r;
            for (int idx = tid; idx < ny * per_row; idx += NT) {
                const int r = idx / per_row, k = idx - r * per_row;
                __half* row = f_y16 + (size_t)(ybase + r) * p.pitch;
                if (k < npad_l) row[k] = row[kY16PadL];
                else            row[c_hi + (k - npad_l)] = row[c_hi - 1];
            }
            __syncthreads();
        }
        const bool top = (gy0 == 0) && ybase > 0, bottom = (gy0 + ny == p.oh) && (ybase + ny < p.y16_rows);
        if (top || bottom) {
            /* this tile's column span of the plane, pads included */
            const int ca = first_col ? 0 : (int)x0 + kY16PadL, cb = last_col ? p.pitch : (int)x0 + nx + kY16PadL;
            const int span = cb - ca;
            if (top) {
                const __half* from = f_y16 + (size_t)ybase * p.pitch + ca;
                for (int idx = tid; idx < ybase * span; idx += NT) {
                    const int r = idx / span, c = idx - r * span;
                    f_y16[(size_t)r * p.pitch + ca + c] = from[c];
                }
            }
            if (bottom) {
                const int rb = ybase + ny, nrep = p.y16_rows - rb;
                const __half* from = f_y16 + (size_t)(rb - 1) * p.pitch + ca;
                for (int idx = tid; idx < nrep * span; idx += NT) {
                    const int r = idx / span, c = idx - r * span;
                    f_y16[(size_t)(rb + r) * p.pitch + ca + c] = from[c];
                }
            }
        }
    }

    /* 3. planes 1 .. D-1 (Cb, Cr, A): bilinear / box tables */
    if (C32 && f_csrc) {
        /* handed over at source resolution (colour-in above) */
    } else if (C32) {
        float* Tc = reinterpret_cast<float*>(T);
        vpass_c32<NC, SC>(Sc, Tc, nc, ny, s_vl[1], s_vwc, warp, lane);
        __syncthreads();
        if (col_ok) {
            const float w0 = (float)wc[0], w1 = (float)wc[1];
            const bool sw = !pow2_or_zero(w1);
            const float wA = sw ? w1 : w0, wB = sw ? w0 : w1;
            const float* tA = Tc + ((size_t)rbeg * SC + (hlc - clo) + (sw ? 1 : 0)) * NC;
            const float* tB = Tc + ((size_t)rbeg * SC + (hlc - clo) + (sw ? 0 : 1)) * NC;
            float* out = p.planes + p.plane_stride + (size_t)(gy0 + rbeg - p.e0) * p.ow + gx;
            const size_t ps = p.plane_stride; const unsigned ow = p.ow;
            auto body = [&](int r) {
                float o[NC];
                cmix<NC>(tA + (size_t)r * (SC * NC), tB + (size_t)r * (SC * NC), wA, wB, o);
                #pragma unroll
                for (int k = 0; k < NC; k++) out[k * ps + (size_t)r * ow] = o[k];
            };
            if (nrows == RPT) {
                #pragma unroll 8
                for (int r = 0; r < RPT; r++) body(r);
            } else {
                #pragma unroll 1
                for (int r = 0; r < nrows; r++) body(r);
            }
        }
    } else {
        const double* tc = T + rbeg * SC + (hlc - clo);
        #pragma unroll 1
        for (int pl = 1; pl < D; pl++) {
            vpass<WC, CFMA, SC>(S + pl * SR * SC, T, nc, ny, s_vl[1], s_vwc, warp, lane);
            __syncthreads();
            if (col_ok) {
                float* out = p.planes + pl * p.plane_stride + (size_t)(gy0 + rbeg - p.e0) * p.ow + gx;
                const unsigned ow = p.ow;
                hpass<WC, CFMA, SC>(tc, wc, nrows, [&](int r, float v) { out[(size_t)r * ow] = v; });
            }
            __syncthreads();
        }
    }
}

/* ====================================================================================================================
 * Front end of the PRODUCT route (x2-type magnification: chroma / alpha handed over at source resolution), second
 * generation.  Same data flow and the same arithmetic as prep_kernel above, restructured around instruction count -- the
 * kernel is issue-bound, not HBM-bound (profiles/r1_ncu_summary.md: 21 M warp instructions for 40 MB):
 *   - every stage works on PAIRS of adjacent columns: 16-byte shared-memory accesses, half the address arithmetic, one
 *     32-bit store of two FP16 conv-input samples;
 *   - stages are mapped on the flattened (row, column-pair) index space, so nearly every lane has work (the first generation
 *     walks 96-column chunks over a 68-column footprint);
 *   - source pixels are fetched as 16-bit / 64-bit words instead of single bytes;
 *   - the footprint and the table slices are found with warp reductions, no shared-memory atomics.
 * FMA = true (tensor-core conv mode): one double fused multiply-add per tap.  The reference rounds the product and the sum
 * separately (src/frawscale.cpp:313-328, 367-381); the fused form differs by < 2^-52 relative BEFORE the rounding to FP32,
 * i.e. the FP32 plane is identical except where the double sum falls within ~1e-16 of an FP32 rounding boundary.  The
 * north-star gate for this plane is 1e-5 relative; tests/test_parity_gpu.py checks it on this kernel's own FP32 output.
 * FMA = false (bit-exact conv mode, plane getters): multiply and add rounded separately, bit-identical to the reference.
 * Chroma / alpha never differ: their colour conversion is the reference's unfused FP32 expression in both modes.
 * ==================================================================================================================== */
/* byte k of `v` as a float without the slow I2F unit: 0x4B0000bb is 2^23 + bb, and the subtraction is exact */
template <int K>
__device__ __forceinline__ float byte_f32(uint32_t v)
{
    return __fsub_rn(__uint_as_float(__byte_perm(v, 0x4B000000u, 0x7440u | K)), 8388608.f);
}

/* two adjacent pixels: the raw words (requested early) and their channels as floats */
/* resident CTAs per SM the compiler plans for (register budget 65536 / (256 x this)) */
#ifndef SRCNN_PREP2_STOP
#define SRCNN_PREP2_STOP 0      /* development: 1 / 2 = leave the kernel after colour-in / after the vertical pass (phase timing) */
#endif
#ifndef SRCNN_PREP2_MINB
#define SRCNN_PREP2_MINB 5      /* measured: 4 -> 5 CTAs per SM (48 registers, 16-44 bytes spilled): C2 42.1 -> 41.8 us, C3 468 -> 449 us; 6: worse */
#endif
template <int D>
__device__ __forceinline__ void load_raw2(const uint8_t* __restrict__ p, uint32_t (&r)[3])
{
    if (D == 4) {
        const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));            /* 8-byte aligned: (row * w + col) even */
        r[0] = v.x; r[1] = v.y; r[2] = 0;
    } else {
        const unsigned short* q = reinterpret_cast<const unsigned short*>(p); /* 2-byte aligned: 3 * (row * w + col), even */
        r[0] = __ldg(q); r[1] = __ldg(q + 1); r[2] = __ldg(q + 2);
    }
}
template <int D>
__device__ __forceinline__ void unpack_px2(const uint32_t (&r)[3], float (&c)[2][D])
{
    if (D == 4) {
        c[0][0] = byte_f32<0>(r[0]); c[0][1] = byte_f32<1>(r[0]); c[0][2] = byte_f32<2>(r[0]); c[0][D - 1] = byte_f32<3>(r[0]);
        c[1][0] = byte_f32<0>(r[1]); c[1][1] = byte_f32<1>(r[1]); c[1][2] = byte_f32<2>(r[1]); c[1][D - 1] = byte_f32<3>(r[1]);
    } else {
        c[0][0] = byte_f32<0>(r[0]); c[0][1] = byte_f32<1>(r[0]); c[0][2] = byte_f32<0>(r[1]);
        c[1][0] = byte_f32<1>(r[1]); c[1][1] = byte_f32<0>(r[2]); c[1][2] = byte_f32<1>(r[2]);
    }
}
/* (double)f for a finite, non-negative, normal-or-zero float -- every Y sample is one -- with integer instructions instead of
 * the conversion unit: sign 0, exponent re-biased by 1023 - 127 = 896, mantissa shifted into place */
__device__ __forceinline__ double widen(float f)
{
    const uint32_t b = __float_as_uint(f);
    return __hiloint2double((int)(b ? (b >> 3) + 0x38000000u : 0u), (int)(b << 29));
}

/* SRCNN_PREP2_F32SM = 1: both shared-memory planes hold FP32 (the values ARE FP32: Y from the colour conversion, the reference's
 * FP32 intermediate plane) and are widened in registers, and the vertical pass takes two ADJACENT destination rows per step, which
 * share all but one of their source rows (WY + 1 loads instead of 2 WY; the second row meets a zero weight on the sample it does
 * not use).  Same arithmetic, bit for bit; ~45 % fewer LSU cycles per tile (the LSU data pipe is this kernel's busiest unit). */
#ifndef SRCNN_PREP2_F32SM
#define SRCNN_PREP2_F32SM 0      /* measured (round 2): C2 39.7 vs 36.4 us, C3 432 vs 362 us -- fewer LSU cycles, more ALU + a longer dependency chain: slower */
#endif
#if SRCNN_PREP2_F32SM
typedef float sm_t;
#else
typedef double sm_t;
#endif
/* (double)f for any finite float, sign and subnormals included */
__device__ __forceinline__ double widen_s(float f)
{
    const uint32_t b = __float_as_uint(f), m = b & 0x7fffffffu;
    if (m != 0u && m < 0x00800000u) return (double)f;                     /* subnormal: the conversion unit (never seen in practice) */
    return __hiloint2double((int)((m ? (m >> 3) + 0x38000000u : 0u) | (b & 0x80000000u)), (int)(b << 29));
}

template <int D, int WY, int SC, bool FMA>
__global__ void __launch_bounds__(NT, SRCNN_PREP2_MINB)
prep2_kernel(const PrepParams p)
{
    pdl_launch_dependents();                          /* the conv kernel may set itself up while this grid drains */
    const uint8_t* const f_src = p.src + (size_t)blockIdx.z * p.src_fstride;
    __half* const f_y16 = p.y16 ? p.y16 + (size_t)blockIdx.z * p.y16_fstride : nullptr;
    float* const f_csrc = p.csrc + (size_t)blockIdx.z * p.csrc_fstride;
    constexpr int NC = D - 1;
    constexpr int NCP = SC / 2;                       /* column pairs per shared-memory row */
    constexpr int RL = NT / NCP;                      /* row lanes of the vertical pass */
    extern __shared__ double sm[];
    sm_t* const S = reinterpret_cast<sm_t*>(sm);      /* [SR][SC]  Y of the source footprint */
    sm_t* const T = S + p.SR * SC;                    /* [TH][SC]  Y after the vertical pass (FP32 values) */
    __shared__ int s_red[8][4];
    __shared__ int s_vl[TH];
    __shared__ double s_vw[TH * WY];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned x0 = blockIdx.x * TW, gy0 = p.e0 + blockIdx.y * TH;
    const int ny = (int)min((unsigned)TH, p.e1 - gy0), nx = (int)min((unsigned)TW, p.ow - x0);

    /* the source bytes of the tile come straight from HBM, and the exact footprint is only known once the tables have been read:
     * ask L2 for the approximate footprint now (same magnification on both axes: source = destination * w / ow), so that the
     * colour-in loads below find their lines in L2 instead of paying a DRAM round trip behind two block barriers */
    if (tid < 96) {
        const unsigned scale_num = p.w, scale_den = p.ow;
        const int r_first = (int)((size_t)gy0 * scale_num / scale_den) - 3, r_last = (int)((size_t)(gy0 + ny) * scale_num / scale_den) + 3;
        const int c_first = max(0, (int)((size_t)x0 * scale_num / scale_den) - 4);
        const int row = r_first + tid / 3, k = tid % 3;
        const int rel = row - (int)p.src_row0;
        if (row <= r_last && rel >= 0 && rel < (int)p.src_rows) {
            const size_t off = ((size_t)rel * p.w + (size_t)c_first) * D + (size_t)k * 128;
            if (off < (size_t)(rel + 1) * p.w * D)                  /* stay inside the row (and with it inside the image) */
                asm volatile("prefetch.global.L2 [%0];" :: "l"(f_src + off));
        }
    }

    /* ---- 0. footprint of the tile: rows [rlo, rlo+nr), columns [clo, clo+nc) with clo, nc even; chroma region inside it */
    {
        int lo = INT_MAX, hi = -1, clo_ = INT_MAX, chi_ = -1;
        if (warp < 4) {                               /* columns: the horizontal tables of this tile's nx destination columns */
            if (tid < nx) {
                const int a = __ldg(p.hyl + x0 + tid), b = __ldg(p.hcl + x0 + tid);
                lo = min(a, b); hi = max(a + WY, b + 2); clo_ = b; chi_ = b + 2;
            }
        } else if (warp == 4) {                       /* rows */
            if (lane < ny) {
                const int a = __ldg(p.vyl + gy0 + lane), b = __ldg(p.vcl + gy0 + lane);
                lo = min(a, b); hi = max(a + WY, b + 2); clo_ = b; chi_ = b + 2;
                s_vl[lane] = a;
            }
        }
        if (warp < 5) {
            lo = __reduce_min_sync(0xffffffffu, lo); hi = __reduce_max_sync(0xffffffffu, hi);
            clo_ = __reduce_min_sync(0xffffffffu, clo_); chi_ = __reduce_max_sync(0xffffffffu, chi_);
            if (lane == 0) { s_red[warp][0] = lo; s_red[warp][1] = hi; s_red[warp][2] = clo_; s_red[warp][3] = chi_; }
        }
        for (int i = tid; i < ny * WY; i += NT) s_vw[i] = __ldg(p.vyw + (size_t)gy0 * WY + i);
    }
    __syncthreads();
    int clo = min(min(s_red[0][0], s_red[1][0]), min(s_red[2][0], s_red[3][0]));
    int chi = max(max(s_red[0][1], s_red[1][1]), max(s_red[2][1], s_red[3][1]));
    int cc0 = min(min(s_red[0][2], s_red[1][2]), min(s_red[2][2], s_red[3][2]));      /* chroma columns handed over by this tile */
    int cc1 = max(max(s_red[0][3], s_red[1][3]), max(s_red[2][3], s_red[3][3]));
    const int rlo = s_red[4][0], nr = s_red[4][1] - rlo, cr0 = s_red[4][2], cr1 = s_red[4][3];
    clo &= ~1; chi = (chi + 1) & ~1; cc0 &= ~1; cc1 = (cc1 + 1) & ~1;
    const int nc = chi - clo, ncp = nc >> 1;

    /* ---- 1. colour-in of the footprint (reference converImgU8toYCbCr, src/libsrcnn.cpp:233-272): a thread keeps ONE column
     * pair and walks down the rows in steps of RL, so everything but the row is loop-invariant.  All of a thread's source
     * words (at most NI rows) are requested before the first one is used: the stage is bound by load latency otherwise */
    {
        constexpr int NI = (28 + RL - 1) / RL;                    /* rows per thread: nr <= 28 <= NI * RL (checked on the host) */
        const int cp = tid % NCP, rl = tid / NCP;
        if (rl < RL && cp < ncp) {
            const int col = clo + 2 * cp;
            const bool ccol = col >= cc0 && col < cc1;
            const uint32_t pix0 = (uint32_t)(rlo + rl - (int)p.src_row0) * p.w + (uint32_t)col;    /* (fits: checked on the host) */
            const uint32_t pstep = (uint32_t)RL * p.w;
            uint32_t raw[NI][3];
            #pragma unroll
            for (int k = 0; k < NI; k++)
                if (rl + k * RL < nr) load_raw2<D>(f_src + (size_t)(pix0 + k * pstep) * D, raw[k]);
            #pragma unroll
            for (int k = 0; k < NI; k++) {
                const int r = rl + k * RL;
                if (r < nr) {
                    float c[2][D];
                    unpack_px2<D>(raw[k], c);
                    float y0, y1, cb0, cb1, cr0_, cr1_;
                    ycbcr_of(c[0][0], c[0][1], c[0][2], y0, cb0, cr0_);
                    ycbcr_of(c[1][0], c[1][1], c[1][2], y1, cb1, cr1_);
#if SRCNN_PREP2_F32SM
                    *reinterpret_cast<float2*>(S + r * SC + 2 * cp) = make_float2(y0, y1);
#else
                    *reinterpret_cast<double2*>(S + r * SC + 2 * cp) = make_double2(widen(y0), widen(y1));
#endif
                    const int row = rlo + r;
                    if (ccol && row >= cr0 && row < cr1) {
                        /* neighbouring tiles' regions overlap by a row / column pair: they store identical values (benign) */
                        float* q = f_csrc + (pix0 + k * pstep);
                        *reinterpret_cast<float2*>(q) = make_float2(cb0, cb1);
                        *reinterpret_cast<float2*>(q + p.csrc_stride) = make_float2(cr0_, cr1_);
                        if (NC == 3) *reinterpret_cast<float2*>(q + 2 * p.csrc_stride) = make_float2(c[0][D - 1], c[1][D - 1]);
                    }
                }
            }
        }
    }
    /* this thread's column pair of the horizontal table (global loads overlap the barrier) */
    const int j = tid & (TW / 2 - 1), rbeg = (tid / (TW / 2)) * (TH / (NT / (TW / 2)));
    constexpr int RPT2 = TH / (NT / (TW / 2));        /* rows per thread in the horizontal pass */
    const unsigned gx = x0 + 2 * j;
    const bool col_ok = 2 * j < nx;
    double wa[WY], wb[WY + 1];
    int o[WY + 1];
    if (col_ok) {
        const int la = __ldg(p.hyl + gx), lb = __ldg(p.hyl + gx + 1);     /* lb - la is 0 or 1 (checked on the host) */
        /* one 16-byte load per tap brings the weights of both columns: the table is stored tap-major for this (a warp then reads
         * 512 contiguous bytes per tap; with the destination-major table every lane's 64-byte-strided loads cost a wavefront
         * each -- a third of all LSU wavefronts of the kernel, and the LSU data pipe is its busiest unit, ncu round 2) */
        double t[WY];
        #pragma unroll
        for (int i = 0; i < WY; i++) {
            const double2 w2 = __ldg(reinterpret_cast<const double2*>(p.hywt + (size_t)i * p.ow + gx));
            wa[i] = w2.x; t[i] = w2.y;
        }
        /* the second column reads the same WY + 1 samples with its weights shifted by lb - la (a zero weight on the sample it
         * does not use: fma(0, v, acc) == acc for the finite v of an initialised footprint) */
        const bool sh = lb != la;
        wb[0] = sh ? 0.0 : t[0];
        #pragma unroll
        for (int i = 1; i < WY; i++) wb[i] = sh ? t[i - 1] : t[i];
        wb[WY] = sh ? t[WY - 1] : 0.0;
        #pragma unroll
        for (int i = 0; i <= WY; i++) o[i] = min(la - clo + i, nc - 1);
    }
    __syncthreads();

#if SRCNN_PREP2_STOP == 1
    return;                                           /* (timing experiment: colour-in only) */
#endif
    /* ---- 2. vertical pass, shared -> shared (reference verticalFilter, src/frawscale.cpp:335-385): thread = column pair x row lane */
#if SRCNN_PREP2_F32SM
    {
        const int cp = tid % NCP, rl = tid / NCP;
        if (rl < RL && cp < ncp) {
            const float2* const S2 = reinterpret_cast<const float2*>(S) + cp;
            float2* const T2 = reinterpret_cast<float2*>(T) + cp;
            /* destination rows (2q, 2q+1) of the tile: their windows start at the same source row or one apart (checked on the
             * host), so both read the WY + 1 rows from the first window's start; four independent accumulator chains */
            #pragma unroll 1
            for (int q = rl; 2 * q < ny; q += RL) {
                const int ya = 2 * q, yb = min(ya + 1, ny - 1);
                const int base = s_vl[ya] - rlo;
                const bool sh = s_vl[yb] != s_vl[ya];
                double vx[WY + 1], vy_[WY + 1];
                #pragma unroll
                for (int i = 0; i <= WY; i++) {
                    const float2 v = S2[min(base + i, nr - 1) * NCP];          /* (the clamped row only ever meets a zero weight) */
                    vx[i] = widen(v.x); vy_[i] = widen(v.y);
                }
                double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
                #pragma unroll
                for (int i = 0; i < WY; i++) {
                    const double w = s_vw[ya * WY + i];
                    a0 = mac<FMA>(a0, w, vx[i]); a1 = mac<FMA>(a1, w, vy_[i]);
                }
                /* second row: its weights shifted by the window offset; 0 * v adds +-0 to a sum that starts at +0 (a no-op) */
                #pragma unroll
                for (int i = 0; i <= WY; i++) {
                    const int k = sh ? i - 1 : i;
                    const double w = (k >= 0 && k < WY) ? s_vw[yb * WY + k] : 0.0;
                    b0 = mac<FMA>(b0, w, vx[i]); b1 = mac<FMA>(b1, w, vy_[i]);
                }
                T2[ya * NCP] = make_float2(__double2float_rn(a0), __double2float_rn(a1));      /* the reference's FP32 intermediate plane */
                T2[yb * NCP] = make_float2(__double2float_rn(b0), __double2float_rn(b1));
            }
        }
    }
#else
    {
        const int cp = tid % NCP, rl = tid / NCP;
        if (rl < RL && cp < ncp) {
            const double2* const S2 = reinterpret_cast<const double2*>(S) + cp;
            double2* const T2 = reinterpret_cast<double2*>(T) + cp;
            /* two destination rows per step: four independent accumulator chains hide the FP64 / shared-memory latency */
            #pragma unroll 1
            for (int y = rl; y < ny; y += 2 * RL) {
                const int y2 = min(y + RL, ny - 1);                /* (the last step may repeat row ny-1: same value stored twice) */
                const double2* sa = S2 + (s_vl[y] - rlo) * NCP;
                const double2* sb = S2 + (s_vl[y2] - rlo) * NCP;
                double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
                #pragma unroll
                for (int i = 0; i < WY; i++) {
                    const double wa_ = s_vw[y * WY + i], wb_ = s_vw[y2 * WY + i];
                    const double2 va = sa[i * NCP], vb = sb[i * NCP];
                    a0 = mac<FMA>(a0, wa_, va.x); a1 = mac<FMA>(a1, wa_, va.y);
                    b0 = mac<FMA>(b0, wb_, vb.x); b1 = mac<FMA>(b1, wb_, vb.y);
                }
                /* the reference's FP32 intermediate plane */
                T2[y * NCP] = make_double2((double)__double2float_rn(a0), (double)__double2float_rn(a1));
                T2[y2 * NCP] = make_double2((double)__double2float_rn(b0), (double)__double2float_rn(b1));
            }
        }
    }
#endif
    __syncthreads();

#if SRCNN_PREP2_STOP == 2
    return;                                           /* (timing experiment: colour-in + vertical pass) */
#endif
    /* ---- 3. horizontal pass, shared -> registers -> global (reference horizontalFilter, src/frawscale.cpp:288-332) */
    const int ybase = (int)gy0 - p.strip_row0 + kHaloRows;       /* conv-input plane row of tile row 0 */
    if (col_ok) {
        const int nrows = max(0, min(ny - rbeg, RPT2));
        const sm_t* t = T + rbeg * SC;
        __half2* o16 = f_y16 ? reinterpret_cast<__half2*>(f_y16 + (size_t)(ybase + rbeg) * p.pitch + gx + kY16PadL) : nullptr;
        float2* o32 = p.write_y32 ? reinterpret_cast<float2*>(p.planes + (size_t)(gy0 + rbeg - p.e0) * p.ow + gx) : nullptr;
        const int pitch2 = p.pitch >> 1; const unsigned ow2 = p.ow >> 1;
        auto body = [&](int r) {
            double v[WY + 1];
            #pragma unroll
#if SRCNN_PREP2_F32SM
            for (int i = 0; i <= WY; i++) v[i] = widen_s(t[r * SC + o[i]]);
#else
            for (int i = 0; i <= WY; i++) v[i] = t[r * SC + o[i]];
#endif
            double a = 0.0, b = 0.0;
            #pragma unroll
            for (int i = 0; i < WY; i++) a = mac<FMA>(a, wa[i], v[i]);
            #pragma unroll
            for (int i = 0; i <= WY; i++) b = mac<FMA>(b, wb[i], v[i]);     /* (0 * v adds +-0: a no-op on a sum that starts at +0) */
            const float fa = __double2float_rn(a), fb = __double2float_rn(b);
            if (o16) o16[r * pitch2] = __floats2half2_rn(fa, fb);
            if (o32) o32[(size_t)r * ow2] = make_float2(fa, fb);
        };
        if (nrows == RPT2) {
            #pragma unroll
            for (int r = 0; r < RPT2; r++) body(r);
        } else {
            #pragma unroll 1
            for (int r = 0; r < nrows; r++) body(r);
        }
    }
    if (!f_y16) return;
    __syncthreads();            /* the tile's conv-input values are visible to the whole CTA */

    /* ---- 3b. replicate padding of the conv-input plane around the image (border tiles only), as in prep_kernel */
    const bool first_col = (x0 == 0), last_col = (x0 + nx == p.ow);
    if (first_col || last_col) {
        const int c_hi = (int)p.ow + kY16PadL, npad_r = last_col ? p.pitch - c_hi : 0, npad_l = first_col ? kY16PadL : 0;
        const int per_row = npad_l + npad_r;
        for (int idx = tid; idx < ny * per_row; idx += NT) {
            const int r = idx / per_row, k = idx - r * per_row;
            __half* row = f_y16 + (size_t)(ybase + r) * p.pitch;
            if (k < npad_l) row[k] = row[kY16PadL];
            else            row[c_hi + (k - npad_l)] = row[c_hi - 1];
        }
        __syncthreads();
    }
    const bool top = (gy0 == 0) && ybase > 0, bottom = (gy0 + ny == p.oh) && (ybase + ny < p.y16_rows);
    if (top || bottom) {
        const int ca = first_col ? 0 : (int)x0 + kY16PadL, cb = last_col ? p.pitch : (int)x0 + nx + kY16PadL;
        const int span = cb - ca;
        if (top) {
            const __half* from = f_y16 + (size_t)ybase * p.pitch + ca;
            for (int idx = tid; idx < ybase * span; idx += NT) {
                const int r = idx / span, c = idx - r * span;
                f_y16[(size_t)r * p.pitch + ca + c] = from[c];
            }
        }
        if (bottom) {
            const int rb = ybase + ny, nrep = p.y16_rows - rb;
            const __half* from = f_y16 + (size_t)(rb - 1) * p.pitch + ca;
            for (int idx = tid; idx < nrep * span; idx += NT) {
                const int r = idx / span, c = idx - r * span;
                f_y16[(size_t)(rb + r) * p.pitch + ca + c] = from[c];
            }
        }
    }
}

/* dynamic shared memory of one CTA (bytes) */
inline size_t prep_smem(int d, int cm, int SR, int SC, bool csrc)
{
    if (cm == 2 && csrc) return ((size_t)SR + TH) * SC * sizeof(double);
    if (cm == 2) return ((size_t)SR + (d == 4 ? 2 : 1) * TH) * SC * sizeof(double) + (size_t)SR * SC * (d - 1) * sizeof(float);
    return ((size_t)d * SR + TH) * SC * sizeof(double);
}

/* Function attributes are per device: every prep_kernel instantiation keeps one "opted in to large dynamic shared memory"
 * flag per CUDA device (contexts on several GPUs of one process, srcnn_b200_multi). */
constexpr int kMaxDevices = 64;

template <int D, int WY, int WC, int CM, int SC>
int launch_one(const PrepParams& p, dim3 grid, cudaStream_t s)
{
    static std::atomic<bool> configured[kMaxDevices];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) { cudaGetLastError(); return 0; }
    if (!configured[dev].load(std::memory_order_acquire)) {
        if (cudaFuncSetAttribute(prep_kernel<D, WY, WC, CM, SC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
            cudaGetLastError(); return 0;
        }
        configured[dev].store(true, std::memory_order_release);
    }
    prep_kernel<D, WY, WC, CM, SC><<<grid, NT, prep_smem(D, CM, p.SR, SC, p.csrc != nullptr), s>>>(p);
    return cudaPeekAtLastError() == cudaSuccess ? 1 : -1;      /* -1: the launch itself failed (sticky error left for the caller) */
}

/* shared-memory row pitch (doubles): 72 covers magnifications >= 2 (128/2 + window), 104 those >= 1.5 */
template <int D, int WY, int CM>
int launch_sc(const PrepParams& p, dim3 grid, cudaStream_t s)
{
    if (p.SC <= 72)  return launch_one<D, WY, 2, CM, 72>(p, grid, s);
    if (p.SC <= 104) return launch_one<D, WY, 2, CM, 104>(p, grid, s);
    return 0;                                        /* shallower magnification: the separate kernels */
}

template <int D, int WY>
int launch_cm(const PrepParams& p, int cm, dim3 grid, cudaStream_t s)
{
    switch (cm) {
        case 2:  return launch_sc<D, WY, 2>(p, grid, s);
        case 1:  return launch_sc<D, WY, 1>(p, grid, s);
        default: return launch_sc<D, WY, 0>(p, grid, s);
    }
}

template <int D>
int launch_d(const PrepParams& p, int wy, int cm, dim3 grid, cudaStream_t s)
{
    switch (wy) {
        case 2:  return launch_cm<D, 2>(p, cm, grid, s);
        case 4:  return launch_cm<D, 4>(p, cm, grid, s);
        default: return launch_cm<D, 6>(p, cm, grid, s);
    }
}

/* largest source footprint (extent along one axis) of any tile of `tile` destination samples starting at d0 */
int max_footprint(const int* ly, int wy, const int* lc, int wc, unsigned d0, unsigned d1, unsigned tile)
{
    int best = 0;
    for (unsigned a = d0; a < d1; a += tile) {
        const unsigned b = std::min(d1, a + tile);
        int lo = INT_MAX, hi = -1;
        for (unsigned u = a; u < b; u++) {
            lo = std::min(lo, std::min(ly[u], lc[u]));
            hi = std::max(hi, std::max(ly[u] + wy, lc[u] + wc));
        }
        best = std::max(best, hi - lo);
    }
    return best;
}



/* host-side: largest even-aligned column extent / row extent of a tile footprint */
int max_footprint_even(const int* ly, int wy, const int* lc, int wc, unsigned d0, unsigned d1, unsigned tile)
{
    int best = 0;
    for (unsigned a = d0; a < d1; a += tile) {
        const unsigned b = std::min(d1, a + tile);
        int lo = INT_MAX, hi = -1;
        for (unsigned u = a; u < b; u++) {
            lo = std::min(lo, std::min(ly[u], lc[u]));
            hi = std::max(hi, std::max(ly[u] + wy, lc[u] + wc));
        }
        best = std::max(best, ((hi + 1) & ~1) - (lo & ~1));
    }
    return best;
}

template <int D, int WY, bool FMA, int SC>
int launch_prep2_sc(const PrepParams& p, dim3 grid, cudaStream_t s)
{
    static std::atomic<bool> configured[kMaxDevices];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) { cudaGetLastError(); return 0; }
    const size_t smem = ((size_t)p.SR + TH) * SC * sizeof(sm_t);
    if (!configured[dev].load(std::memory_order_acquire)) {
        if (cudaFuncSetAttribute(prep2_kernel<D, WY, SC, FMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess) {
            cudaGetLastError(); return 0;
        }
        configured[dev].store(true, std::memory_order_release);
    }
    if (smem > 100 * 1024) return 0;
    prep2_kernel<D, WY, SC, FMA><<<grid, NT, smem, s>>>(p);
    return cudaPeekAtLastError() == cudaSuccess ? 1 : -1;
}

/* shared-memory row pitch (doubles): 72 covers magnifications >= 2, 104 those >= 1.5 */
template <int D, int WY, bool FMA>
int launch_prep2(const PrepParams& p, dim3 grid, cudaStream_t s)
{
    return p.SC <= 72 ? launch_prep2_sc<D, WY, FMA, 72>(p, grid, s) : launch_prep2_sc<D, WY, FMA, 104>(p, grid, s);
}

template <int D, bool FMA>
int launch_prep2_wy(const PrepParams& p, int wy, dim3 grid, cudaStream_t s)
{
    switch (wy) {
        case 2:  return launch_prep2<D, 2, FMA>(p, grid, s);
        case 4:  return launch_prep2<D, 4, FMA>(p, grid, s);
        case 6:  return launch_prep2<D, 6, FMA>(p, grid, s);
        default: return 0;
    }
}

} // namespace

bool prep_chroma_handover(const AxisTableDev& vc, const AxisTableDev& hc, bool approx_ok)
{
    if (getenv_flag("SRCNN_B200_PREP_F64")) return false;
    if (vc.pow2pair && hc.pow2pair) return true;
    return approx_ok && vc.pack2 && hc.pack2 && !getenv_flag("SRCNN_B200_CHROMA_EXACT");
}

bool prep_supported(const AxisTableDev& vy, const AxisTableDev& vc, const AxisTableDev& hy, const AxisTableDev& hc)
{
    if (!vy.cleft || !vc.cleft || !hy.cleft || !hc.cleft) return false;
    if (vy.cwin != hy.cwin || (vy.cwin != 2 && vy.cwin != 4 && vy.cwin != 6)) return false;
    return vc.cwin == 2 && hc.cwin == 2;
}

/* host-side: every pair of adjacent destination columns (2j, 2j+1) starts its window at the same source column or one later */
bool prep2_pairs_ok(const int* h_hyl, unsigned ow)
{
    if (ow & 1u) return false;
    for (unsigned x = 0; x + 1 < ow; x += 2) {
        const int dlt = h_hyl[x + 1] - h_hyl[x];
        if (dlt != 0 && dlt != 1) return false;
    }
    return true;
}

int launch_prep2(const uint8_t* src, unsigned w, unsigned d, unsigned src_row0, unsigned src_rows, unsigned ow, unsigned oh,
                 unsigned e0, unsigned e1, float* y32, bool fused_mac, __half* y16, Y16Geom g, unsigned strip_row0,
                 const AxisTableDev& vy, const AxisTableDev& vc, const AxisTableDev& hy, const AxisTableDev& hc,
                 const int* h_vyl, const int* h_vcl, const int* h_hyl, const int* h_hcl, bool pairs_ok, cudaStream_t s,
                 float* csrc, size_t csrc_stride, unsigned nframes, size_t src_fstride, size_t y16_fstride, size_t csrc_fstride)
{
    if (!csrc || !pairs_ok || (w & 1u) || (ow & 1u) || (csrc_stride & 1u) || (csrc_fstride & 1u) || (d != 3 && d != 4)) return 0;
    if (((uintptr_t)src & 7u) || ((src_fstride) & 7u) || ((uintptr_t)csrc & 7u)) return 0;
    if (!prep_supported(vy, vc, hy, hc) || !prep_chroma_handover(vc, hc, fused_mac)) return 0;
    if (nframes > 1 && y32) return 0;
    if (y16 && ((g.pitch & 1) || ((uintptr_t)y16 & 3u) || (y16_fstride & 1u))) return 0;
    if (y32 && (((uintptr_t)y32 & 7u))) return 0;
    PrepParams p{};
    p.src = src; p.w = w; p.src_row0 = src_row0; p.src_rows = src_rows; p.ow = ow; p.oh = oh; p.e0 = e0; p.e1 = e1;
    p.planes = y32; p.plane_stride = 0; p.write_y32 = y32 ? 1 : 0;
    p.y16 = y16; p.pitch = g.pitch; p.y16_rows = g.rows; p.strip_row0 = (int)strip_row0;
    p.vyl = vy.cleft; p.vcl = vc.cleft; p.hyl = hy.cleft; p.hcl = hc.cleft;
    p.vyw = vy.cw; p.vcw = vc.cw; p.hyw = hy.cw; p.hcw = hc.cw;
    p.hywt = hy.cwt; p.hcwt = hc.cwt;
    if (!p.hywt) return 0;
    p.SR = max_footprint(h_vyl, vy.cwin, h_vcl, vc.cwin, e0, e1, TH);
    p.SC = max_footprint_even(h_hyl, hy.cwin, h_hcl, hc.cwin, 0, ow, TW);
    if (p.SC > 104 || p.SR > 28) return 0;                      /* prep2_kernel: SC = 72 / 104, at most 28 footprint rows */
#if SRCNN_PREP2_F32SM
    for (unsigned y = e0; y + 1 < e1; y++)                      /* the vertical pass pairs adjacent rows: windows 0 or 1 apart */
        if (h_vyl[y + 1] - h_vyl[y] != 0 && h_vyl[y + 1] - h_vyl[y] != 1) return 0;
#endif
    if ((size_t)(oh ? (size_t)w * ((size_t)oh / 1 + 64) : 0) * 4 >= ((size_t)1 << 31)) return 0;       /* 32-bit pixel offsets (source rows <= oh) */
    p.csrc = csrc; p.csrc_stride = csrc_stride;
    p.src_fstride = src_fstride; p.y16_fstride = y16_fstride; p.csrc_fstride = csrc_fstride;
    if ((size_t)oh * 8 >= (size_t)INT_MAX || (size_t)ow * 8 >= (size_t)INT_MAX) return 0;
    dim3 grid((ow + TW - 1) / TW, (e1 - e0 + TH - 1) / TH, nframes ? nframes : 1);
    if (grid.y > 65535u || grid.z > 65535u) return 0;
    if (d == 3) return fused_mac ? launch_prep2_wy<3, true>(p, vy.cwin, grid, s) : launch_prep2_wy<3, false>(p, vy.cwin, grid, s);
    return fused_mac ? launch_prep2_wy<4, true>(p, vy.cwin, grid, s) : launch_prep2_wy<4, false>(p, vy.cwin, grid, s);
}

int launch_prep(const uint8_t* src, unsigned w, unsigned d, unsigned src_row0, unsigned ow, unsigned oh,
                unsigned e0, unsigned e1, float* planes, size_t plane_stride, bool write_y32,
                __half* y16, Y16Geom g, unsigned strip_row0,
                const AxisTableDev& vy, const AxisTableDev& vc, const AxisTableDev& hy, const AxisTableDev& hc,
                const int* h_vyl, const int* h_vcl, const int* h_hyl, const int* h_hcl, cudaStream_t s,
                float* csrc, size_t csrc_stride, unsigned nframes, size_t src_fstride, size_t y16_fstride, size_t csrc_fstride)
{
    if (nframes > 1 && (planes || write_y32 || !csrc)) return 0;     /* groups: the x2 hand-over route only */
    if (!prep_supported(vy, vc, hy, hc) || (d != 3 && d != 4)) return 0;
    PrepParams p{};
    p.src = src; p.w = w; p.src_row0 = src_row0; p.ow = ow; p.oh = oh; p.e0 = e0; p.e1 = e1;
    p.planes = planes; p.plane_stride = plane_stride; p.write_y32 = write_y32 ? 1 : 0;
    p.y16 = y16; p.pitch = g.pitch; p.y16_rows = g.rows; p.strip_row0 = (int)strip_row0;
    p.vyl = vy.cleft; p.vcl = vc.cleft; p.hyl = hy.cleft; p.hcl = hc.cleft;
    p.vyw = vy.cw; p.vcw = vc.cw; p.hyw = hy.cw; p.hcw = hc.cw;
    p.hywt = hy.cwt; p.hcwt = hc.cwt;
    p.SR = max_footprint(h_vyl, vy.cwin, h_vcl, vc.cwin, e0, e1, TH);
    p.SC = max_footprint(h_hyl, hy.cwin, h_hcl, hc.cwin, 0, ow, TW);
    /* chroma arithmetic: 2 = FP32 two-tap (pow2pair tables), 1 = FP64 fused multiply-add (exact29), 0 = FP64 mul + add */
    const int cm = (vc.pow2pair && hc.pow2pair && !getenv_flag("SRCNN_B200_PREP_F64")) ? 2 : (vc.exact29 && hc.exact29) ? 1 : 0;
    if (cm != 2) csrc = nullptr;
    if (nframes > 1 && !csrc) return 0;
    p.csrc = csrc; p.csrc_stride = csrc_stride;
    p.src_fstride = src_fstride; p.y16_fstride = y16_fstride; p.csrc_fstride = csrc_fstride;
    if (prep_smem((int)d, cm, p.SR, 104, csrc != nullptr) > 160 * 1024) return 0;
    if ((size_t)oh * 8 >= (size_t)INT_MAX || (size_t)ow * 8 >= (size_t)INT_MAX) return 0;
    dim3 grid((ow + TW - 1) / TW, (e1 - e0 + TH - 1) / TH, nframes ? nframes : 1);
    if (grid.y > 65535u || grid.z > 65535u) return 0;
    return d == 3 ? launch_d<3>(p, vy.cwin, cm, grid, s) : launch_d<4>(p, vy.cwin, cm, grid, s);
}

} // namespace srcnn
