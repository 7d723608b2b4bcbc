/* Colour-space kernels (HBM-bound, vectorised: 4 pixels per thread).
 *
 *  color_in  = reference converImgU8toYCbCr   (src/libsrcnn.cpp:233-272)
 *  color_out = reference convertImgF32XtoImgU8 (src/libsrcnn.cpp:274-308) + conv-output plane (:889-915)
 *  color_out_up = the same with the x2 chroma / alpha resampling (FRAWResizeEngine::scale, bilinear / box) folded in
 *
 * Every FP32 operation uses the round-to-nearest non-fused intrinsics in the reference's association order
 * (its x86-64 -O2 build has no FMA contraction), so both kernels are bit-exact against the oracle.
 */
#include "kernels.h"
#include <cstdlib>
#include "color_math.cuh"

namespace srcnn {

__device__ __forceinline__ void ycbcr_of(float r, float g, float b, float& y, float& cb, float& cr)
{
    /* Y = .299R + .587G + .114B ; Cb = 128 - .1687R - .3313G + .5B ; Cr = 128 + .5R - .4187G - .0813B */
    y  = __fadd_rn(__fadd_rn(__fmul_rn(0.299f, r), __fmul_rn(0.587f, g)), __fmul_rn(0.114f, b));
    cb = __fadd_rn(__fsub_rn(__fsub_rn(128.f, __fmul_rn(0.1687f, r)), __fmul_rn(0.3313f, g)), __fmul_rn(0.5f, b));
    cr = __fsub_rn(__fsub_rn(__fadd_rn(128.f, __fmul_rn(0.5f, r)), __fmul_rn(0.4187f, g)), __fmul_rn(0.0813f, b));
}

/* 4 pixels per thread: 12 B (RGB) or 16 B (RGBA) in, one float4 per plane out */
template <int D, bool VEC>
__global__ void __launch_bounds__(256)
color_in_kernel(const uint8_t* __restrict__ src, size_t npix, float* __restrict__ planes, size_t plane_stride)
{
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      /* quad index */
    const size_t i0 = q * 4;
    if (i0 >= npix) return;
    float Y[4], CB[4], CR[4], A[4];
    if (VEC && i0 + 4 <= npix) {
        uint32_t wds[D];
        if (D == 4) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(src) + q);
            wds[0] = v.x; wds[1] = v.y; wds[2] = v.z; wds[D - 1] = v.w;
        } else {
            const uint32_t* p = reinterpret_cast<const uint32_t*>(src) + q * 3;
            wds[0] = __ldg(p); wds[1] = __ldg(p + 1); wds[2] = __ldg(p + 2);
        }
        #pragma unroll
        for (int k = 0; k < 4; k++) {
            auto byte_at = [&](int b) -> float { return (float)((wds[b >> 2] >> ((b & 3) * 8)) & 0xffu); };
            ycbcr_of(byte_at(k * D + 0), byte_at(k * D + 1), byte_at(k * D + 2), Y[k], CB[k], CR[k]);
            if (D == 4) A[k] = byte_at(k * D + 3);
        }
        *reinterpret_cast<float4*>(planes + i0)                    = make_float4(Y[0], Y[1], Y[2], Y[3]);
        *reinterpret_cast<float4*>(planes + plane_stride + i0)     = make_float4(CB[0], CB[1], CB[2], CB[3]);
        *reinterpret_cast<float4*>(planes + 2 * plane_stride + i0) = make_float4(CR[0], CR[1], CR[2], CR[3]);
        if (D == 4)
            *reinterpret_cast<float4*>(planes + 3 * plane_stride + i0) = make_float4(A[0], A[1], A[2], A[3]);
    } else {
        for (size_t i = i0; i < npix && i < i0 + 4; i++) {
            float y, cb, cr;
            ycbcr_of((float)src[i * D], (float)src[i * D + 1], (float)src[i * D + 2], y, cb, cr);
            planes[i] = y; planes[plane_stride + i] = cb; planes[2 * plane_stride + i] = cr;
            if (D == 4) planes[3 * plane_stride + i] = (float)src[i * D + 3];
        }
    }
}

void launch_color_in(const uint8_t* src, size_t npix, unsigned d, float* planes, size_t plane_stride, cudaStream_t s)
{
    const size_t quads = (npix + 3) / 4;
    const unsigned blocks = (unsigned)((quads + 255) / 256);
    /* the 4-pixel vector path needs 16 B (RGBA) / 4 B (RGB) aligned input and 16 B aligned planes */
    const bool vec = (((uintptr_t)src % (d == 4 ? 16 : 4)) == 0) && (((uintptr_t)planes % 16) == 0) &&
                     (plane_stride % 4 == 0);
    if (d == 4) { if (vec) color_in_kernel<4, true ><<<blocks, 256, 0, s>>>(src, npix, planes, plane_stride);
                  else     color_in_kernel<4, false><<<blocks, 256, 0, s>>>(src, npix, planes, plane_stride); }
    else        { if (vec) color_in_kernel<3, true ><<<blocks, 256, 0, s>>>(src, npix, planes, plane_stride);
                  else     color_in_kernel<3, false><<<blocks, 256, 0, s>>>(src, npix, planes, plane_stride); }
}

/* ------------------------------------------------------------------------------------------------------------ */

template <int D, bool CONV, bool VEC>
__global__ void __launch_bounds__(256)
color_out_kernel(const float* __restrict__ yp, const float* __restrict__ cb, const float* __restrict__ cr,
                 const float* __restrict__ al, size_t npix, uint8_t* __restrict__ out, uint8_t* __restrict__ conv)
{
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t i0 = q * 4;
    if (i0 >= npix) return;
    if (VEC && i0 + 4 <= npix) {
        const float4 y4 = __ldg(reinterpret_cast<const float4*>(yp + i0));
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(cb + i0));
        const float4 r4 = __ldg(reinterpret_cast<const float4*>(cr + i0));
        float4 a4 = make_float4(0, 0, 0, 0);
        if (D == 4) a4 = __ldg(reinterpret_cast<const float4*>(al + i0));
        const float Y[4] = {y4.x, y4.y, y4.z, y4.w}, B[4] = {b4.x, b4.y, b4.z, b4.w},
                    R[4] = {r4.x, r4.y, r4.z, r4.w}, A[4] = {a4.x, a4.y, a4.z, a4.w};
        uint32_t bytes[4 * D];
        #pragma unroll
        for (int k = 0; k < 4; k++) {
            rgb_of(Y[k], B[k], R[k], bytes[k * D], bytes[k * D + 1], bytes[k * D + 2]);
            if (D == 4) bytes[k * D + 3] = to_u8(A[k]);
        }
        uint32_t wds[D];
        #pragma unroll
        for (int wi = 0; wi < D; wi++)
            wds[wi] = bytes[wi * 4] | (bytes[wi * 4 + 1] << 8) | (bytes[wi * 4 + 2] << 16) | (bytes[wi * 4 + 3] << 24);
        if (D == 4) {
            reinterpret_cast<uint4*>(out)[q] = make_uint4(wds[0], wds[1], wds[2], wds[D - 1]);
        } else {
            uint32_t* p = reinterpret_cast<uint32_t*>(out) + q * 3;
            p[0] = wds[0]; p[1] = wds[1]; p[2] = wds[2];
        }
        if (CONV) {   /* (unsigned char) of the already clamped conv III plane */
            reinterpret_cast<uint32_t*>(conv)[q] = __float2uint_rz(Y[0]) | (__float2uint_rz(Y[1]) << 8) |
                                                   (__float2uint_rz(Y[2]) << 16) | (__float2uint_rz(Y[3]) << 24);
        }
    } else {
        for (size_t i = i0; i < npix && i < i0 + 4; i++) {
            uint32_t r, g, b;
            rgb_of(yp[i], cb[i], cr[i], r, g, b);
            out[i * D] = (uint8_t)r; out[i * D + 1] = (uint8_t)g; out[i * D + 2] = (uint8_t)b;
            if (D == 4) out[i * D + 3] = (uint8_t)to_u8(al[i]);
            if (CONV) conv[i] = (uint8_t)__float2uint_rz(yp[i]);
        }
    }
}

void launch_color_out(const float* yp, const float* cb, const float* cr, const float* a, size_t npix, unsigned d,
                      uint8_t* out, uint8_t* conv, cudaStream_t s)
{
    const size_t quads = (npix + 3) / 4;
    const unsigned blocks = (unsigned)((quads + 255) / 256);
    auto al16 = [](const void* p) { return ((uintptr_t)p % 16) == 0; };
    const bool vec = al16(yp) && al16(cb) && al16(cr) && (d != 4 || al16(a)) &&
                     (((uintptr_t)out % (d == 4 ? 16 : 4)) == 0) && (!conv || ((uintptr_t)conv % 4) == 0);
#define SRCNN_CO(D_, C_, V_) color_out_kernel<D_, C_, V_><<<blocks, 256, 0, s>>>(yp, cb, cr, a, npix, out, conv)
    if (d == 4) {
        if (conv) { if (vec) SRCNN_CO(4, true, true);  else SRCNN_CO(4, true, false); }
        else      { if (vec) SRCNN_CO(4, false, true); else SRCNN_CO(4, false, false); }
    } else {
        if (conv) { if (vec) SRCNN_CO(3, true, true);  else SRCNN_CO(3, true, false); }
        else      { if (vec) SRCNN_CO(3, false, true); else SRCNN_CO(3, false, false); }
    }
#undef SRCNN_CO
}


/* ------------------------------------------------------------------------------------------------------------
 * Colour-out with the chroma / alpha resampling folded in (magnification where the bilinear / box tables are "pow2pair",
 * i.e. x2: weights 0.25 / 0.75 / 1 / 0).  Cb, Cr (, A) arrive at SOURCE resolution (written by the fused front end) and are
 * resampled here exactly as FRAWResizeEngine::scale does it -- vertical pass first, one rounding to FP32 after each pass
 * (src/frawscale.cpp:238-278) -- with fmaf(wA, vA, wB * vB): the product with the 0-or-power-of-two weight is exact, so
 * the fused multiply-add rounds the exact two-tap sum once, which is what the reference's double accumulation + cast
 * yields for these operands (prep.cu header).  A CTA owns 1024 pixels of one output row: it resamples the source columns
 * they touch vertically into shared memory (coalesced loads, each column once), then every thread mixes two of those per
 * pixel with its packed table entries (one 16-byte load per pixel) and runs the colour-out arithmetic for 4 pixels.
 * Saves the 4 B/px/plane write + read of the output-size chroma planes (66 MB of 4K RGB traffic each way). */
__device__ __forceinline__ float mix2f(float wA, float vA, float wB, float vB) { return __fmaf_rn(wA, vA, __fmul_rn(wB, vB)); }

constexpr int kUpBlockPx = 1024;              /* output pixels of one row per CTA (8 warps x 32 threads x 4 pixels) */
constexpr int kUpWarpPx = 128;                /* ... per warp */
constexpr int kUpIters = 3;                   /* staging passes of 32 columns */
constexpr int kUpMaxCols = 32 * kUpIters;     /* source columns a warp may need: magnification >= ~1.37 (x2: 66) */
constexpr int kUpRowsMax = 16;                /* consecutive output rows per CTA (the horizontal table entries stay in registers);
                                                 fewer for small images so that the grid still fills the SMs */

__device__ __forceinline__ int up_pos(int j) { return (j >> 1) + (j & 1) * (kUpMaxCols / 2); }

/* pack2 entry of a two-tap table (built on the host, srcnn_ctx.cu): x = left + (index of tap A), y = wA bits, z = wB bits,
 * w = left + (index of tap B); B is the tap whose weight is 0 or a power of two */
template <int D, bool CONV>
__global__ void __launch_bounds__(256)
color_out_up_kernel(const float* __restrict__ yp, const float* __restrict__ csrc, size_t cstride, unsigned w, unsigned src_row0,
                    const int4* __restrict__ vpack, const int4* __restrict__ hpack,
                    unsigned ow, unsigned row0, unsigned nrows, unsigned rpc, uint8_t* __restrict__ out, uint8_t* __restrict__ conv,
                    size_t yp_fstride, size_t csrc_fstride, size_t out_fstride, size_t conv_fstride)
{
    constexpr int NC = D - 1;
    if (gridDim.z > 1) {                              /* frame group: this CTA's frame */
        yp += (size_t)blockIdx.z * yp_fstride; csrc += (size_t)blockIdx.z * csrc_fstride;
        out += (size_t)blockIdx.z * out_fstride;
        if (CONV) conv += (size_t)blockIdx.z * conv_fstride;
    }
    /* the intermediate (vertically resampled) plane: every WARP stages the source columns of its own 128 pixels, so rows need
     * only __syncwarp and warps run independently.  One buffer per warp (two warp barriers per row): the shared-memory
     * addresses of a lane are then loop-invariant registers -- this kernel is bound by instruction issue, not by HBM */
    __shared__ float tS[8][NC][kUpMaxCols];
    const unsigned wq = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned xw = blockIdx.x * kUpBlockPx + wq * kUpWarpPx;     /* first pixel of this warp */
    if (xw >= ow) return;                                             /* (whole warps only: no block barrier below) */
    const unsigned xe = min(xw + kUpWarpPx, ow) - 1;                  /* last pixel of this warp */
    const int4 hb = __ldg(hpack + xw), he = __ldg(hpack + xe);
    const int c_lo = min(hb.x, hb.w), ncols = max(he.x, he.w) - c_lo + 1;   /* source columns [c_lo, c_lo + ncols) */
    const unsigned x0 = xw + lane * 4;
    const bool active = x0 < ow;
    float (*t)[kUpMaxCols] = tS[wq];
    const float* ta[4]; const float* tb[4]; float hA[4], hB[4];
    #pragma unroll
    for (int i = 0; i < 4; i++) {
        const int4 hp = active ? __ldg(hpack + x0 + i) : make_int4(c_lo, 0, 0, c_lo);
        /* shared-memory position of staged column j: even columns in the first half of the row, odd ones in the second.  The
         * four pixels of a lane sit two source columns after its neighbour's, so a warp reads stride-2 columns: with the
         * natural order that is a 2-way bank conflict on every load (ncu, round 1: 42 % of this kernel's shared wavefronts) */
        ta[i] = &t[0][up_pos(hp.x - c_lo)]; tb[i] = &t[0][up_pos(hp.w - c_lo)];
        hA[i] = __int_as_float(hp.y); hB[i] = __int_as_float(hp.z);
    }
    /* where this lane stages its columns lane, lane + 32, lane + 64: up_pos(lane + 32 it) == up_pos(lane) + 16 it */
    float* const ts0 = &t[0][up_pos((int)lane)];
    const unsigned ry0 = blockIdx.y * rpc, ry1 = min(ry0 + rpc, nrows);
    /* 32-bit element offsets (checked on the host): one 64-bit multiply-add per source row instead of one per load */
    const unsigned cs = (unsigned)cstride, col0 = (unsigned)c_lo + lane;
    /* software pipeline over the rows: the source samples (and Y') of row r+1 are requested before row r is computed, so
     * their latency hides behind the colour arithmetic */
    float pa[kUpIters][NC], pb[kUpIters][NC], pvA = 0.f, pvB = 0.f;
    float4 py = make_float4(0.f, 0.f, 0.f, 0.f);
    auto prefetch = [&](unsigned ry) {
        const int4 vp = __ldg(vpack + row0 + ry);                     /* vertical taps of image row row0 + ry */
        pvA = __int_as_float(vp.y); pvB = __int_as_float(vp.z);
        const float* rA = csrc + ((unsigned)(vp.x - (int)src_row0) * w + col0);
        const float* rB = csrc + ((unsigned)(vp.w - (int)src_row0) * w + col0);
        #pragma unroll
        for (int k = 0; k < NC; k++) {
            const float* qa = rA + k * cs; const float* qb = rB + k * cs;
            #pragma unroll
            for (int it = 0; it < kUpIters; it++)
                if ((int)lane + 32 * it < ncols) { pa[it][k] = __ldg(qa + 32 * it); pb[it][k] = __ldg(qb + 32 * it); }
        }
        if (active) py = __ldg(reinterpret_cast<const float4*>(yp + ((size_t)ry * ow + x0)));
    };
    /* up to here only constant tables were read: as a programmatic dependent launch this much overlaps the conv kernel's
     * tail; Y' (and, two kernels back, the chroma samples) is read from here on */
    pdl_wait();
    if (ry0 < ry1) prefetch(ry0);
    uint32_t* o32 = reinterpret_cast<uint32_t*>(out) + ((size_t)ry0 * ow + x0) / 4 * D;
    uint32_t* c32 = CONV ? reinterpret_cast<uint32_t*>(conv) + ((size_t)ry0 * ow + x0) / 4 : nullptr;
    const unsigned ostep = ow / 4 * D, cstep = ow / 4;
    for (unsigned ry = ry0; ry < ry1; ry++, o32 += ostep, c32 += cstep) {
        #pragma unroll
        for (int it = 0; it < kUpIters; it++) {
            if ((int)lane + 32 * it < ncols) {
                #pragma unroll
                for (int k = 0; k < NC; k++) ts0[16 * it + k * kUpMaxCols] = mix2f(pvA, pa[it][k], pvB, pb[it][k]);
            }
        }
        const float4 y4 = py;
        __syncwarp();
        if (ry + 1 < ry1) prefetch(ry + 1);
        if (active) {                       /* (a lane without pixels still staged its columns above) */
            const float Y[4] = { y4.x, y4.y, y4.z, y4.w };
            uint32_t bytes[4 * D];
            #pragma unroll
            for (int i = 0; i < 4; i++) {
                float c[NC];
                #pragma unroll
                for (int k = 0; k < NC; k++) c[k] = mix2f(hA[i], ta[i][k * kUpMaxCols], hB[i], tb[i][k * kUpMaxCols]);
                rgb_of(Y[i], c[0], c[1], bytes[i * D], bytes[i * D + 1], bytes[i * D + 2]);
                if (D == 4) bytes[i * D + 3] = to_u8(c[NC - 1]);
            }
            uint32_t wds[D];
            #pragma unroll
            for (int wi = 0; wi < D; wi++) wds[wi] = pack4(bytes[wi * 4], bytes[wi * 4 + 1], bytes[wi * 4 + 2], bytes[wi * 4 + 3]);
            if (D == 4) {
                *reinterpret_cast<uint4*>(o32) = make_uint4(wds[0], wds[1], wds[2], wds[D - 1]);
            } else {
                o32[0] = wds[0]; o32[1] = wds[1]; o32[2] = wds[2];
            }
            if (CONV) *c32 = pack4(__float2uint_rz(Y[0]), __float2uint_rz(Y[1]), __float2uint_rz(Y[2]), __float2uint_rz(Y[3]));
        }
        __syncwarp();                       /* every lane has read this row's columns before the next row overwrites them */
    }
}

/* host-side: the source columns one warp (kUpWarpPx output pixels) reads fit its shared-memory row, and lefts are monotonic */
bool pdl_enabled()
{
    static const bool on = [] { const char* e = getenv("SRCNN_B200_PDL"); return !(e && *e == '0'); }();   /* thread-safe once */
    return on;
}

bool color_out_up_quads_ok(const int* h_hcl, unsigned ow)
{
    if (ow % 4) return false;
    for (unsigned x = 0; x < ow; x += kUpWarpPx) {
        const unsigned e = (x + kUpWarpPx < ow ? x + kUpWarpPx : ow) - 1;
        if (h_hcl[e] + 1 - h_hcl[x] + 1 > kUpMaxCols) return false;
    }
    for (unsigned x = 1; x < ow; x++) if (h_hcl[x] < h_hcl[x - 1]) return false;
    return true;
}

bool launch_color_out_up(const float* yp, const float* csrc, size_t cstride, unsigned w, unsigned src_row0,
                         const AxisTableDev& vc, const AxisTableDev& hc, unsigned ow, unsigned row0, unsigned row1, unsigned d,
                         uint8_t* out, uint8_t* conv, cudaStream_t s, int num_sms,
                         unsigned nframes, size_t yp_fstride, size_t csrc_fstride, size_t out_fstride, size_t conv_fstride)
{
    auto al = [](const void* p, unsigned a) { return ((uintptr_t)p % a) == 0; };
    if (nframes == 0) nframes = 1;
    if (nframes > 1 && ((yp_fstride % 4) || (out_fstride % 16) || (conv && (conv_fstride % 4)))) return false;
    if (!vc.pack2 || !hc.pack2 || (ow % 4) || row1 <= row0) return false;
    if (cstride * (size_t)(d - 1) >= ((size_t)1 << 30)) return false;       /* 32-bit element offsets into the chroma planes */
    if (!al(yp, 16) || !al(out, d == 4 ? 16 : 4) || (conv && !al(conv, 4))) return false;
    const unsigned gx = (ow + kUpBlockPx - 1) / kUpBlockPx * nframes, nrows = row1 - row0;   /* (x nframes: for the wave estimate) */
    /* rows per CTA: pick the row count that minimises (number of waves of 4 CTAs per SM) x (rows + one row-time of prologue), i.e.
     * avoids a nearly empty last wave.  (Measured, round 2: modelling the RGBA instantiations' real residency of 3 CTAs per SM and
     * allowing up to 48 rows per CTA picks one wave of 38-row CTAs for a 1080-row strip and is SLOWER -- 53 vs 50 us -- than two
     * short waves of 16-row CTAs: CTAs are scheduled dynamically, many short ones balance better than the wave model predicts.) */
    const int sms = num_sms > 0 ? num_sms : 148;
    const size_t resident = (size_t)sms * 4;
    unsigned rpc = 2; size_t best = ~(size_t)0;
    for (unsigned r = 2; r <= (unsigned)kUpRowsMax; r++) {
        const size_t ctas = (size_t)gx * ((nrows + r - 1) / r), waves = (ctas + resident - 1) / resident;
        const size_t cost = waves * (r + 1);
        if (cost <= best) { best = cost; rpc = r; }
    }
    dim3 grid(gx / nframes, (nrows + rpc - 1) / rpc, nframes);
    if (grid.y > 65535u || grid.z > 65535u) return false;
#define SRCNN_COU(D_, C_) launch_pdl(color_out_up_kernel<D_, C_>, grid, dim3(256), 0, s, yp, csrc, cstride, w, src_row0, \
        reinterpret_cast<const int4*>(vc.pack2), reinterpret_cast<const int4*>(hc.pack2), ow, row0, nrows, rpc, out, conv, yp_fstride, csrc_fstride, out_fstride, conv_fstride)
    if (d == 4) { if (conv) SRCNN_COU(4, true); else SRCNN_COU(4, false); }
    else        { if (conv) SRCNN_COU(3, true); else SRCNN_COU(3, false); }
#undef SRCNN_COU
    return true;
}

} // namespace srcnn
