/* Colour-space kernels (HBM-bound, vectorised: 4 pixels per thread).
 *
 *  color_in  = reference converImgU8toYCbCr   (src/libsrcnn.cpp:233-272)
 *  color_out = reference convertImgF32XtoImgU8 (src/libsrcnn.cpp:274-308) + conv-output plane (:889-915)
 *
 * Every FP32 operation uses the round-to-nearest non-fused intrinsics in the reference's association order
 * (its x86-64 -O2 build has no FMA contraction), so both kernels are bit-exact against the oracle.
 */
#include "kernels.h"
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

} // namespace srcnn
