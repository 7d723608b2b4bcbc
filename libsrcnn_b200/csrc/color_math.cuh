/* Internal: the reference's colour-out arithmetic (convertImgF32XtoImgU8, src/libsrcnn.cpp:274-308), shared by the
 * stand-alone colour-out kernel (color.cu) and the fused store stage of the tensor-core conv kernel (conv_tc.cu).
 * Unfused FP32 operations in the reference's association order -> bit-exact. */
#ifndef SRCNN_B200_COLOR_MATH_CUH
#define SRCNN_B200_COLOR_MATH_CUH

#include <stdint.h>

namespace srcnn {

__device__ __forceinline__ uint32_t to_u8(float f)
{
    /* (unsigned char)MAX(0.f, MIN(255.f, f)) with MIN(a,b) = a<b?a:b, truncating cast.  A float -> u8 conversion in PTX
     * clamps to the destination range, so one cvt.rzi.u8.f32 is that whole expression (no NaNs arise on this path). */
    uint32_t u;
    asm("cvt.rzi.u8.f32 %0, %1;" : "=r"(u) : "f"(f));
    return u;
}

__device__ __forceinline__ void rgb_of(float y, float cb, float cr, uint32_t& r, uint32_t& g, uint32_t& b)
{
    /* x / 32.f and x / 64.f are written as multiplications by the exact reciprocal: bit-identical to the IEEE division
     * whenever the quotient is a normal number, and it always is here (cb - 128 and cr - 128 are 0 or >= 2^-17 in
     * magnitude, so a non-zero numerator is >= 2^-17) */
    const float fcb = __fsub_rn(cb, 128.f), fcr = __fsub_rn(cr, 128.f);
    r = to_u8(__fadd_rn(y, __fmul_rn(__fmul_rn(45.f, fcr), 0.03125f)));
    g = to_u8(__fsub_rn(y, __fmul_rn(__fadd_rn(__fmul_rn(11.f, fcb), __fmul_rn(23.f, fcr)), 0.03125f)));
    b = to_u8(__fadd_rn(y, __fmul_rn(__fmul_rn(113.f, fcb), 0.015625f)));
}

} // namespace srcnn
#endif
