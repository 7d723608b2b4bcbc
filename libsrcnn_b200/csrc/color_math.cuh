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
    /* The reference computes 45.f * fCr / 32.f, (11.f * fCb + 23.f * fCr) / 32.f and 113.f * fCb / 64.f.  Dividing by a power
     * of two is exact and commutes with every rounding as long as nothing underflows: fl(45 c) / 32 == fl((45/32) c) and
     * fl(fl(11 c) + fl(23 d)) / 32 == fl(fl((11/32) c) + fl((23/32) d)), with 45/32, 11/32, 23/32 and 113/64 exact in
     * FP32.  Nothing underflows: cb - 128 and cr - 128 are 0 or >= 2^-17 in magnitude.  So each division folds into its
     * multiplication, bit for bit (tests: test_stage_color_out_bit_exact and every exact-mode image test). */
    const float fcb = __fsub_rn(cb, 128.f), fcr = __fsub_rn(cr, 128.f);
    r = to_u8(__fadd_rn(y, __fmul_rn(1.40625f, fcr)));
    g = to_u8(__fsub_rn(y, __fadd_rn(__fmul_rn(0.34375f, fcb), __fmul_rn(0.71875f, fcr))));
    b = to_u8(__fadd_rn(y, __fmul_rn(1.765625f, fcb)));
}

/* four byte values (each 0..255 in its own register) -> one little-endian word, three byte permutes */
__device__ __forceinline__ uint32_t pack4(uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3)
{
    return __byte_perm(__byte_perm(b0, b1, 0x0040), __byte_perm(b2, b3, 0x0040), 0x5410);
}

} // namespace srcnn
#endif
