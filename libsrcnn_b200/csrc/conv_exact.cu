/* Fused conv I + II + III on CUDA cores in FP32, bit-exact against the reference -- the validation mode
 * (SRCNN_B200_CONV_EXACT) and the A/B baseline for the tcgen05 kernel (conv_tc.cu).
 *
 *  conv I   = reference convolution99 x64 (src/libsrcnn.cpp:350-422): Y clamp-padded by 4, taps row-major,
 *             float running sum from 0 with separate multiply and add, + bias, ReLU
 *  conv II  = reference convolution11 x32 (src/libsrcnn.cpp:424-447): sum over the 64 inputs in order, + bias, ReLU
 *  conv III = reference convolution55     (src/libsrcnn.cpp:449-529): conv II planes clamp-padded by 2 (their own
 *             coordinates), per input channel a double partial over 25 taps of float products, float running sum
 *             over channels, + bias, clamp to [0,255]
 *
 * One 512-thread CTA produces a 28x12 output tile: conv I/II are evaluated at the 32x16 in-image positions conv III needs, one
 * position per thread (the 64- and 32-channel intermediates live in registers / shared memory only), then conv III reads the
 * tile with the reference's coordinate clamping.  Conv I / II weights are staged in shared memory (see below).
 */
#include "kernels.h"
#include <atomic>
#include <mutex>
#include "srcnn_weights_data.h"

namespace srcnn {

__constant__ float cB1[64];
__constant__ float cB2[32];
__constant__ float cW3[25][32];
__constant__ float cB3;

int conv_exact_init()
{
    cudaError_t e;
    if ((e = cudaMemcpyToSymbol(cB1, SRCNN_B1, sizeof(SRCNN_B1))) != cudaSuccess) return (int)e;
    if ((e = cudaMemcpyToSymbol(cB2, SRCNN_B2, sizeof(SRCNN_B2))) != cudaSuccess) return (int)e;
    if ((e = cudaMemcpyToSymbol(cW3, SRCNN_W3, sizeof(SRCNN_W3))) != cudaSuccess) return (int)e;
    if ((e = cudaMemcpyToSymbol(cB3, &SRCNN_B3, sizeof(float))) != cudaSuccess) return (int)e;
    return 0;
}

/* Tile = 28 x 12 outputs: conv I / II are needed at the 32 x 16 = 512 in-image positions around it (conv III's +-2), one position
 * per thread of a 512-thread CTA.  Round 2: the weights come from SHARED memory as 16-byte broadcast loads and the Y window is
 * re-read from shared memory per tap instead of living in 81 registers.  ncu on the round-1 kernel (weights as constant-bank
 * loads through the uniform datapath, 124 registers, 16 warps per SM): 54 % of all stall samples were FMULs waiting for their
 * LDCU, 17 % the barrier behind a half-empty second pass over 400 positions; issue slots 40 % used.  Now ~64 registers, 32 warps
 * per SM, no idle lanes.  The arithmetic (order of every sum, separate multiply and add) is unchanged. */
constexpr int TW_ = 28, TH_ = 12;
constexpr int PW = TW_ + 4, PH = TH_ + 4;          /* 32 x 16 conv II positions */
constexpr int YW = TW_ + 12, YH = TH_ + 12;        /* 40 x 24 Y samples         */
constexpr int NTHR = PW * PH;                      /* 512 */
constexpr int L2P = NTHR + 1;                      /* channel pitch of the conv II tile in shared memory */
#ifndef SRCNN_EXACT_KB
#define SRCNN_EXACT_KB 16     /* measured: 4 -> 2.339 ms, 8 -> 2.244 ms, 16 -> 2.201 ms per 1920x1080 pass */
#endif
constexpr int KB = SRCNN_EXACT_KB;                              /* conv I filters per step: independent chains of dependent adds */
constexpr size_t EXACT_SMEM = (size_t)(YW * YH + 32 * L2P + 81 * 64 + 64 * 32) * sizeof(float);

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__global__ void __launch_bounds__(NTHR, 2)
conv_exact_kernel(const float* __restrict__ y, int dw, int dh, int win_row0, int win_row1,
                  float* __restrict__ out, int row0, int row1, const float* __restrict__ w1g, const float* __restrict__ w2g)
{
    extern __shared__ __align__(16) float smem[];
    float* sW1 = smem;                      /* [81][64] */
    float* sW2 = sW1 + 81 * 64;             /* [64][32] */
    float* sY  = sW2 + 64 * 32;             /* [YH][YW] */
    float* sL2 = sY + YW * YH;              /* [32][L2P] */

    const int tx0 = blockIdx.x * TW_, ty0 = row0 + blockIdx.y * TH_;
    const int tid = threadIdx.x;

    for (int i = tid; i < (81 * 64 + 64 * 32) / 4; i += NTHR)
        reinterpret_cast<float4*>(smem)[i] = __ldg(reinterpret_cast<const float4*>(i < 81 * 16 ? w1g : w2g - 81 * 64) + i);
    /* Y tile with the reference's clamp padding: sY[i][j] = Y(clamp(ty0-6+i), clamp(tx0-6+j)) */
    for (int i = tid; i < YW * YH; i += NTHR) {
        /* rows past the strip's window only feed outputs past row1, which are discarded: keep the read in bounds */
        const int r = clampi(clampi(ty0 - 6 + i / YW, 0, dh - 1), win_row0, win_row1 - 1);
        const int c = clampi(tx0 - 6 + i % YW, 0, dw - 1);
        sY[i] = __ldg(y + (size_t)(r - win_row0) * dw + c);
    }
    __syncthreads();

    /* conv I + II at the 32 x 16 positions (tile +-2); positions outside the image are never read */
    {
        const int pr = tid / PW, pc = tid % PW;
        const int gr = ty0 - 2 + pr, gc = tx0 - 2 + pc;
        if (gr >= 0 && gr < dh && gc >= 0 && gc < dw) {
            const uint32_t ywa = (uint32_t)__cvta_generic_to_shared(sY + pr * YW + pc);
            float acc2[32];
            #pragma unroll
            for (int n = 0; n < 32; n++) acc2[n] = 0.f;
            #pragma unroll 1
            for (int k = 0; k < 64; k += KB) {
                /* KB filters at a time: each filter's 81-term sum is a serial chain of dependent adds (the reference's order is
                 * kept), KB independent chains hide the add latency */
                float t[KB];
                #pragma unroll
                for (int j = 0; j < KB; j++) t[j] = 0.f;
                #pragma unroll
                for (int i = 0; i < 9; i++)
                    #pragma unroll
                    for (int jj = 0; jj < 9; jj++) {
                        float v;          /* (asm volatile: re-read per step -- hoisted out of the k loop the window would cost 81 registers) */
                        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(ywa + 4u * (uint32_t)(i * YW + jj)));
                        #pragma unroll
                        for (int q = 0; q < KB / 4; q++) {
                            const float4 w = *reinterpret_cast<const float4*>(sW1 + (i * 9 + jj) * 64 + k + 4 * q);
                            t[4 * q + 0] = __fadd_rn(t[4 * q + 0], __fmul_rn(w.x, v));
                            t[4 * q + 1] = __fadd_rn(t[4 * q + 1], __fmul_rn(w.y, v));
                            t[4 * q + 2] = __fadd_rn(t[4 * q + 2], __fmul_rn(w.z, v));
                            t[4 * q + 3] = __fadd_rn(t[4 * q + 3], __fmul_rn(w.w, v));
                        }
                    }
                /* the conv II sums take the KB results in filter order */
                #pragma unroll
                for (int j = 0; j < KB; j++) {
                    float u = __fadd_rn(t[j], cB1[k + j]);
                    u = (u >= 0.f) ? u : 0.f;
                    #pragma unroll
                    for (int n = 0; n < 32; n += 4) {
                        const float4 w = *reinterpret_cast<const float4*>(sW2 + (k + j) * 32 + n);
                        acc2[n + 0] = __fadd_rn(acc2[n + 0], __fmul_rn(u, w.x));
                        acc2[n + 1] = __fadd_rn(acc2[n + 1], __fmul_rn(u, w.y));
                        acc2[n + 2] = __fadd_rn(acc2[n + 2], __fmul_rn(u, w.z));
                        acc2[n + 3] = __fadd_rn(acc2[n + 3], __fmul_rn(u, w.w));
                    }
                }
            }
            #pragma unroll
            for (int n = 0; n < 32; n++) {
                const float t = __fadd_rn(acc2[n], cB2[n]);
                sL2[n * L2P + tid] = (t >= 0.f) ? t : 0.f;
            }
        }
    }
    __syncthreads();

    /* conv III: one output per thread (the first TW_ x TH_ = 336 threads) */
    if (tid >= TW_ * TH_) return;
    const int orow = ty0 + tid / TW_, ocol = tx0 + tid % TW_;
    if (orow >= row1 || ocol >= dw) return;
    int pidx[25];
    #pragma unroll
    for (int dy = 0; dy < 5; dy++)
        #pragma unroll
        for (int dx = 0; dx < 5; dx++) {
            const int rr = clampi(orow + dy - 2, 0, dh - 1), cc = clampi(ocol + dx - 2, 0, dw - 1);
            pidx[dy * 5 + dx] = (rr - (ty0 - 2)) * PW + (cc - (tx0 - 2));
        }
    float t = 0.f;
    #pragma unroll 1
    for (int ch = 0; ch < 32; ch++) {
        double tp = 0.0;
        #pragma unroll
        for (int tap = 0; tap < 25; tap++)
            tp = __dadd_rn(tp, (double)__fmul_rn(cW3[tap][ch], sL2[ch * L2P + pidx[tap]]));
        t = __double2float_rn(__dadd_rn((double)t, tp));
    }
    t = __fadd_rn(t, cB3);
    t = (t > 0.f) ? t : 0.f;
    t = (t < 255.f) ? t : 255.f;
    out[(size_t)(orow - row0) * dw + ocol] = t;
}

/* -> 0, or the CUDA error of the one-time per-device set-up (weights copy, shared-memory opt-in); launch errors surface through
 * cudaGetLastError like those of every other kernel */
int launch_conv_exact(const float* y, unsigned dw, unsigned dh, unsigned win_row0, unsigned win_rows,
                      float* out, unsigned row0, unsigned row1, cudaStream_t s)
{
    /* the opt-in and the global copy of the conv I / II weights are per device: set up once per device, not on every launch */
    static std::atomic<bool> configured[64];
    static float* d_w[64];
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return (int)cudaErrorInvalidDevice;
    if (!configured[dev].load(std::memory_order_acquire)) {
        std::lock_guard<std::mutex> lock(mu);
        if (!configured[dev].load(std::memory_order_relaxed)) {
            float* w = nullptr;
            cudaError_t e = cudaMalloc(&w, (81 * 64 + 64 * 32) * sizeof(float));
            if (e != cudaSuccess) return (int)e;
            if ((e = cudaMemcpy(w, SRCNN_W1, sizeof(SRCNN_W1), cudaMemcpyHostToDevice)) != cudaSuccess ||
                (e = cudaMemcpy(w + 81 * 64, SRCNN_W2, sizeof(SRCNN_W2), cudaMemcpyHostToDevice)) != cudaSuccess ||
                (e = cudaFuncSetAttribute(conv_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EXACT_SMEM)) != cudaSuccess) {
                cudaFree(w); return (int)e;
            }
            d_w[dev] = w;
            configured[dev].store(true, std::memory_order_release);
        }
    }
    dim3 grid((dw + TW_ - 1) / TW_, (row1 - row0 + TH_ - 1) / TH_);
    conv_exact_kernel<<<grid, NTHR, EXACT_SMEM, s>>>(y, (int)dw, (int)dh, (int)win_row0, (int)(win_row0 + win_rows), out, (int)row0, (int)row1,
                                                      d_w[dev], d_w[dev] + 81 * 64);
    return 0;
}

} // namespace srcnn
