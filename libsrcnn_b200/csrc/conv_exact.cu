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
 * One CTA produces a 16x16 output tile: conv I/II are evaluated at the 20x20 in-image positions conv III needs
 * (the 64- and 32-channel intermediates live in registers / shared memory only), then conv III reads the tile
 * with the reference's coordinate clamping.
 */
#include "kernels.h"
#include <atomic>
#include "srcnn_weights_data.h"

namespace srcnn {

__constant__ float cW1[81][64];
__constant__ float cB1[64];
__constant__ float cW2[64][32];
__constant__ float cB2[32];
__constant__ float cW3[25][32];
__constant__ float cB3;

int conv_exact_init()
{
    cudaError_t e;
    if ((e = cudaMemcpyToSymbol(cW1, SRCNN_W1, sizeof(SRCNN_W1))) != cudaSuccess) return (int)e;
    if ((e = cudaMemcpyToSymbol(cB1, SRCNN_B1, sizeof(SRCNN_B1))) != cudaSuccess) return (int)e;
    if ((e = cudaMemcpyToSymbol(cW2, SRCNN_W2, sizeof(SRCNN_W2))) != cudaSuccess) return (int)e;
    if ((e = cudaMemcpyToSymbol(cB2, SRCNN_B2, sizeof(SRCNN_B2))) != cudaSuccess) return (int)e;
    if ((e = cudaMemcpyToSymbol(cW3, SRCNN_W3, sizeof(SRCNN_W3))) != cudaSuccess) return (int)e;
    if ((e = cudaMemcpyToSymbol(cB3, &SRCNN_B3, sizeof(float))) != cudaSuccess) return (int)e;
    return 0;
}

constexpr int TILE = 16;
constexpr int POS  = TILE + 4;          /* 20: conv II positions per tile edge */
constexpr int YT   = TILE + 12;         /* 28: Y samples per tile edge         */
constexpr int L2P  = POS * POS + 1;     /* channel pitch in shared memory (odd -> no bank conflicts) */

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__global__ void __launch_bounds__(256)
conv_exact_kernel(const float* __restrict__ y, int dw, int dh, int win_row0, int win_row1,
                  float* __restrict__ out, int row0, int row1)
{
    extern __shared__ float smem[];
    float* sY  = smem;                  /* [YT][YT]            */
    float* sL2 = smem + YT * YT;        /* [32][L2P]           */

    const int tx0 = blockIdx.x * TILE, ty0 = row0 + blockIdx.y * TILE;
    const int tid = threadIdx.x;

    /* Y tile with the reference's clamp padding: sY[i][j] = Y(clamp(ty0-6+i), clamp(tx0-6+j)) */
    for (int i = tid; i < YT * YT; i += 256) {
        /* rows past the strip's window only feed outputs past row1, which are discarded: keep the read in bounds */
        const int r = clampi(clampi(ty0 - 6 + i / YT, 0, dh - 1), win_row0, win_row1 - 1);
        const int c = clampi(tx0 - 6 + i % YT, 0, dw - 1);
        sY[i] = __ldg(y + (size_t)(r - win_row0) * dw + c);
    }
    __syncthreads();

    /* conv I + II at the 20x20 positions (tile +-2); positions outside the image are never read */
    for (int p = tid; p < POS * POS; p += 256) {
        const int pr = p / POS, pc = p % POS;
        const int gr = ty0 - 2 + pr, gc = tx0 - 2 + pc;
        if (gr < 0 || gr >= dh || gc < 0 || gc >= dw) continue;
        float win[81];
        #pragma unroll
        for (int i = 0; i < 9; i++)
            #pragma unroll
            for (int j = 0; j < 9; j++) win[i * 9 + j] = sY[(pr + i) * YT + pc + j];
        float acc2[32];
        #pragma unroll
        for (int n = 0; n < 32; n++) acc2[n] = 0.f;
        /* four conv I filters at a time: each filter's 81-term sum is a serial chain of dependent adds (the reference's order must
         * be kept), four independent chains hide the add latency; the conv II sums then take the four results in filter order */
        constexpr int KB = 4;
        #pragma unroll 1
        for (int k = 0; k < 64; k += KB) {
            float t[KB];
            #pragma unroll
            for (int j = 0; j < KB; j++) t[j] = 0.f;
            #pragma unroll
            for (int tap = 0; tap < 81; tap++)
                #pragma unroll
                for (int j = 0; j < KB; j++) t[j] = __fadd_rn(t[j], __fmul_rn(cW1[tap][k + j], win[tap]));
            #pragma unroll
            for (int j = 0; j < KB; j++) {
                float u = __fadd_rn(t[j], cB1[k + j]);
                u = (u >= 0.f) ? u : 0.f;
                #pragma unroll
                for (int n = 0; n < 32; n++) acc2[n] = __fadd_rn(acc2[n], __fmul_rn(u, cW2[k + j][n]));
            }
        }
        #pragma unroll
        for (int n = 0; n < 32; n++) {
            float t = __fadd_rn(acc2[n], cB2[n]);
            sL2[n * L2P + p] = (t >= 0.f) ? t : 0.f;
        }
    }
    __syncthreads();

    /* conv III: one output per thread */
    const int orow = ty0 + tid / TILE, ocol = tx0 + tid % TILE;
    if (orow >= row1 || ocol >= dw) return;
    int pidx[25];
    #pragma unroll
    for (int dy = 0; dy < 5; dy++)
        #pragma unroll
        for (int dx = 0; dx < 5; dx++) {
            const int rr = clampi(orow + dy - 2, 0, dh - 1), cc = clampi(ocol + dx - 2, 0, dw - 1);
            pidx[dy * 5 + dx] = (rr - (ty0 - 2)) * POS + (cc - (tx0 - 2));
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

void launch_conv_exact(const float* y, unsigned dw, unsigned dh, unsigned win_row0, unsigned win_rows,
                       float* out, unsigned row0, unsigned row1, cudaStream_t s)
{
    const size_t smem = (size_t)(YT * YT + 32 * L2P) * sizeof(float);
    /* the opt-in is a per-device function attribute: set it once per device, not on every launch */
    static std::atomic<bool> configured[64];
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64 && !configured[dev].load(std::memory_order_acquire)) {
        if (cudaFuncSetAttribute(conv_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess)
            configured[dev].store(true, std::memory_order_release);
    }
    dim3 grid((dw + TILE - 1) / TILE, (row1 - row0 + TILE - 1) / TILE);
    conv_exact_kernel<<<grid, 256, smem, s>>>(y, (int)dw, (int)dh, (int)win_row0, (int)(win_row0 + win_rows), out, (int)row0, (int)row1);
}

} // namespace srcnn
