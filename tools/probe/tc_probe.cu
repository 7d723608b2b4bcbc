// Standalone probe of the tcgen05 primitives the fused conv kernel relies on (development tool, not product):
//   D[128 x N] = A[128 x K] * B[N x K]^T, fp16 operands, fp32 accumulate in TMEM
//   mode 0: A from shared memory (SS), K-major no-swizzle canonical layout
//   mode 1: A from tensor memory (TS), written by tcgen05.st 32x32b (lane = row, column = 2 packed K elements)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tc_probe.cu ; run: ./tc_probe
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;          // descriptor version 1 (Blackwell)
    return d;                        // layout_type = 0 (no swizzle), base_offset = 0
}
__device__ __forceinline__ uint32_t make_idesc(int N)
{
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // f16 x f16 -> f32, K-major A and B
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" :: "r"(bar), "r"(parity) : "memory");
}

template <int N, int K>
__global__ void __launch_bounds__(128) probe(const __half* A, const __half* Bimg, float* D, int mode)
{
    __shared__ __align__(128) unsigned char sB[N * K * 2];
    __shared__ __align__(128) unsigned char sA[128 * K * 2];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int t = threadIdx.x, warp = t >> 5;
    constexpr int KG = K / 8;                 // core matrices along K
    // B image is already in canonical layout (host-built): plain copy
    for (int i = t; i < N * K * 2 / 4; i += 128) ((uint32_t*)sB)[i] = ((const uint32_t*)Bimg)[i];
    // A (row-major in global) -> canonical K-major no-swizzle: (m/8)*SBO + (k/8)*128 + (m%8)*16 + (k%8)*2, SBO = KG*128
    for (int i = t; i < 128 * K; i += 128) {
        int m = i / K, k = i % K;
        *(__half*)(sA + (m / 8) * (KG * 128) + (k / 8) * 128 + (m % 8) * 16 + (k % 8) * 2) = A[i];
    }
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" :: "r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base_s;
    const uint32_t colD = 0, colA = 128;
    const uint32_t lane_addr = tb + ((uint32_t)(warp * 32) << 16);

    if (mode == 1) {
        // thread t owns row t: pack K halves into K/2 words, write 16 columns at a time
        const uint32_t* arow = (const uint32_t*)(A + (size_t)t * K);
        #pragma unroll
        for (int c0 = 0; c0 < K / 2; c0 += 16) {
            uint32_t r[16];
            #pragma unroll
            for (int j = 0; j < 16; j++) r[j] = arow[c0 + j];
            asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                         :: "r"(lane_addr + colA + c0),
                            "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                            "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (t == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = make_idesc(N);
        for (int s = 0; s < K / 16; s++) {
            const uint64_t bdesc = make_desc(smem_u32(sB) + s * 256, 128, KG * 128);
            const uint32_t acc = s > 0;
            if (mode == 0) {
                const uint64_t adesc = make_desc(smem_u32(sA) + s * 256, 128, KG * 128);
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                             :: "r"(tb + colD), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
            } else {
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                             :: "r"(tb + colD), "r"(tb + colA + s * 8), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(smem_u32(&bar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    #pragma unroll
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(lane_addr + colD + c0) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        #pragma unroll
        for (int j = 0; j < 16; j++) D[(size_t)t * N + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" :: "r"(tb) : "memory");
}

template <int N, int K>
int run(int mode)
{
    std::vector<__half> A(128 * K), Bimg(N * K);
    std::vector<float> Af(128 * K), Bf(N * K), ref(128 * N), out(128 * N);
    srand(7 + N + K);
    for (int i = 0; i < 128 * K; i++) { float v = (rand() % 2001 - 1000) / 250.0f; A[i] = __float2half(v); Af[i] = __half2float(A[i]); }
    for (int n = 0; n < N; n++) for (int k = 0; k < K; k++) {
        float v = (rand() % 2001 - 1000) / 1000.0f; __half hv = __float2half(v); Bf[n * K + k] = __half2float(hv);
        size_t off = (size_t)(n / 8) * ((K / 8) * 128) + (k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2;
        Bimg[off / 2] = hv;
    }
    for (int m = 0; m < 128; m++) for (int n = 0; n < N; n++) { double s = 0; for (int k = 0; k < K; k++) s += (double)Af[m * K + k] * Bf[n * K + k]; ref[m * N + n] = (float)s; }
    __half *dA, *dB; float* dD;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, Bimg.size() * 2); cudaMalloc(&dD, out.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, Bimg.data(), Bimg.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, out.size() * 4);
    probe<N, K><<<1, 128>>>(dA, dB, dD, mode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d K=%d mode=%d: CUDA error %s\n", N, K, mode, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int i = 0; i < 128 * N; i++) { double d = fabs((double)out[i] - ref[i]); if (!(d <= 1e-2)) bad++; if (d > maxerr || d != d) maxerr = d; }
    printf("N=%d K=%d mode=%s: max|err| = %.3e, bad = %d / %d   (sample out[0]=%f ref[0]=%f out[last]=%f ref[last]=%f)\n",
           N, K, mode ? "TS(A in TMEM)" : "SS(A in smem)", maxerr, bad, 128 * N, out[0], ref[0], out[128 * N - 1], ref[128 * N - 1]);
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return bad != 0;
}

int main()
{
    int f = 0;
    f |= run<64, 96>(0); f |= run<64, 96>(1);
    f |= run<32, 64>(0); f |= run<32, 64>(1);
    f |= run<32, 32>(0); f |= run<32, 32>(1);
    printf(f ? "PROBE FAILED\n" : "PROBE OK\n");
    return f;
}
