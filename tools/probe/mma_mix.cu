// Micro-benchmark: does the tensor pipe lose time when tcgen05.mma groups of DIFFERENT shapes / accumulators, issued by
// different warps, interleave?  Per "row" the conv kernel issues 6 x (N=64) + 5 x (N=32) + 6 x (N=16) K=16 steps (A in
// TMEM).  Modes: 0 = one warp issues the three groups back to back; 1 = three warps issue their groups concurrently with
// no ordering between them; 2 = three warps, each group taken under a shared-memory ticket lock (groups stay contiguous).
// Development tool (informs the issuer structure of conv_tc.cu); not part of the product.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                 :: "r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ constexpr uint32_t idesc_of(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | (8u << 24); }

template <int N, int KS>
__device__ __forceinline__ void group(uint32_t d, uint32_t a, uint64_t db)
{
    #pragma unroll
    for (int s = 0; s < KS; s++) mma_ts(d, a + 8 * (s & 1), db + 16 * (s & 3), idesc_of(N), s > 0);
}

template <int MODE, int COMMIT>
__global__ void __launch_bounds__(160) k(int iters, long long* out)
{
    __shared__ __align__(128) unsigned char sB[256 * 16 * 2 * 4];
    __shared__ __align__(8) uint64_t bar[3];
    __shared__ __align__(8) uint64_t cbar[3];     /* COMMIT = 1: a tcgen05.commit after every group (as the conv kernel does), never waited for */
    __shared__ uint32_t s_tmem;
    __shared__ volatile int ticket;
    const int t = threadIdx.x, warp = t >> 5;
    for (int i = t; i < (int)sizeof(sB) / 4; i += blockDim.x) ((uint32_t*)sB)[i] = 0;
    if (t == 0) {
        for (int i = 0; i < 3; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar[i])));
        for (int i = 0; i < 3; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&cbar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        ticket = 0;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&s_tmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint64_t db = make_desc(smem_u32(sB), 128, 256);
    const uint32_t A = tmem + 480;
    const long long t0 = clock64();
    if (MODE == 0) {
        if (warp == 1 && elect_one()) {
            #pragma unroll 1
            for (int it = 0; it < iters; it++) {
                group<64, 6>(tmem + (it & 1) * 64, A, db);
                if (COMMIT) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&cbar[0])) : "memory");
                group<32, 5>(tmem + 128 + (it & 1) * 32, A, db);
                if (COMMIT) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&cbar[1])) : "memory");
                group<16, 6>(tmem + 192, A, db);
                if (COMMIT) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&cbar[2])) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar[0])) : "memory");
            mbar_wait(smem_u32(&bar[0]), 0);
            out[blockIdx.x] = clock64() - t0;
        }
    } else if (warp >= 1 && warp <= 3) {
        const int w = warp - 1;
        if (elect_one()) {
            #pragma unroll 1
            for (int it = 0; it < iters; it++) {
                if (MODE == 2) { while (ticket != 3 * it + w) { } }
                if (w == 0) group<64, 6>(tmem + (it & 1) * 64, A, db);
                else if (w == 1) group<32, 5>(tmem + 128 + (it & 1) * 32, A, db);
                else group<16, 6>(tmem + 192, A, db);
                if (COMMIT) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&cbar[w])) : "memory");
                if (MODE == 2) { __threadfence_block(); ticket = 3 * it + w + 1; }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar[w])) : "memory");
            mbar_wait(smem_u32(&bar[w]), 0);
            if (w == 0) out[blockIdx.x] = clock64() - t0;
            else out[148 * w + blockIdx.x] = clock64() - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}

template <int MODE, int COMMIT = 0> void run(const char* name)
{
    long long* d; cudaMalloc(&d, 3 * 148 * 8); cudaMemset(d, 0, 3 * 148 * 8);
    const int iters = 4000;
    k<MODE, COMMIT><<<148, 160>>>(iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[3 * 148]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < 3 * 148; i++) mx = h[i] > mx ? h[i] : mx;
    printf("%-44s : %7.1f clk per row (ideal 6*32 + 5*16 + 6*9.1 = 326.6)  (%s)\n", name, (double)mx / iters, cudaGetErrorString(e));
    cudaFree(d);
}
int main()
{
    run<0>("one issuer, groups back to back");
    run<1>("three issuers, unordered");
    run<2>("three issuers, ticket lock per group");
    run<0>("one issuer, groups back to back");
    run<1>("three issuers, unordered");
    run<0, 1>("one issuer, a commit after every group");
    run<1, 1>("three issuers, a commit after every group");
    return 0;
}
