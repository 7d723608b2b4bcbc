// Micro-benchmark: issue rate of tcgen05.mma (kind::f16, M=128, A in TMEM, B in smem) as a function of N.
// Development tool (informs the K/N shapes of conv_tc.cu); not part of the product.
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
template <int N, int SS, int KS>
__global__ void __launch_bounds__(128) k(int iters, long long* out)
{
    __shared__ __align__(128) unsigned char sB[256 * 16 * 2 * 4];
    __shared__ __align__(128) unsigned char sA[128 * 16 * 2 * 2];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    const int t = threadIdx.x, warp = t >> 5;
    for (int i = t; i < (int)sizeof(sB) / 4; i += 128) ((uint32_t*)sB)[i] = 0;
    for (int i = t; i < (int)sizeof(sA) / 4; i += 128) ((uint32_t*)sA)[i] = 0;
    if (t == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&s_tmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    if (warp == 1 && elect_one()) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
        const uint64_t db = make_desc(smem_u32(sB), 128, 256), da = make_desc(smem_u32(sA), 128, 256);
        const long long t0 = clock64();
        #pragma unroll 1
        for (int it = 0; it < iters; it++) {
            #pragma unroll
            for (int s = 0; s < KS; s++) {
                if (SS)
                    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                                 :: "r"(tmem + (it & 1) * 256), "l"(da), "l"(db + 16 * (s & 3)), "r"(idesc), "r"((uint32_t)(s > 0)) : "memory");
                else
                    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                                 :: "r"(tmem + (it & 1) * 256), "r"(tmem + 480 + 8 * (s & 1)), "l"(db + 16 * (s & 3)), "r"(idesc), "r"((uint32_t)(s > 0)) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
        mbar_wait(smem_u32(&bar), 0);
        const long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}
template <int N, int SS, int KS> void run1(int grid, long long* d)
{
    {
        const int ks = KS;
        const int iters = 4000;
        k<N, SS, KS><<<grid, 128>>>(iters, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
        long long mx = 0; for (int i = 0; i < grid; i++) mx = h[i] > mx ? h[i] : mx;
        printf("%s N=%3d grid=%3d ksteps/group=%2d : %6.2f clk per MMA (K=16)  (%s)\n", SS ? "SS" : "TS", N, grid, ks, (double)mx / ((double)iters * ks), cudaGetErrorString(e));
    }
}
template <int N, int SS> void run(int grid)
{
    long long* d; cudaMalloc(&d, 148 * 8);
    run1<N, SS, 2>(grid, d); run1<N, SS, 6>(grid, d); run1<N, SS, 10>(grid, d);
    cudaFree(d);
}
int main()
{
    for (int grid : {148}) {
        run<16, 0>(grid); run<32, 0>(grid); run<64, 0>(grid); run<128, 0>(grid); run<256, 0>(grid);
        run<32, 1>(grid); run<64, 1>(grid);
    }
    return 0;
}
