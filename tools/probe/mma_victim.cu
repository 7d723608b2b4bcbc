// Micro-benchmark: what does a tcgen05.mma issuer do to the OTHER warps of its SM sub-partition?
// Warp 1 (sub-partition 1) issues M=128 N=64 K=16 MMAs; warps 4..7 (one per sub-partition) run a fixed victim loop and time it.
//   issuer modes : 0 idle | 1 back to back (tensor queue always full) | 2 batches of 6 + commit + wait (queue never full)
//   victim loops : 0 integer ALU | 1 ld.shared | 2 tcgen05.st x4 + wait::st | 3 tcgen05.ld x16 + wait::ld
// Also prints the clock after each of the first 24 MMAs issued into an idle pipe (= depth of the tensor queue).
// Development tool (informs the warp placement of conv_tc.cu); not part of the product.
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
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" :: "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mma(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                 :: "r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
constexpr int VIT = 2000;
template <int IM, int VW>
__global__ void __launch_bounds__(256) k(long long* out, uint32_t* sink)
{
    __shared__ __align__(128) unsigned char sB[64 * 16 * 2 * 4];
    __shared__ __align__(8) uint64_t bar_[10];
    uint64_t& bar = bar_[0];
    __shared__ uint32_t s_tmem;
    __shared__ volatile int stop;
    __shared__ uint32_t sdat[1024];
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    for (int i = t; i < (int)sizeof(sB) / 4; i += 256) ((uint32_t*)sB)[i] = 0;
    for (int i = t; i < 1024; i += 256) sdat[i] = i;
    if (t == 0) { stop = 0; for (int i = 0; i < 10; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar) + 8 * i)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&s_tmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | (8u << 24);
    const uint64_t db = make_desc(smem_u32(sB), 128, 256);
    if (warp == 1 && IM == 5) {
        /* the conv kernel's issuer loop: per batch  if (elect) { 6 MMAs; commit; stamp }  __syncwarp();  stamp */
        long long ca[8], cb[8], cc[8];
        const long long c0 = clock64();
        #pragma unroll
        for (int b = 0; b < 8; b++) {
            ca[b] = cb[b] = 0;
            if (elect_one()) {
                const uint64_t dbb = make_desc(smem_u32(sB) + ((sdat[b] & 1) << 4), 128, 256);
                ca[b] = clock64();
                #pragma unroll
                for (int s = 0; s < 6; s++) mma(tmem + (b & 1) * 64, tmem + 480, dbb + 16 * (s & 3), idesc, s > 0);
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar) + 8 * (b + 1)) : "memory");
                cb[b] = clock64();
            }
            __syncwarp();
            cc[b] = clock64();
        }
        mbar_wait(smem_u32(&bar) + 8 * 8, 0);
        for (int b = 0; b < 8; b++) {
            long long x = ca[b], y = cb[b];
            for (int o = 16; o; o >>= 1) { x = max(x, __shfl_xor_sync(~0u, x, o)); y = max(y, __shfl_xor_sync(~0u, y, o)); }
            if (lane == 0 && blockIdx.x == 0) { out[16 + 3 * b] = x - c0; out[17 + 3 * b] = y - c0; out[18 + 3 * b] = cc[b] - c0; }
        }
    } else if (warp == 1 && IM == 6) {
        /* the conv kernel's issuer loop: per batch  if (elect) { 6 MMAs; commit; stamp }  __syncwarp();  stamp */
        long long ca[8], cb[8], cc[8];
        const long long c0 = clock64();
        #pragma unroll
        for (int b = 0; b < 8; b++) {
            ca[b] = cb[b] = 0;
            if (elect_one()) {
                const uint64_t dbb = make_desc(smem_u32(sB) + ((sdat[b] & 1) << 4), 128, 256);
                ca[b] = clock64();
                #pragma unroll
                for (int s = 0; s < 6; s++) mma(tmem + (b & 1) * 64, tmem + 480, dbb + 16 * (s & 3), idesc, s > 0);
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar) + 8 * (b + 1)) : "memory");
                cb[b] = clock64();
            }
            cc[b] = clock64();
        }
        mbar_wait(smem_u32(&bar) + 8 * 8, 0);
        for (int b = 0; b < 8; b++) {
            long long x = ca[b], y = cb[b];
            for (int o = 16; o; o >>= 1) { x = max(x, __shfl_xor_sync(~0u, x, o)); y = max(y, __shfl_xor_sync(~0u, y, o)); }
            if (lane == 0 && blockIdx.x == 0) { out[16 + 3 * b] = x - c0; out[17 + 3 * b] = y - c0; out[18 + 3 * b] = cc[b] - c0; }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            if (IM == 3) {
                /* queue depth: issue 24 MMAs into an idle pipe, clock after each */
                long long c[25]; c[0] = clock64();
                #pragma unroll
                for (int s = 0; s < 24; s++) { mma(tmem, tmem + 480, db, idesc, 1); c[s + 1] = clock64(); }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
                mbar_wait(smem_u32(&bar), 0);
                if (blockIdx.x == 0) for (int s = 0; s < 25; s++) out[16 + s] = c[s] - c[0];
            } else if (IM == 4) {
                /* issuer-loop emulation: batches of 6 MMAs + commit, NO wait for completion; clock after the commit and after
                 * the next batch's uniform-register setup (does the warp run ahead of the tensor pipe?) */
                long long c[17]; c[0] = clock64();
                uint32_t a_dummy = sdat[3];
                #pragma unroll
                for (int b = 0; b < 8; b++) {
                    const uint64_t dbb = make_desc(smem_u32(sB) + ((sdat[b] & 1) << 4), 128, 256);     /* defeats hoisting of the descriptor setup */
                    #pragma unroll
                    for (int s = 0; s < 6; s++) mma(tmem + (b & 1) * 64, tmem + 480, dbb + 16 * (s & 3), idesc, s > 0);
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar) + 8 * (b + 1)) : "memory");
                    c[2 * b + 1] = clock64();
                    a_dummy = a_dummy * 3 + 1;
                    c[2 * b + 2] = clock64();
                }
                mbar_wait(smem_u32(&bar) + 8 * 8, 0);
                sink[blockIdx.x] = a_dummy;
                if (blockIdx.x == 0) for (int s = 0; s < 17; s++) out[16 + s] = c[s] - c[0];
            } else if (IM == 1) {
                while (!stop) {
                    #pragma unroll
                    for (int s = 0; s < 8; s++) mma(tmem, tmem + 480, db, idesc, 1);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
                mbar_wait(smem_u32(&bar), 0);
            } else if (IM == 2) {
                uint32_t ph = 0;
                while (!stop) {
                    #pragma unroll
                    for (int s = 0; s < 6; s++) mma(tmem, tmem + 480, db, idesc, 1);
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
                    mbar_wait(smem_u32(&bar), ph); ph ^= 1;
                }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t a0 = lane, a1 = lane * 3, a2 = lane * 5, a3 = lane * 7;
        __syncwarp();
        const long long t0 = clock64();
        #pragma unroll 1
        for (int i = 0; i < VIT; i++) {
            if (VW == 0) {
                #pragma unroll
                for (int j = 0; j < 8; j++) { a0 = a0 * 3 + a1; a1 = a1 * 5 + a2; a2 = a2 * 7 + a3; a3 = a3 * 9 + a0; }
            } else if (VW == 1) {
                #pragma unroll
                for (int j = 0; j < 4; j++) { a0 += sdat[(a0 + j) & 1023]; }
            } else if (VW == 2) {
                asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" :: "r"(lane_base + 256), "r"(a0), "r"(a1), "r"(a2), "r"(a3) : "memory");
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            } else {
                uint32_t r[16];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(lane_base + 256) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                a0 += r[0] + r[15];
            }
        }
        const long long t1 = clock64();
        if (lane == 0 && blockIdx.x == 0) out[warp - 4] = t1 - t0;
        sink[blockIdx.x * 256 + t] = a0 + a1 + a2 + a3;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (warp == 4 && lane == 0) stop = 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}
template <int IM, int VW> void run(long long* d, uint32_t* sink)
{
    static const char* im[] = { "issuer idle", "issuer back to back (queue full)", "issuer batches of 6 + wait", "" };
    static const char* vw[] = { "ALU", "ld.shared", "tcgen05.st x4 + wait", "tcgen05.ld x16 + wait" };
    cudaMemset(d, 0, 64 * 8);
    k<IM, VW><<<148, 256>>>(d, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[48]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    if (IM == 6) {
        printf("the same WITHOUT the explicit __syncwarp: per batch [start, after commit, after the if]:");
        for (int b = 0; b < 8; b++) printf(" [%lld %lld %lld]", h[16 + 3 * b], h[17 + 3 * b], h[18 + 3 * b]);
        printf("  (%s)\n", cudaGetErrorString(e));
    } else if (IM == 5) {
        printf("issuer loop as in conv_tc (elect / 6 MMAs + commit / __syncwarp): per batch [start, after commit, after __syncwarp]:");
        for (int b = 0; b < 8; b++) printf(" [%lld %lld %lld]", h[16 + 3 * b], h[17 + 3 * b], h[18 + 3 * b]);
        printf("  (%s)\n", cudaGetErrorString(e));
    } else if (IM == 4) {
        printf("issuer loop, 8 batches of 6 x (N=64) + commit, no waits: clock after commit / after a following instruction:");
        for (int s = 1; s < 17; s++) printf(" %lld", h[16 + s]);
        printf("  (%s)\n", cudaGetErrorString(e));
    } else if (IM == 3) {
        printf("clock after each MMA issued into an idle pipe (N=64: 32 clk each to execute):");
        for (int s = 1; s < 25; s++) printf(" %lld", h[16 + s]);
        printf("  (%s)\n", cudaGetErrorString(e));
    } else
        printf("%-34s victim %-22s clk per iteration on sub-partition 0..3: %7.1f %7.1f %7.1f %7.1f  (%s)\n", im[IM], vw[VW],
               (double)h[0] / VIT, (double)h[1] / VIT, (double)h[2] / VIT, (double)h[3] / VIT, cudaGetErrorString(e));
}
int main()
{
    long long* d; uint32_t* sink; cudaMalloc(&d, 64 * 8); cudaMalloc(&sink, 148 * 256 * 4);
    run<3, 0>(d, sink);
    run<4, 0>(d, sink);
    run<5, 0>(d, sink);
    run<6, 0>(d, sink);
    run<0, 0>(d, sink); run<1, 0>(d, sink); run<2, 0>(d, sink);
    run<0, 1>(d, sink); run<1, 1>(d, sink); run<2, 1>(d, sink);
    run<0, 2>(d, sink); run<1, 2>(d, sink); run<2, 2>(d, sink);
    run<0, 3>(d, sink); run<1, 3>(d, sink); run<2, 3>(d, sink);
    return 0;
}
