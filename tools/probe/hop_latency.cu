// Micro-benchmark: the latencies a warp-specialised tcgen05 pipeline is built from (one CTA, otherwise idle SM).
//   ping-pong   : warp A arrives on mbarrier X, warp B observes X and arrives on Y, A observes Y  -> arrive-to-observe latency
//   commit      : 1 MMA (M=128 N=16 K=16) + tcgen05.commit, the same lane observes the mbarrier       -> MMA + commit + observe
//   commit-other: the commit is observed by ANOTHER warp, which answers with a plain arrive
//   tmem ld/st  : dependent tcgen05.ld / tcgen05.st + wait loops of several shapes
// Development tool (informs conv_tc.cu); not part of the product.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
template <int MODE> __device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    if (MODE == 0)
        asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" :: "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
    else if (MODE == 1)
        asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" :: "r"(bar), "r"(parity) : "memory");
    else
        asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory"); }
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mma16(uint32_t d, uint32_t a, uint64_t b, uint32_t acc)
{
    const uint32_t idesc = (1u << 4) | ((uint32_t)(16 >> 3) << 17) | (8u << 24);
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                 :: "r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
#define LD(N, regs, addr) asm volatile("tcgen05.ld.sync.aligned.32x32b.x" #N ".b32 " regs ", [%" #N "];"
constexpr int IT = 2000;

template <int WM>
__global__ void __launch_bounds__(256) k(long long* out, uint32_t* sink)
{
    __shared__ __align__(128) unsigned char sB[16 * 16 * 2 * 4];
    __shared__ __align__(8) uint64_t bars[4];
    __shared__ uint32_t s_tmem;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    for (int i = t; i < (int)sizeof(sB) / 4; i += 256) ((uint32_t*)sB)[i] = 0;
    if (t == 0) { for (int i = 0; i < 4; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bars[i]))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&s_tmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint32_t X = smem_u32(&bars[0]), Y = smem_u32(&bars[1]), Z = smem_u32(&bars[2]), W = smem_u32(&bars[3]);
    const uint64_t db = make_desc(smem_u32(sB), 128, 256);
    uint32_t acc = 0;

    /* 1. ping-pong between warp 1 and warp 6 (different sub-partitions), single lanes */
    if (warp == 1 && lane == 0) {
        const long long t0 = clock64();
        for (int i = 0; i < IT; i++) { mbar_arrive(X); mbar_wait<WM>(Y, i & 1); }
        out[0] = clock64() - t0;
    } else if (warp == 6 && lane == 0) {
        for (int i = 0; i < IT; i++) { mbar_wait<WM>(X, i & 1); mbar_arrive(Y); }
    }
    __syncthreads();
    /* 1b. the same with whole warps (fence + __syncwarp + lane 0 arrives, all lanes wait): the conv kernel's hand-over */
    if (warp == 1) {
        const long long t0 = clock64();
        for (int i = 0; i < IT; i++) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncwarp(); if (lane == 0) mbar_arrive(Z);
            mbar_wait<WM>(W, i & 1); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        if (lane == 0) out[1] = clock64() - t0;
    } else if (warp == 6) {
        for (int i = 0; i < IT; i++) {
            mbar_wait<WM>(Z, i & 1); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncwarp(); if (lane == 0) mbar_arrive(W);
        }
    }
    __syncthreads();
    /* 2. MMA + commit observed by the issuing lane itself */
    if (warp == 1 && lane == 0) {
        const long long t0 = clock64();
        for (int i = 0; i < IT; i++) {
            mma16(tmem, tmem + 480, db, 0);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(X) : "memory");
            mbar_wait<WM>(X, i & 1);
        }
        out[2] = clock64() - t0;
    }
    __syncthreads();
    /* 2b. commit observed by another warp, which answers with an arrive (one full issuer <-> worker round trip, no TMEM access) */
    if (warp == 1) {
        const long long t0 = clock64();
        for (int i = 0; i < IT; i++) {
            if (elect_one()) {
                mma16(tmem, tmem + 480, db, 0);
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(Y) : "memory");
            }
            __syncwarp();
            mbar_wait<WM>(Z, i & 1); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        if (lane == 0) out[3] = clock64() - t0;
    } else if (warp == 6) {
        for (int i = 0; i < IT; i++) {
            mbar_wait<WM>(Y, i & 1); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncwarp(); if (lane == 0) mbar_arrive(Z);
        }
    }
    __syncthreads();
    /* 2c. as 2b, and the worker drains 32 accumulator columns (x32 load + wait) before it answers */
    if (warp == 1) {
        const long long t0 = clock64();
        for (int i = 0; i < IT; i++) {
            if (elect_one()) {
                mma16(tmem, tmem + 480, db, 0);
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(W) : "memory");
            }
            __syncwarp();
            mbar_wait<WM>(X, i & 1); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        if (lane == 0) out[4] = clock64() - t0;
    } else if (warp == 5) {                   /* lane quarter 1 */
        const uint32_t lb = tmem + (32u << 16);
        for (int i = 0; i < IT; i++) {
            uint32_t r[32];
            mbar_wait<WM>(W, i & 1); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(lb) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncwarp(); if (lane == 0) mbar_arrive(X);
            acc += r[0] + r[31];
        }
    }
    __syncthreads();
    /* 3. TMEM access latencies, one warp, dependent iterations */
    if (warp == 4) {
        const uint32_t lb = tmem;
        uint32_t r[32];
        #pragma unroll
        for (int i = 0; i < 32; i++) r[i] = lane + i;
        long long t0 = clock64();
        for (int i = 0; i < IT; i++) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(lb + (r[0] & 1)) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        if (lane == 0) out[5] = clock64() - t0;
        t0 = clock64();
        for (int i = 0; i < IT; i++) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(lb + (r[0] & 1)) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        if (lane == 0) out[6] = clock64() - t0;
        t0 = clock64();
        for (int i = 0; i < IT; i++) {
            asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(lb + 64), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                    "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        if (lane == 0) out[7] = clock64() - t0;
        t0 = clock64();
        for (int i = 0; i < IT; i++) {
            asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" :: "r"(lb + 64), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
            asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" :: "r"(lb + 68), "r"(r[4]), "r"(r[5]) : "memory");
            asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" :: "r"(lb + 70), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]) : "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        if (lane == 0) out[8] = clock64() - t0;
        #pragma unroll
        for (int i = 0; i < 32; i++) acc += r[i];
    }
    sink[t] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}
template <int WM> void run(long long* d, uint32_t* sink)
{
    static const char* wm[] = { "try_wait + suspend hint", "try_wait", "test_wait (spin)" };
    cudaMemset(d, 0, 128);
    k<WM><<<1, 256>>>(d, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[9]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("wait = %-24s (%s)\n", wm[WM], cudaGetErrorString(e));
    printf("  mbarrier ping-pong, single lanes            : %6.1f clk per one-way hand-over\n", (double)h[0] / IT / 2);
    printf("  ... whole warps (fences, syncwarp, arrive)  : %6.1f clk per one-way hand-over\n", (double)h[1] / IT / 2);
    printf("  MMA(N=16) + commit -> same lane observes    : %6.1f clk\n", (double)h[2] / IT);
    printf("  elect/MMA/commit -> worker -> arrive -> issuer: %6.1f clk per round trip\n", (double)h[3] / IT);
    printf("  ... worker drains 32 columns (ld x32) first : %6.1f clk per round trip\n", (double)h[4] / IT);
    printf("  tcgen05.ld x16 + wait %6.1f | ld x32 + wait %6.1f | st x16 + wait %6.1f | st x4,x2,x4 + wait %6.1f clk\n",
           (double)h[5] / IT, (double)h[6] / IT, (double)h[7] / IT, (double)h[8] / IT);
}
int main()
{
    long long* d; uint32_t* sink; cudaMalloc(&d, 128); cudaMalloc(&sink, 1024);
    run<0>(d, sink); run<1>(d, sink); run<2>(d, sink);
    return 0;
}
