// Micro-benchmark: does the tensor pipe slow down when other warps use TMEM / shared memory?
// Warp 1 issues M=128 N=64 K=16 MMAs (A in TMEM, B in smem) back to back for a fixed count and times itself; NV victim warps
// run one of: nothing | tcgen05.ld x32 + wait | tcgen05.st x16 + wait | ld.shared (conflict-free) | ld.shared (8-way conflicts)
// Development tool (informs conv_tc.cu); not part of the product.
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
__device__ __forceinline__ void mma(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                 :: "r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
constexpr int NMMA = 6000;
template <int VW, int N>
__global__ void __launch_bounds__(768) k(long long* out, uint32_t* sink, int nvict)
{
    __shared__ __align__(128) unsigned char sB[64 * 16 * 2 * 4];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    __shared__ volatile int stop;
    __shared__ uint32_t sdat[4096];
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    for (int i = t; i < (int)sizeof(sB) / 4; i += 768) ((uint32_t*)sB)[i] = 0;
    for (int i = t; i < 4096; i += 768) sdat[i] = 0;
    if (t == 0) { stop = 0; asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&s_tmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    const uint64_t db = make_desc(smem_u32(sB), 128, 256);
    if (warp == 1) {
        if (elect_one()) {
            const long long t0 = clock64();
            #pragma unroll 1
            for (int i = 0; i < NMMA / 6; i++) {
                #pragma unroll
                for (int s = 0; s < 6; s++) mma(tmem + (i & 1) * 64, tmem + 480 + 8 * (s & 1), db + 16 * (s & 3), idesc, s > 0);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
            mbar_wait(smem_u32(&bar), 0);
            const long long t1 = clock64();
            if (blockIdx.x == 0) out[0] = t1 - t0;
            stop = 1;
        }
        __syncwarp();
    } else if (warp >= 4 && warp < 4 + nvict && VW > 0) {
        const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t acc = lane; long long n = 0;
        const long long t0 = clock64();
        while (!stop) {
            if (VW == 1) {
                uint32_t r[32];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(lane_base + 256) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc += r[0] + r[31];
            } else if (VW == 2) {
                asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" :: "r"(lane_base + 256 + 16 * (warp >> 2)), "r"(acc) : "memory");
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            } else {
                const int stride = VW == 3 ? 1 : 8;
                #pragma unroll
                for (int j = 0; j < 8; j++) acc += sdat[((lane * stride) + j * 32 + (acc & 1)) & 4095];
            }
            n++;
        }
        const long long t1 = clock64();
        if (lane == 0 && blockIdx.x == 0 && warp == 4) { out[1] = t1 - t0; out[2] = n; }
        sink[blockIdx.x * 768 + t] = acc;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}
template <int VW, int N> void run(long long* d, uint32_t* sink, int nvict)
{
    static const char* vw[] = { "none", "tcgen05.ld x32 + wait", "tcgen05.st x16 + wait", "ld.shared x8 conflict-free", "ld.shared x8 8-way conflicts" };
    cudaMemset(d, 0, 64);
    k<VW, N><<<148, 768>>>(d, sink, nvict);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[3]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("N=%2d  %2d victim warps: %-30s : %6.2f clk per MMA   (victim: %.1f clk per iteration)  (%s)\n", N, VW ? nvict : 0, vw[VW],
           (double)h[0] / NMMA, h[2] ? (double)h[1] / h[2] : 0.0, cudaGetErrorString(e));
}
int main()
{
    long long* d; uint32_t* sink; cudaMalloc(&d, 64); cudaMalloc(&sink, 148 * 768 * 4);
    run<0, 64>(d, sink, 0);
    for (int nv : {4, 16}) { run<1, 64>(d, sink, nv); run<2, 64>(d, sink, nv); run<3, 64>(d, sink, nv); run<4, 64>(d, sink, nv); }
    run<0, 16>(d, sink, 0);
    for (int nv : {16}) { run<1, 16>(d, sink, nv); run<2, 16>(d, sink, nv); run<3, 16>(d, sink, nv); run<4, 16>(d, sink, nv); }
    return 0;
}
