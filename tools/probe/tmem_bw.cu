// Micro-benchmark: throughput of tcgen05.ld / tcgen05.st (32x32b) per SM, as a function of resident warps.
// Development tool (informs the epilogue design of conv_tc.cu); not part of the product.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>   // 0: ld x32, 1: ld x64 (two x32 back to back, one wait), 2: st x16, 3: st x32(2 x16), 4: ld x32 + st x16
__global__ void k(int iters, long long* out, uint32_t* sink)
{
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&s_tmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = s_tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64) % 448;
    uint32_t r[32]; uint32_t acc = 0;
    #pragma unroll
    for (int i = 0; i < 32; i++) r[i] = threadIdx.x + i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (MODE == 0 || MODE == 1 || MODE == 4) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(base) : "memory");
            if (MODE == 1) {
                uint32_t q[32];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]),
                   "=r"(q[8]), "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15]),
                   "=r"(q[16]), "=r"(q[17]), "=r"(q[18]), "=r"(q[19]), "=r"(q[20]), "=r"(q[21]), "=r"(q[22]), "=r"(q[23]),
                   "=r"(q[24]), "=r"(q[25]), "=r"(q[26]), "=r"(q[27]), "=r"(q[28]), "=r"(q[29]), "=r"(q[30]), "=r"(q[31])
                 : "r"(base + 32) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                #pragma unroll
                for (int i = 0; i < 32; i++) acc += q[i];
            } else {
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            }
            #pragma unroll
            for (int i = 0; i < 32; i++) acc += r[i];
        }
        if (MODE == 2 || MODE == 3 || MODE == 4) {
            asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(base), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                    "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
            if (MODE == 3)
                asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(base + 16), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
                    "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            r[0] += it;
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x % 32 == 0) out[warp] = t1 - t0;
    sink[threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(s_tmem) : "memory");
}

template <int MODE> void run(const char* name, int bytes_per_warp_iter)
{
    long long* d; uint32_t* sink; cudaMalloc(&d, 64 * 8); cudaMalloc(&sink, 1024 * 4);
    for (int nw : {4, 8, 16, 32}) {
        const int iters = 2000;
        k<MODE><<<1, nw * 32>>>(iters, d, sink);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[32]; cudaMemcpy(h, d, nw * 8, cudaMemcpyDeviceToHost);
        long long mx = 0; for (int i = 0; i < nw; i++) mx = h[i] > mx ? h[i] : mx;
        printf("%-28s warps=%2d  %7.1f clk/iter/warp  -> %6.1f B/clk/SM  (%s)\n", name, nw, (double)mx / iters,
               (double)bytes_per_warp_iter * nw * iters / mx, cudaGetErrorString(e));
    }
    cudaFree(d); cudaFree(sink);
}
int main()
{
    run<0>("ld 32x32b.x32 + wait", 32 * 32 * 4);
    run<1>("ld 2 x (x32) + wait", 64 * 32 * 4);
    run<2>("st 32x32b.x16 + wait", 16 * 32 * 4);
    run<3>("st 2 x (x16) + wait", 32 * 32 * 4);
    run<4>("ld x32 + st x16 + waits", 48 * 32 * 4);
    return 0;
}
