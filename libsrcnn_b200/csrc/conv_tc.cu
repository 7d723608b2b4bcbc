/* Fused conv I + II + III on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), sm_100a only -- the product path.
 *
 * Replaces the reference's three OpenMP stages convolution99 x64 -> convolution11 x32 -> convolution55
 * (src/libsrcnn.cpp:350-529, drivers :783-846).  The 64- and 32-channel intermediates never leave the SM.
 *
 * Mapping.  A persistent CTA (one per SM) owns bands of 128 adjacent image columns (124 output columns + 2x2 halo for
 * conv III) and walks down row segments; TMEM lane t = image column c0 + t.  For every image row the three layers are
 * chained M=128 implicit GEMMs whose A operand lives in TENSOR MEMORY (written by tcgen05.st: lane = pixel,
 * column = two packed FP16 K elements) and whose B operand (the weights) sits in shared memory in the canonical
 * K-major no-swizzle core-matrix layout:
 *
 *   conv I  : A1[px][k = 10*slot + j] = Y16(row held in ring slot `slot`, c-4+j)   K = 96 (81 taps; the j = 9 column and
 *             k >= 92 carry zero weights; k = 90/91 are constant 1.0 inputs whose weights are the bias split hi/lo), N = 64.
 *             The 9-row window is a ring of 9 register slots per thread; the issuer picks one of 9 ROTATED weight images
 *             (row mod 9) instead of rolling the window, so a new row costs six 32-bit shared loads and only the two
 *             slots that changed since the same-parity row two rows earlier (10 TMEM columns) are rewritten.
 *   conv II : A2[px][k] = fp16(ReLU(D1[px][k]))          K = 64, N = 32, plus one K = 16 step for the bias whose A operand
 *             is the last K-step of A1 (it contains the constant-1 inputs; all other k meet zero weights).
 *   conv III: on the tensor core too, one SET per pair of output rows: K = 6 conv II rows x 32 channels out of an 8-row
 *             ring in TMEM, N = 16 (2 rows x 5 dx); the remaining 5-term dx sum crosses lanes through shared memory in
 *             the gather group, then bias, clamp, FP32 store.
 *
 * Warp specialisation (24 warps): TMA producer (Y16 row chunks -> 4-stage smem ring), three single-lane MMA issuers
 * (one per layer), and five 4-warp groups -- im2col, epilogue I x2 (one per row parity), epilogue II, gather/store --
 * chained by mbarriers with every TMEM operand / accumulator region double-buffered (A3: 4-pair ring), so the tensor pipe
 * works on row n while the groups prepare row n+1 and drain rows n-1, n-2.  The issuing lanes are the critical threads
 * of the pipeline: their instruction streams are kept as short as possible.
 *
 * Border semantics follow the reference exactly: conv I clamps Y coordinates (the Y16 plane is replicate-padded by
 * the front end), conv III clamps the coordinates of the conv II plane itself (src/libsrcnn.cpp:461-491), which
 * here becomes a choice of ring slot per K-step (rows) and a clamp on the lane index in the gather (columns).
 *
 * Arithmetic: FP16 operands (round-to-nearest), FP32 accumulation in TMEM, FP32 epilogues.  SURVEY.md 7.4 shows this
 * keeps |dY'| < 1 (final 8-bit output within 1 LSB); tests/test_parity_gpu.py asserts it against the oracle.
 */
#include "kernels.h"
#include "color_math.cuh"
#include "srcnn_weights_data.h"

#include <cuda.h>
#include <string>
#include <vector>
#include <cstring>
#include <cstdio>
#include <type_traits>


namespace srcnn {

namespace {

#ifndef SRCNN_TMEM0
#define SRCNN_TMEM0 0     /* measured by reading the SASS: the immediates land in vector registers first and need MORE R2UR, not fewer */
#endif
/* SRCNN_ELECT_OUTER = 1: an issuer warp elects its lane ONCE, around its whole row loop, instead of once per row: the per-row
 * ELECT / BSSY / BSYNC / BRA.DIV / WARPSYNC scaffolding (~9 of ~42 instructions per row) disappears from the single-lane issue
 * loops, which pace the pipeline (three issuers x ~45 serial instructions per image row at ~10 clk each = the 455 clk row period) */
#ifndef SRCNN_ELECT_OUTER
#define SRCNN_ELECT_OUTER 0     /* measured (round 2): 0.1320 ms vs 0.1280 ms -- 33 instead of 42 instructions per row in the issue loops, and slower */
#endif
/* SRCNN_ONE_LANE = 1: lanes 1..31 of the three issuer warps EXIT right after the register hand-back, so an issuer is a one-thread
 * warp: no elect, no divergent region, no reconvergence per row (tools/probe/mma_victim.cu: the elect / issue / __syncwarp loop
 * costs 222 clk per batch on an idle SM, 88 of them between the commit and the reconverged warp; a lane that runs the whole loop
 * alone issues at the pipe's 192).  Unlike SRCNN_ELECT_OUTER the other lanes are gone instead of parked at a convergence barrier. */
#ifndef SRCNN_ONE_LANE
#define SRCNN_ONE_LANE 0
#endif
#if SRCNN_ONE_LANE
#define ISSUER_BEGIN {
#define ISSUE_LANE() true
#define ISSUE_SYNC() ((void)0)
#elif SRCNN_ELECT_OUTER
#define ISSUER_BEGIN if (elect_one()) {
#define ISSUE_LANE() true
#define ISSUE_SYNC() ((void)0)
#else
#define ISSUER_BEGIN {
#define ISSUE_LANE() elect_one()
#define ISSUE_SYNC() __syncwarp()
#endif
#define ISSUER_END }
/* SRCNN_A1_EARLY = 1: the conv I issuer runs its six K-steps in a rotated order -- first the (at most three) K-steps that read
 * the two oldest window rows, i.e. the ten operand columns im2col overwrites next -- and commits to a second barrier (A1_FREE)
 * right after them; im2col waits for that barrier instead of for the whole row's MMAs and starts its partial rewrite while the
 * remaining K-steps are still running.  Needs SRCNN_A1_PARTIAL. */
#ifndef SRCNN_A1_EARLY
#define SRCNN_A1_EARLY 0
#endif
/* SRCNN_E1_SPLIT = 1: both epilogue I groups work on EVERY row, each on half of the 64 conv I channels (instead of one group
 * per row parity on all 64): the D1 -> A2 hand-over of a row takes half as long, which shortens the conv I -> conv II hop that
 * every row has to pass */
#ifndef SRCNN_E1_SPLIT
#define SRCNN_E1_SPLIT 0     /* measured (round 2): 0.1317 ms vs 0.1281 ms */
#endif
/* SRCNN_PI = 1 ("producer issues"): there are no dedicated MMA-issuing warps.  The group that writes a layer's A operand
 * synchronises on a named barrier (bar.sync over its 4 warps) and its first warp issues that layer's MMAs straight away:
 * im2col -> conv I, epilogue I -> conv II, epilogue II -> conv III.  This removes one mbarrier hand-over (arrive -> wake-up of
 * a sleeping issuer warp) from every layer of every row -- the pipeline is latency-bound on exactly those hops (the tensor pipe
 * idles ~28 % of the time while all three issuers wait for operands, profiles/r1_conv_wait_profile.txt). */
#ifndef SRCNN_PI
#define SRCNN_PI 0     /* measured (round 2, profiles/r2_conv_variants.txt): 0.2285 ms vs 0.1283 ms -- an issuing lane blocks while the tensor
                          queue is full, and with it the whole producer group at its next named barrier */
#endif
#ifndef SRCNN_E2_DECOUPLE
#define SRCNN_E2_DECOUPLE 0   /* 1: epilogue II drains both rows of a pair before waiting for the A3 slot (measured: no gain) */
#endif

constexpr int K1 = 96, N1 = 64, NROT = 9;  /* conv I   GEMM; NROT rotated weight images (the A1 window is a ring of 9 row slots) */
constexpr int K2 = 64, N2 = 32;            /* conv II  GEMM (+ one K=16 step against a constant-1 operand that adds the bias) */
constexpr int N3 = 16, KS3 = 12;           /* conv III GEMM per PAIR of output rows: K = 6 rows x 32 ch = 12 K-steps, N = 2 rows x 5 dx */
constexpr int B1_BYTES = N1 * K1 * 2, B2_BYTES = N2 * K2 * 2, B2B_BYTES = N2 * 16 * 2, B3_TILE = N3 * 16 * 2;
constexpr int B3_BYTES = KS3 * B3_TILE;
constexpr int W_B2 = NROT * B1_BYTES, W_B2B = W_B2 + B2_BYTES, W_B3 = W_B2B + B2B_BYTES;
constexpr int WIMG_BYTES = W_B3 + B3_BYTES;

/* TMEM columns (512 allocated).  D1/D2/A1/A2 are double-buffered by row parity, A3 is a ring of 8 conv II rows
 * (4 pair slots), D3 is double-buffered by set parity (one set per two output rows). */
/* (Measured and removed again, see profiles/r2_conv_pipeline_investigation.md and commit eb98414: a timing-only build with
 * triple-buffered D1 / A1 -- 0.1402 vs 0.1349 ms --, timing-only builds without single stages or waits, and SRCNN_E1IM, the
 * epilogue I groups doing im2col's operand rewrite as well -- bit-identical, 0.2246 vs 0.1334 ms.) */
constexpr uint32_t D1_COL = 0, D2_COL = 128, A2_COL = 192, A1_COL = 256, A3_COL = 352, D3_COL = 480;
constexpr uint32_t D1_STR = 64, D2_STR = 32, A1_STR = 48, A2_STR = 32, A3_SLOT = 16, D3_STR = 16;
/* conv II bias K-step: its A operand is the LAST K-step of the conv I operand, A1[b] columns 40..47 = k 80..95: k = 90, 91
 * are the constant 1.0 inputs (the bias sits at those k in the B image), every other k meets a zero weight there.  The
 * values under the zero weights are Y samples (finite), whichever row im2col has in flight, so the product is exactly 0. */
constexpr uint32_t ABIAS_OFF = 40;
constexpr int KBIAS = 10;                  /* k (within that K-step) of the first constant-1 element */

/* Y16 staging ring */
constexpr int CH_ROWS = 8, BOX_W = 144, NST = 4;
constexpr int CH_BYTES = CH_ROWS * BOX_W * 2;

/* warp roles: one control warpgroup + five 4-warp groups (TMEM lane quarter = warp & 3); the order of the groups made no
 * measurable difference on B200 */
/* SRCNN_ISSUE_SPLIT = 1: the conv I and conv II issuers are split by row parity (a second control warpgroup: warps 24 / 25 issue
 * the odd rows, 26 / 27 idle).  An issuer's per-row loop is serial -- observe the operand barrier, fence, elect, ~10 uniform-register
 * moves, the MMAs, commit, reconverge: ~220 clk per batch on an idle SM (tools/probe/mma_victim.cu), ~340 clk inside this kernel
 * (wait profile: the conv I issuer is outside its barrier wait 66 % of the time) -- and with one warp per layer every row of the
 * CTA passes through it. */
#ifndef SRCNN_ISSUE_SPLIT
#define SRCNN_ISSUE_SPLIT 0
#endif
constexpr int W_E3 = 0, W_E2 = 4, W_E1 = 8, W_IM2COL = 16, W_CTL = 20, W_CTL2 = 24, NWARPS = SRCNN_ISSUE_SPLIT ? 28 : 24;
constexpr int W_MMA1B = W_CTL2, W_MMA2B = W_CTL2 + 1;
/* SRCNN_CTLPERM: sub-partition (= warp index within the control warpgroup) of the TMA producer and the three issuers, one hex
 * digit each: 0x3210 = TMA 0, conv I 1, conv II 2, conv III 3 */
#ifndef SRCNN_CTLPERM
#define SRCNN_CTLPERM 0x3210
#endif
constexpr int W_TMA = W_CTL + (SRCNN_CTLPERM & 15), W_MMA1 = W_CTL + ((SRCNN_CTLPERM >> 4) & 15), W_MMA2 = W_CTL + ((SRCNN_CTLPERM >> 8) & 15),
              W_MMA3 = W_CTL + ((SRCNN_CTLPERM >> 12) & 15);

/* barrier indices.  M?_DONE[b] fires when the MMAs of layer ? for a row with parity b have completed: it both frees the
 * A operand region for the producer group and publishes the accumulator to the consumer group. */
/* SRCNN_MERGE_BAR: "operand written" and "accumulator drained" of conv I (and of conv II) arrive on ONE barrier of count 8
 * (4 producer warps + 4 drainer warps; the drainers pre-arrive once for the first row of each parity), so the issuing lane
 * -- the pipeline's critical thread -- probes one barrier per row instead of two. */
#ifndef SRCNN_MERGE_BAR
#define SRCNN_MERGE_BAR 1
#endif
enum { Y_FULL = 0, Y_EMPTY = NST, A1_FULL = 2 * NST, M1_DONE = A1_FULL + 2, D1_EMPTY_ = M1_DONE + 2,
       A2_FULL = D1_EMPTY_ + 2, M2_DONE = A2_FULL + 2, D2_EMPTY_ = M2_DONE + 2,
       A3_FULL = D2_EMPTY_ + 2, A3_EMPTY = A3_FULL + 4, M3_DONE = A3_EMPTY + 4, D3_EMPTY = M3_DONE + 2, A1_FREE = D3_EMPTY + 2,
       NBAR = A1_FREE + 2,
       D1_EMPTY = SRCNN_MERGE_BAR ? A1_FULL : D1_EMPTY_, D2_EMPTY = SRCNN_MERGE_BAR ? A2_FULL : D2_EMPTY_ };

/* dynamic shared memory layout (offsets from a 1024-byte aligned base) */
constexpr int S_Y = 0, S_W = NST * CH_BYTES, S_P = S_W + WIMG_BYTES, S_OUT = S_P + 2 * 10 * 128 * 4, S_BAR = S_OUT + 2 * 2 * 384,
              S_TMEM = S_BAR + NBAR * 8;
constexpr int SMEM_BYTES = S_TMEM + 16 + 1024;
static_assert(S_W % 128 == 0 && S_P % 16 == 0 && S_BAR % 8 == 0, "smem layout");

__constant__ float cBias3;
/* SRCNN_BIAS2_EPI = 1: the conv II bias is added by epilogue II in FP32 (a constant-bank operand of the add: no register, no
 * load) instead of by a fifth K-step of the conv II GEMM against a constant-1 operand -- one MMA issue and 16 tensor-pipe
 * clocks less per image row */
#ifndef SRCNN_BIAS2_EPI
#define SRCNN_BIAS2_EPI 0     /* measured (round 2): 0.1330 ms vs 0.1283 ms -- the 32 extra adds per row lengthen the conv II -> conv III
                                 hand-over by more than the K-step they save.  (Re-loading the drained accumulator with the bias
                                 by tcgen05.st instead: 0.151 ms, the store + its wait sit on the D2 release path.) */
#endif
__constant__ float cBias2[32];

/* ---------------------------------------------------------------- PTX helpers */
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

/* shared-memory matrix descriptor, K-major, no swizzle: core matrix = 8 rows x 16 B stored contiguously (128 B);
 * LBO = byte distance between the two core matrices along K, SBO = between 8-row groups along M/N */
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}
/* instruction descriptor: D = F32, A = B = F16, both K-major, M = 128 */
__device__ __forceinline__ constexpr uint32_t umma_idesc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | (8u << 24); }

/* The descriptor is handed over as two 32-bit words: only the low one (address and LBO fields) varies between K-steps
 * and weight images, the high one (SBO, descriptor version) is a compile-time constant per operand -- this keeps the
 * single issuing lane's address arithmetic 32-bit (it sits on the pipeline's critical path) */
__device__ __forceinline__ constexpr uint32_t umma_desc_hi(uint32_t sbo) { return ((sbo >> 4) & 0x3FFFu) | (1u << 14); }
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t bdesc_lo, uint32_t bdesc_hi, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n.reg .pred p;\n.reg .b64 bd;\nsetp.ne.b32 p, %5, 0;\nmov.b64 bd, {%2, %3};\n"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], bd, %4, p;\n}\n"
                 :: "r"(d_tmem), "r"(a_tmem), "r"(bdesc_lo), "r"(bdesc_hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}
/* SRCNN_WAITMODE: 0 = try_wait with a suspend-time hint (compiles to TRYWAIT / NANOSLEEP.SYNCS / PHASECHK / BRA), 1 = plain try_wait
 * loop (TRYWAIT / BRA), 2 = try_wait + nanosleep back-off */
#ifndef SRCNN_WAITMODE
#define SRCNN_WAITMODE 0
#endif
#ifndef SRCNN_WAITNS
#define SRCNN_WAITNS 32
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
#if SRCNN_WAITMODE == 0
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
                 "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" :: "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
#elif SRCNN_WAITMODE == 1
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" :: "r"(bar), "r"(parity) : "memory");
#else
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\nnanosleep.u32 %2;\nbra WAIT_%=;\nDONE_%=:\n}\n" :: "r"(bar), "r"(parity), "n"(SRCNN_WAITNS) : "memory");
#endif
}
/* latency-critical waits (the single-lane MMA issuers): poll instead of suspending -- a suspended warp resumes noticeably
 * later than the phase flips, and that wake-up latency sits on every operand / accumulator hand-over of the pipeline */
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\n"
                 "mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" :: "r"(bar), "r"(parity) : "memory");
}
/* uncritical waits (the TMA producer runs up to four 8-row chunks ahead): back off between probes */
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\nnanosleep.u32 256;\nbra WAIT_%=;\nDONE_%=:\n}\n" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(dst), "l"(map), "r"(x), "r"(y), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st()    { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld()    { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(N)); }
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(N)); }

__device__ __forceinline__ void tmem_st1(uint32_t addr, uint32_t r0)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" :: "r"(addr), "r"(r0) : "memory");
}
__device__ __forceinline__ void tmem_st2(uint32_t addr, uint32_t r0, uint32_t r1)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" :: "r"(addr), "r"(r0), "r"(r1) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t* r)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t addr, const uint32_t* r)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                    "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t addr, const uint32_t* r)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                 "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                 :: "r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                    "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
                    "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
                    "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t* r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(addr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t* r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(addr) : "memory");
}
/* two FP32 -> packed FP16x2 with ReLU in one instruction; `lo` lands in the low half (even K index) */
__device__ __forceinline__ uint32_t cvt_relu_h2(uint32_t lo, uint32_t hi)
{
    uint32_t d;
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
    return d;
}
/* warp-uniform code, one elected lane issues (keeps descriptors / addresses in uniform registers) */
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
    return pred != 0;
}
/* one lane per warp signals for the whole warp (barrier counts are 4 = warps per group) */
/* SRCNN_ARRIVE_ALL = 1: every lane arrives for itself (barrier counts x 32): no __syncwarp and no predicated single-lane
 * arrive on the hand-over path */
#ifndef SRCNN_ARRIVE_ALL
#define SRCNN_ARRIVE_ALL 1      /* measured (round 2, three interleaved runs each): 0.1333-0.1337 vs 0.1341-0.1344 ms */
#endif
constexpr uint32_t ARV = SRCNN_ARRIVE_ALL ? 32u : 1u;      /* arrivals per warp */
__device__ __forceinline__ void warp_arrive(uint32_t bar, int lane)
{
#if SRCNN_ARRIVE_ALL
    mbar_arrive(bar);
#else
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
#endif
}

/* ---------------------------------------------------------------- the kernel */
/* Per-site choice of the wait flavour (development switch, A/B measured): bit set = poll (mbarrier.test_wait) instead of
 * the suspending try_wait.  Sites: 0 MMA1.A1_FULL  1 MMA1.D1_EMPTY  2 im2col.M1_DONE  3 epilogueI.M1_DONE  4 MMA2.A2_FULL
 * 5 MMA2.D2_EMPTY  6 epilogueI.M2_DONE  7 epilogueII.M2_DONE  8 MMA3.A3_FULL  9 MMA3.D3_EMPTY  10 gather.M3_DONE */
#ifndef SRCNN_SPINMASK
#define SRCNN_SPINMASK 0
#endif
#ifdef SRCNN_PROF
/* development build: every barrier wait accumulates its duration per call site; CTA 0 dumps (per warp) the totals */
#define WAITP(id, bar, par) do { const long long t0_ = clock64(); mbar_wait(bar, par); prof_acc[id] += clock64() - t0_; } while (0)
#define WAITS(site, id, bar, par) do { const long long t0_ = clock64(); if ((SRCNN_SPINMASK >> (site)) & 1) mbar_wait_spin(bar, par); else mbar_wait(bar, par); prof_acc[id] += clock64() - t0_; } while (0)
#else
#define WAITP(id, bar, par) mbar_wait(bar, par)
#define WAITS(site, id, bar, par) do { if ((SRCNN_SPINMASK >> (site)) & 1) mbar_wait_spin(bar, par); else mbar_wait(bar, par); } while (0)
#endif
/* SRCNN_TL (development build): CTA 0 records clock64() of the pipeline events of TL_N consecutive rows starting at row TL_R0
 * (CTA-global row counter) into p.prof[row][32]; tools/conv_timeline.py reads the dump and attributes every tensor-pipe gap */
#ifdef SRCNN_TL
#ifndef SRCNN_TL_CTA
#define SRCNN_TL_CTA 0
#endif
constexpr int TL_R0 = 96, TL_N = 128, TL_W = 128;      /* per row: 32 events x 4 warps of a group */
#define TL(ev, rowidx, cond) do { const int r_ = (int)(rowidx) - TL_R0; \
    if (blockIdx.x == SRCNN_TL_CTA && (threadIdx.x & 31) == 0 && r_ >= 0 && r_ < TL_N) p.prof[r_ * TL_W + (ev) * 4 + ((threadIdx.x >> 5) & 3)] = clock64(); } while (0)
#define TLI(ev, rowidx) do { const int r_ = (int)(rowidx) - TL_R0; \
    if (blockIdx.x == SRCNN_TL_CTA && r_ >= 0 && r_ < TL_N) p.prof[r_ * TL_W + (ev) * 4] = clock64(); } while (0)
#else
#define TL(ev, rowidx, cond) ((void)0)
#define TLI(ev, rowidx) ((void)0)
#endif
/* SRCNN_GWAIT = 1: only ONE warp of a four-warp group polls an mbarrier; its three siblings block in a named barrier until the
 * poller joins them.  mbarrier.try_wait does not really suspend a warp on this part: with or without a suspend-time hint the retry
 * loop (TRYWAIT / NANOSLEEP.SYNCS / PHASECHK / BRA) spins, and ncu's source page attributes ~45 % of the kernel's issued warp
 * instructions to those loops.  bar.sync waits cost no issue slots.  Measured: 0.1446 vs 0.1282 ms -- the extra hop costs more than
 * the freed issue slots give (profiles/r2_conv_pipeline_investigation.md).  SRCNN_LEADS: which warp of each group polls, one hex digit per
 * group (gather, epilogue II, epilogue I odd, epilogue I even, im2col) -- spread over the sub-partitions next to the issuers. */
#ifndef SRCNN_GWAIT
#define SRCNN_GWAIT 0
#endif
/* SRCNN_PRESLEEP = ns: groups whose next wait is known to be long (epilogue I after its hand-over: the next row of its parity is
 * two row periods away; gather after a pair; the TMA producer) sleep once before they start polling -- less polling traffic on
 * the barrier unit without a longer wake-up, as long as the sleep is shorter than the wait */
#ifndef SRCNN_PRESLEEP
#define SRCNN_PRESLEEP 0
#endif
__device__ __forceinline__ void presleep() { if (SRCNN_PRESLEEP) asm volatile("nanosleep.u32 %0;" :: "n"(SRCNN_PRESLEEP)); }
#ifndef SRCNN_LEADS
#define SRCNN_LEADS 0x10320
#endif
constexpr int NB_IM = 8, NB_E1 = 9, NB_E2 = 11, NB_G = 12;          /* named barriers of the group waits (E1: 9 + parity) */
__device__ __forceinline__ void named_sync(uint32_t id) { asm volatile("bar.sync %0, 128;" :: "r"(id) : "memory"); }
#define GWAITS(lead, nb, site, id, bar, par) do { if (!SRCNN_GWAIT || (lead)) WAITS(site, id, bar, par); if (SRCNN_GWAIT) named_sync(nb); } while (0)
#define GWAITP(lead, nb, id, bar, par) do { if (!SRCNN_GWAIT || (lead)) WAITP(id, bar, par); if (SRCNN_GWAIT) named_sync(nb); } while (0)
/* 1: im2col rewrites only the two ring slots (10 TMEM columns) that changed since the same-parity row two rows earlier */
#ifndef SRCNN_A1_PARTIAL
#define SRCNN_A1_PARTIAL 1
#endif

struct TcParams {
    int pitch, y16_rows;
    int dw, dh;
    float* out;                 /* FP32 Y', rows [row0,row1) x dw, or NULL */
    const float *cb, *cr, *al;  /* fused colour-out (rgb != NULL): chroma / alpha planes, u8 outputs; all at row row0 */
    unsigned char *rgb, *conv8;
    int d;
    int row0, row1;
    int bands, ppb;             /* column bands; row pairs per band = ceil((row1-row0) / 2) */
    int nframes;                /* same-geometry frames in this launch: frame f's conv-input plane starts at plane row
                                   f * y16_rows, its Y' at out + f * out_fstride */
    size_t out_fstride;
    const unsigned char* wimg;  /* B1 x 9 rotations | B2 | B2-bias | B3 operand images */
#if defined(SRCNN_PROF) || defined(SRCNN_TL)
    long long* prof;
#endif
};

/* Work partition.  The (band, row pair) grid is walked band-major and cut into gridDim.x contiguous spans of equal
 * length, one per CTA; a span that crosses a band boundary splits into units.  A work unit = one 128-column band x
 * the rows [ra, rb); it streams the conv II rows q = ra-2 .. ra-3+nrows, nrows even (output rows are produced in
 * pairs; an odd tail computes one discarded extra row).  Long contiguous units keep the per-unit cost (4 halo rows,
 * 8 window-fill rows, pipeline fill and drain) to about two per CTA. */
struct Unit { int frame, band, ra, rb, npairs, nrows; };
struct Span { int s, e, nunits; };
__device__ __forceinline__ Span span_of(const TcParams& p)
{
    const long long T = (long long)p.nframes * p.bands * p.ppb;      /* (frame, band) pairs are walked like bands of one image */
    Span x;
    x.s = (int)(T * blockIdx.x / gridDim.x);
    x.e = (int)(T * (blockIdx.x + 1) / gridDim.x);
    x.nunits = x.e > x.s ? (x.e - 1) / p.ppb - x.s / p.ppb + 1 : 0;
    return x;
}
__device__ __forceinline__ Unit unit_of(const TcParams& p, const Span& sp, int k)
{
    Unit x;
    const int gb = sp.s / p.ppb + k;                  /* band index over all frames */
    x.frame = gb / p.bands;
    x.band = gb - x.frame * p.bands;
    const int p0 = k == 0 ? sp.s - gb * p.ppb : 0;
    const int p1 = min(p.ppb, sp.e - gb * p.ppb);
    x.ra = p.row0 + 2 * p0;
    x.rb = min(p.row0 + 2 * p1, p.row1);
    x.npairs = p1 - p0;
    x.nrows = 2 * x.npairs + 4;
    return x;
}
__device__ __forceinline__ int total_rows(const TcParams& p, const Span& sp)
{
    int total = 0;
    for (int k = 0; k < sp.nunits; k++) total += unit_of(p, sp, k).nrows;
    return total;
}


/* FUSE: the gather group's store stage also does colour-out (optional, measured slower); the default instantiation
 * carries none of that code */
template <bool FUSE>
__global__ void __launch_bounds__(NWARPS * 32, 1)
conv_tc_kernel(const TcParams p, const __grid_constant__ CUtensorMap ymap)
{
    extern __shared__ unsigned char smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* const sgen = smem_raw + (sbase - smem_u32(smem_raw));
    const uint32_t sY = sbase + S_Y, sW = sbase + S_W, bar0 = sbase + S_BAR;
    unsigned char* const sOut = sgen + S_OUT;                          /* [set parity][row of the pair][124 px x 3 B] */

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef SRCNN_PROF
    long long prof_acc[2] = { 0, 0 };
    const long long prof_t0 = clock64();
#endif
    auto BAR = [&](int i) -> uint32_t { return bar0 + 8u * (uint32_t)i; };

    /* one-time setup: weights -> smem (generic proxy, then made visible to the async proxy), barriers, TMEM */
    for (int i = tid; i < WIMG_BYTES / 16; i += NWARPS * 32)
        reinterpret_cast<uint4*>(sgen + S_W)[i] = __ldg(reinterpret_cast<const uint4*>(p.wimg) + i);
    if (tid == 0) {
        for (int i = 0; i < NST; i++) { mbar_init(BAR(Y_FULL + i), 1); mbar_init(BAR(Y_EMPTY + i), 4 * ARV); }
        for (int b = 0; b < 2; b++) {
            /* (SRCNN_PI: the operand-written half of the merged barriers is a bar.sync inside the producer group; only the
             * four drainer warps arrive here) */
            constexpr int e1w = SRCNN_E1_SPLIT ? 8 : 4;               /* epilogue I warps working on one row */
            mbar_init(BAR(A1_FULL + b), ARV * ((SRCNN_MERGE_BAR && !SRCNN_PI) ? 4 + e1w : (SRCNN_PI ? e1w : 4))); mbar_init(BAR(M1_DONE + b), 1);
            mbar_init(BAR(D1_EMPTY_ + b), ARV * e1w);
            mbar_init(BAR(A2_FULL + b), ARV * ((SRCNN_MERGE_BAR && !SRCNN_PI) ? e1w + 4 : (SRCNN_PI ? 4 : e1w))); mbar_init(BAR(M2_DONE + b), 1);
            mbar_init(BAR(D2_EMPTY_ + b), ARV * 4);
        }
        for (int j = 0; j < 4; j++) { mbar_init(BAR(A3_FULL + j), 4 * ARV); mbar_init(BAR(A3_EMPTY + j), 1); }
        for (int b = 0; b < 2; b++) { mbar_init(BAR(M3_DONE + b), 1); mbar_init(BAR(D3_EMPTY + b), 4 * ARV); mbar_init(BAR(A1_FREE + b), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_MMA2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(sbase + S_TMEM) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    /* everything above touched only the (constant) weight image; from here on the conv-input plane of the front end is read
     * and Y' is written: wait for the preceding kernel if this is a programmatic dependent launch, then let the colour-out
     * kernel be scheduled as soon as CTAs of this grid retire */
    pdl_wait();
    pdl_launch_dependents();
#if SRCNN_TMEM0
    /* The CTA allocates all 512 columns and is alone on its SM (144 KB of shared memory), so the allocation starts at column 0,
     * lane 0: address 0.  With the base a compile-time constant every TMEM address of the issuing lanes is an immediate -- the
     * compiler materialises them in uniform registers directly instead of moving them over from vector registers (R2UR) before
     * every tcgen05.mma: the single-lane issue loops, ~45 serial instructions per image row each, are the pipeline's pacemakers */
    if (*reinterpret_cast<volatile uint32_t*>(sgen + S_TMEM) != 0u) __trap();
    constexpr uint32_t tmem = 0u;
#else
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sgen + S_TMEM);
#endif
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);     /* this warp's TMEM lane quarter */
    const Span span = span_of(p);

    if (warp >= W_CTL) {
        /* ============================ control warpgroup: TMA producer + three MMA issuers ==================== */
        reg_dec<48>();                      /* the setmaxnreg pool is what the dec's release: 4096 + 2048 + 2048 regs */
        if (SRCNN_ONE_LANE && !SRCNN_PI && warp != W_TMA && lane != 0) return;      /* (exited threads do not take part in the final barrier) */
        if (warp == W_TMA) {
            /* Y16 rows of the current unit, 8-row chunks into a 4-stage ring (rows past the plane are zero-filled) */
            uint32_t ck = 0;
            for (int u = 0; u < span.nunits; u++) {
                const Unit un = unit_of(p, span, u);
                const int x0 = (un.band * kBandCols + 2) & ~7;            /* 16 B aligned box origin */
                const int yr0 = un.frame * p.y16_rows + un.ra - p.row0;   /* plane row of image row ra-6 (of this frame) */
                const int nch = (un.nrows + 8 + CH_ROWS - 1) / CH_ROWS;
                for (int j = 0; j < nch; j++, ck++) {
                    const uint32_t s = ck % NST, k = ck / NST;
#if (SRCNN_GWAIT || SRCNN_PRESLEEP) && !defined(SRCNN_PROF)
                    mbar_wait_sleep(BAR(Y_EMPTY + s), (k & 1) ^ 1);
#else
                    WAITP(0, BAR(Y_EMPTY + s), (k & 1) ^ 1);
#endif
                    if (elect_one()) {
                        mbar_expect_tx(BAR(Y_FULL + s), CH_BYTES);
                        tma_load_2d(sY + s * CH_BYTES, &ymap, x0, yr0 + j * CH_ROWS, BAR(Y_FULL + s));
                    }
                    __syncwarp();
                }
            }
        } else if (SRCNN_PI) {
            /* (the producer groups issue their own MMAs; these three warps only wait for the end of the kernel) */
        } else if (warp >= W_CTL2 + 2) {
            /* (idle: completes the second control warpgroup) */
        } else if (warp == W_MMA1 || warp == W_MMA1B) {
            const uint32_t myb = warp == W_MMA1B;
            ISSUER_BEGIN
            /* conv I: the A1 window is a ring of 9 row slots, so the weight image rotates with the row (R = n mod 9) */
            const uint32_t db0 = (uint32_t)umma_desc(sW, 128, (K1 / 8) * 128);
            constexpr uint32_t dh1 = umma_desc_hi((K1 / 8) * 128);
            uint32_t pr = 0;                                               /* global row-pair counter */
            for (int u = 0; u < span.nunits; u++) {
                const Unit un = unit_of(p, span, u);
                uint32_t R = (uint32_t)(un.ra % 9);   /* weight-image rotation of conv II row q: (q + 2) mod 9 */
                #pragma unroll 1
                for (int n = 0; n < un.nrows; n += 2, pr++) {
                    const uint32_t ph = pr & 1;
                    #pragma unroll
                    for (uint32_t b = 0; b < 2; b++) {
                        if (SRCNN_ISSUE_SPLIT && b != myb) { R = (R == NROT - 1) ? 0 : R + 1; continue; }
                        if (ISSUE_LANE()) TLI(22, 2 * pr + b);
                        WAITS(0, 0, BAR(A1_FULL + b), ph);
                        if (!SRCNN_MERGE_BAR) WAITS(1, 1, BAR(D1_EMPTY + b), ph ^ 1);
                        tc_fence_after();
                        if (ISSUE_LANE()) {
                            TLI(3, 2 * pr + b);
                            const uint32_t db = db0 + R * (B1_BYTES >> 4);
#if SRCNN_A1_EARLY
                            /* window rows 0 and 1 of this output row sit in ring slots R and R+1 (mod 9), columns 5R .. 5R+9
                             * (mod 45) = K-steps s0 .. s0+2 (mod 6) with s0 = floor(5R / 8): 0 0 1 1 2 3 3 4 5 for R = 0 .. 8 */
                            uint32_t s = (uint32_t)((0x543321100ull >> (4 * R)) & 7u);
                            #pragma unroll
                            for (int i = 0; i < K1 / 16; i++) {
                                umma_ts(tmem + D1_COL + b * D1_STR, tmem + A1_COL + b * A1_STR + 8 * s, db + 16 * s, dh1, umma_idesc(N1), i > 0);
                                if (i == (R == 3 ? 2 : 1)) umma_commit(BAR(A1_FREE + b));      /* (only R = 3 straddles three K-steps) */
                                s = (s == K1 / 16 - 1) ? 0 : s + 1;
                            }
#else
                            #pragma unroll
                            for (int s = 0; s < K1 / 16; s++) {
                                umma_ts(tmem + D1_COL + b * D1_STR, tmem + A1_COL + b * A1_STR + 8 * s, db + 16 * s, dh1, umma_idesc(N1), s > 0);
                            }
#endif
                            umma_commit(BAR(M1_DONE + b));
                            TLI(4, 2 * pr + b);
                        }
                        ISSUE_SYNC();
#ifdef SRCNN_TL
                        if (ISSUE_LANE()) TLI(25, 2 * pr + b);
#endif
                        R = (R == NROT - 1) ? 0 : R + 1;
                    }
                }
            }
            ISSUER_END
        } else if (warp == W_MMA2 || warp == W_MMA2B) {
            const uint32_t myb = warp == W_MMA2B;
            ISSUER_BEGIN
            const uint32_t db = (uint32_t)umma_desc(sW + W_B2, 128, (K2 / 8) * 128);
            const uint32_t dbb = (uint32_t)umma_desc(sW + W_B2B, 128, 256);
            constexpr uint32_t dh2 = umma_desc_hi((K2 / 8) * 128), dhb = umma_desc_hi(256);
            const int total = total_rows(p, span);
            #pragma unroll 1
            for (int g = 0; g < total; g += 2) {
                const uint32_t ph = (g >> 1) & 1;
                #pragma unroll
                for (uint32_t b = 0; b < 2; b++) {
                    if (SRCNN_ISSUE_SPLIT && b != myb) continue;
                    if (ISSUE_LANE()) TLI(23, g + b);
                    WAITS(4, 0, BAR(A2_FULL + b), ph);
                    if (!SRCNN_MERGE_BAR) WAITS(5, 1, BAR(D2_EMPTY + b), ph ^ 1);
                    tc_fence_after();
                    if (ISSUE_LANE()) {
                        TLI(9, g + b);
                        #pragma unroll
                        for (int s = 0; s < K2 / 16; s++)
                            umma_ts(tmem + D2_COL + b * D2_STR, tmem + A2_COL + b * A2_STR + 8 * s, db + 16 * s, dh2, umma_idesc(N2), s > 0);
                        if (!SRCNN_BIAS2_EPI)
                            umma_ts(tmem + D2_COL + b * D2_STR, tmem + A1_COL + b * A1_STR + ABIAS_OFF, dbb, dhb, umma_idesc(N2), 1);     /* + bias */
                        umma_commit(BAR(M2_DONE + b));
                        TLI(10, g + b);
                    }
                    ISSUE_SYNC();
                }
            }
            ISSUER_END
        } else {
            ISSUER_BEGIN
            /* conv III, dy on the tensor core: one set per pair of output rows (r0, r0+1):
             *   D3[px][5*o + dx] = sum_{rho<6} sum_ch A3[row clamp(r0-2+rho)][px][ch] * W3[ch][dy = rho-o-2][dx]
             * the row clamp of the reference (src/libsrcnn.cpp:461-491) is a choice of ring slot per K-step */
            const uint32_t db3 = (uint32_t)umma_desc(sW + W_B3, 128, 256);
            constexpr uint32_t dh3 = umma_desc_hi(256);
            uint32_t kg0 = 0, sg = 0;                                      /* global pair / set counters */
            for (int u = 0; u < span.nunits; u++) {
                const Unit un = unit_of(p, span, u);
                #pragma unroll 1
                for (int pp = 0; pp < un.npairs; pp++, sg++) {
                    const uint32_t kn = kg0 + (uint32_t)pp + 2;           /* newest pair this set reads */
                    if (pp == 0) {
                        /* the first set of a unit also reads the unit's first two pairs.  Their completion is implied by the
                         * newest pair's (same four producer warps, in order), but every completed phase of a barrier gets its own
                         * wait: no phase is ever skipped (compute-sanitizer synccheck: "missing wait") */
                        WAITS(8, 0, BAR(A3_FULL) + 8u * (kg0 & 3), (kg0 >> 2) & 1);
                        WAITS(8, 0, BAR(A3_FULL) + 8u * ((kg0 + 1) & 3), ((kg0 + 1) >> 2) & 1);
                    }
                    if (ISSUE_LANE()) TLI(24, 2 * sg + 4);
                    WAITS(8, 0, BAR(A3_FULL) + 8u * (kn & 3), (kn >> 2) & 1);
                    if (ISSUE_LANE()) TLI(14, 2 * sg + 4);
                    WAITS(9, 1, BAR(D3_EMPTY + (sg & 1)), ((sg >> 1) & 1) ^ 1);
                    tc_fence_after();
                    if (ISSUE_LANE()) {
                        TLI(15, 2 * sg + 4);
                        const int r0 = un.ra + 2 * pp;
                        const uint32_t d3 = tmem + D3_COL + (sg & 1) * D3_STR;
                        if (r0 >= 2 && r0 + 3 < p.dh) {
                            /* interior set (all but the first and last of an image): the six rows are three consecutive
                             * pair slots of the ring, no clamping -- the short instruction path matters because this single
                             * lane issues 12 MMAs per set on the pipeline's critical loop */
                            const uint32_t kp = kg0 + (uint32_t)pp;
                            #pragma unroll
                            for (int j = 0; j < 3; j++) {
                                const uint32_t pa = tmem + A3_COL + ((kp + j) & 3u) * (2 * A3_SLOT);
                                #pragma unroll
                                for (int h = 0; h < 4; h++) {
                                    umma_ts(d3, pa + 8 * h, db3 + (uint32_t)((4 * j + h) * (B3_TILE >> 4)), dh3, umma_idesc(N3), (j | h) > 0);
                                }
                                TLI(26 + j, 2 * sg + 4);
                            }
                        } else {
                            #pragma unroll
                            for (int rho = 0; rho < 6; rho++) {
                                int q = r0 - 2 + rho;
                                q = q < 0 ? 0 : (q > p.dh - 1 ? p.dh - 1 : q);
                                const uint32_t slot = (2u * kg0 + (uint32_t)(q - (un.ra - 2))) & 7u;
                                const uint32_t a = tmem + A3_COL + slot * A3_SLOT;
                                umma_ts(d3, a,     db3 + (uint32_t)((2 * rho) * (B3_TILE >> 4)),     dh3, umma_idesc(N3), rho > 0);
                                umma_ts(d3, a + 8, db3 + (uint32_t)((2 * rho + 1) * (B3_TILE >> 4)), dh3, umma_idesc(N3), 1);
                            }
                        }
                        umma_commit(BAR(M3_DONE + (sg & 1)));
                        /* pair kg0+pp is dead after this set; the last set of a unit retires the remaining two */
                        umma_commit(BAR(A3_EMPTY) + 8u * ((kg0 + (uint32_t)pp) & 3));
                        if (pp == un.npairs - 1) {
                            umma_commit(BAR(A3_EMPTY) + 8u * ((kg0 + (uint32_t)pp + 1) & 3));
                            umma_commit(BAR(A3_EMPTY) + 8u * ((kg0 + (uint32_t)pp + 2) & 3));
                        }
                        TLI(16, 2 * sg + 4);
                    }
                    ISSUE_SYNC();
                }
                kg0 += (uint32_t)un.npairs + 2;
            }
            ISSUER_END
        }
    } else if (warp >= W_IM2COL && warp < W_IM2COL + 4) {
        /* ===================== im2col: smem Y16 ring -> 9-slot register ring -> A1 in TMEM ==================== */
        reg_inc<96>();
        const int t = (warp - W_IM2COL) * 32 + lane;
        const bool lead = (warp & 3) == (SRCNN_LEADS & 15);
        uint32_t it = 0, ck = 0;
#if SRCNN_PI
        const uint32_t db1_0 = (uint32_t)umma_desc(sW, 128, (K1 / 8) * 128);
        constexpr uint32_t dh1 = umma_desc_hi((K1 / 8) * 128);
#endif
        for (int u = 0; u < span.nunits; u++) {
            const Unit un = unit_of(p, span, u);
            const int hw = ((un.band * kBandCols + 2) & 7) + t;       /* half offset of this lane's window in the box */
            const uint32_t woff = sY + (uint32_t)(hw >> 1) * 4u, sh = (uint32_t)(hw & 1) * 16u;
            const int last_rho = un.nrows + 7;
            uint32_t W[9][5];                                         /* slot s holds Y row (rho mod 9 == s), two halves per word */
            auto fetch = [&](int rho, uint32_t (&w5)[5]) {
                const uint32_t c = ck + (uint32_t)(rho / CH_ROWS), s = c % NST;
                if ((rho % CH_ROWS) == 0) GWAITP(lead, NB_IM, 0, BAR(Y_FULL + s), (c / NST) & 1);
                const uint32_t a = woff + s * CH_BYTES + (uint32_t)(rho % CH_ROWS) * (BOX_W * 2);
                uint32_t x[6];
                #pragma unroll
                for (int j = 0; j < 6; j++) asm volatile("ld.shared.b32 %0, [%1];" : "=r"(x[j]) : "r"(a + 4u * j));
                #pragma unroll
                for (int j = 0; j < 5; j++) w5[j] = __funnelshift_r(x[j], x[j + 1], sh);
                if ((rho % CH_ROWS) == CH_ROWS - 1 || rho == last_rho) warp_arrive(BAR(Y_EMPTY + s), lane);
            };
            /* Y row y always lives in slot (y + 6) mod 9, i.e. the k position of a tap depends on the absolute image row only
             * (so the FP32 accumulation order -- and the result, bit for bit -- does not depend on how rows are cut into
             * units or strips).  Row n (q = ra-2+n): the new Y row rho = n+8 lands in slot (ra+n+8) mod 9, window row i sits
             * in slot (ra+n+i) mod 9 and the MMA issuer picks the weight image rotated by R = (ra+n) mod 9.  The register ->
             * TMEM column map never moves.  The first 8 passes (n < 0) only fill the window. */
            auto row = [&](auto RC, int n) {
                constexpr int R = decltype(RC)::value;
                fetch(n + 8, W[(R + 8) % 9]);
                if (n < 0) return;
                const uint32_t b = it & 1, ph = (it >> 1) & 1;
                TL(0, it, t == 0);
#if SRCNN_A1_EARLY && SRCNN_A1_PARTIAL && !SRCNN_PI
                /* partial rewrite: only the K-steps that read the two oldest window rows of row it-2 must have run */
                WAITS(2, 1, BAR(A1_FREE) + 8u * b, ph ^ 1);
                if (n < 2) WAITS(2, 1, BAR(M1_DONE) + 8u * b, ph ^ 1);     /* whole operand rewritten: all of row it-2's MMAs */
#else
                GWAITS(lead, NB_IM, 2, 1, BAR(M1_DONE) + 8u * b, ph ^ 1);         /* MMA1 of row it-2 has consumed A1[b] */
#endif
                tc_fence_after();
                TL(1, it, t == 0);
                const uint32_t dst = lane_base + A1_COL + b * A1_STR;
                if (!SRCNN_A1_PARTIAL || n < 2) {
                    /* first two rows of a unit: the whole operand (both parity buffers still hold another unit's rows) */
                    uint32_t r[16];
                    #pragma unroll
                    for (int m = 0; m < 16; m++) r[m] = W[m / 5][m % 5];
                    tmem_st16(dst, r);
                    #pragma unroll
                    for (int m = 0; m < 16; m++) r[m] = W[(m + 16) / 5][(m + 16) % 5];
                    tmem_st16(dst + 16, r);
                    #pragma unroll
                    for (int m = 0; m < 13; m++) r[m] = W[(m + 32) / 5][(m + 32) % 5];
                    r[13] = 0x3C003C00u;                          /* k = 90, 91: constant 1.0 (bias inputs) */
                    r[14] = 0u; r[15] = 0u;
                    tmem_st16(dst + 32, r);
                } else {
                    /* A1[b] still holds the operand of row n-2, and the slot -> column map never moves: only the two ring
                     * slots refilled since then (Y rows n+7 and n+8) differ -- 10 of the 48 columns */
                    constexpr int SA = (R + 7) % 9, SB = (R + 8) % 9;
                    const uint32_t (&wa)[5] = W[SA];
                    const uint32_t (&wb)[5] = W[SB];
                    tmem_st4(dst + 5 * SA, wa[0], wa[1], wa[2], wa[3]);
                    if (SB == SA + 1) {
                        tmem_st2(dst + 5 * SA + 4, wa[4], wb[0]);
                        tmem_st4(dst + 5 * SB + 1, wb[1], wb[2], wb[3], wb[4]);
                    } else {
                        tmem_st1(dst + 5 * SA + 4, wa[4]);
                        tmem_st4(dst + 5 * SB, wb[0], wb[1], wb[2], wb[3]);
                        tmem_st1(dst + 5 * SB + 4, wb[4]);
                    }
                }
                tmem_wait_st();
                tc_fence_before();
                TL(2, it, t == 0);
#if SRCNN_PI
                asm volatile("bar.sync 3, 128;" ::: "memory");       /* all four lane quarters of A1[b] are written */
                if (warp == W_IM2COL) {
                    /* conv I of this row: D1[b] must have been drained by epilogue I (row it-2), then six K-steps against the
                     * weight image rotated by R = (image row) mod 9 -- a compile-time constant here */
                    WAITS(1, 1, BAR(D1_EMPTY) + 8u * b, ph);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t db = db1_0 + (uint32_t)R * (B1_BYTES >> 4);
                        #pragma unroll
                        for (int s = 0; s < K1 / 16; s++)
                            umma_ts(tmem + D1_COL + b * D1_STR, tmem + A1_COL + b * A1_STR + 8 * s, db + 16 * s, dh1, umma_idesc(N1), s > 0);
                        umma_commit(BAR(M1_DONE) + 8u * b);
                    }
                    __syncwarp();
                }
#else
                warp_arrive(BAR(A1_FULL) + 8u * b, lane);
#endif
                it++;
            };
            int n = -8;
#define SRCNN_ROW(R) case R: row(std::integral_constant<int, R>{}, n); if (++n == un.nrows) break;
            switch ((un.ra + 1) % 9) {
                for (;;) {
                    SRCNN_ROW(0) SRCNN_ROW(1) SRCNN_ROW(2) SRCNN_ROW(3) SRCNN_ROW(4) SRCNN_ROW(5) SRCNN_ROW(6) SRCNN_ROW(7) SRCNN_ROW(8)
                }
            }
#undef SRCNN_ROW
            ck += (uint32_t)((un.nrows + 8 + CH_ROWS - 1) / CH_ROWS);
        }
    } else if (warp >= W_E1 && warp < W_E1 + 8) {
        /* ===================== epilogue I (two groups, one per row parity): D1 -> ReLU -> FP16 -> A2 ========== */
        reg_inc<88>();
#if SRCNN_E1_SPLIT && !SRCNN_PI
        {
            const uint32_t half = (uint32_t)(warp - W_E1) >> 2;       /* which 32 of the 64 conv I channels */
            const int total = total_rows(p, span);
            if (SRCNN_MERGE_BAR && total > 0) { warp_arrive(BAR(D1_EMPTY + 0), lane); warp_arrive(BAR(D1_EMPTY + 1), lane); }
            #pragma unroll 1
            for (int g = 0; g < total; g++) {
                const uint32_t b = (uint32_t)g & 1u, ph = ((uint32_t)g >> 1) & 1u;
                const uint32_t d1 = lane_base + D1_COL + b * D1_STR + 32u * half, a2 = lane_base + A2_COL + b * A2_STR + 16u * half;
                uint32_t v[32];
                WAITS(3, 0, BAR(M1_DONE) + 8u * b, ph);
                tc_fence_after();
                tmem_ld32(d1, v);
                tmem_wait_ld();
                tc_fence_before();
                warp_arrive(BAR(D1_EMPTY) + 8u * b, lane);
                #pragma unroll
                for (int m = 0; m < 16; m++) v[m] = cvt_relu_h2(v[2 * m], v[2 * m + 1]);
                WAITS(6, 1, BAR(M2_DONE) + 8u * b, ph ^ 1);          /* MMA2 of the previous row of this parity has consumed A2[b] */
                tc_fence_after();
                tmem_st16(a2, v);
                tmem_wait_st();
                tc_fence_before();
                warp_arrive(BAR(A2_FULL) + 8u * b, lane);
            }
        }
#else
        const uint32_t b = (uint32_t)(warp - W_E1) >> 2;
        const bool lead = (uint32_t)(warp & 3) == ((SRCNN_LEADS >> (4 + 4 * b)) & 15u);
        const int total = total_rows(p, span);
        const uint32_t d1 = lane_base + D1_COL + b * D1_STR;
        const uint32_t a2 = lane_base + A2_COL + b * A2_STR;
#if SRCNN_PI
        const uint32_t db2 = (uint32_t)umma_desc(sW + W_B2, 128, (K2 / 8) * 128);
        const uint32_t dbb2 = (uint32_t)umma_desc(sW + W_B2B, 128, 256);
        constexpr uint32_t dh2 = umma_desc_hi((K2 / 8) * 128), dhb2 = umma_desc_hi(256);
#endif
        if (SRCNN_MERGE_BAR && total > 0) warp_arrive(BAR(D1_EMPTY) + 8u * b, lane);     /* D1[b] starts out drained */
        #pragma unroll 1
        for (int g = 0; g < total; g += 2) {
            const uint32_t ph = (g >> 1) & 1;
            uint32_t v[64];
            GWAITS(lead, NB_E1 + b, 3, 0, BAR(M1_DONE) + 8u * b, ph);
            tc_fence_after();
            TL(5, g + b, ((warp & 3) | lane) == 0);
            tmem_ld32(d1, v);
            tmem_ld32(d1 + 32, v + 32);
            tmem_wait_ld();
            tc_fence_before();
            TL(6, g + b, ((warp & 3) | lane) == 0);
            warp_arrive(BAR(D1_EMPTY) + 8u * b, lane);
            #pragma unroll
            for (int m = 0; m < 32; m++) v[m] = cvt_relu_h2(v[2 * m], v[2 * m + 1]);
            GWAITS(lead, NB_E1 + b, 6, 1, BAR(M2_DONE) + 8u * b, ph ^ 1);              /* MMA2 of the previous row of this parity has consumed A2[b] */
            tc_fence_after();
            TL(7, g + b, ((warp & 3) | lane) == 0);
            tmem_st16(a2, v);
            tmem_st16(a2 + 16, v + 16);
            tmem_wait_st();
            tc_fence_before();
            TL(8, g + b, ((warp & 3) | lane) == 0);
#if SRCNN_PI
            if (b == 0) asm volatile("bar.sync 4, 128;" ::: "memory"); else asm volatile("bar.sync 5, 128;" ::: "memory");
            if (((warp - W_E1) & 3) == 0) {
                /* conv II of this row: D2[b] drained by epilogue II (previous row of this parity), four K-steps */
                WAITS(5, 1, BAR(D2_EMPTY) + 8u * b, ph);
                tc_fence_after();
                if (elect_one()) {
                    #pragma unroll
                    for (int s = 0; s < K2 / 16; s++)
                        umma_ts(tmem + D2_COL + b * D2_STR, tmem + A2_COL + b * A2_STR + 8 * s, db2 + 16 * s, dh2, umma_idesc(N2), s > 0);
                    if (!SRCNN_BIAS2_EPI)
                        umma_ts(tmem + D2_COL + b * D2_STR, tmem + A1_COL + b * A1_STR + ABIAS_OFF, dbb2, dhb2, umma_idesc(N2), 1);
                    umma_commit(BAR(M2_DONE) + 8u * b);
                }
                __syncwarp();
            }
#else
            warp_arrive(BAR(A2_FULL) + 8u * b, lane);
            presleep();
#endif
        }
#endif
    } else if (warp >= W_E2 && warp < W_E2 + 4) {
        /* ===================== epilogue II: D2 (bias included) -> ReLU -> FP16 -> A3 ring ===================== */
#if SRCNN_E2_DECOUPLE
        reg_inc<96>();                      /* holds a whole pair (32 + 32 registers); pool: 4096 + 2048 released = 3 x 2048 taken */
#elif SRCNN_PI
        /* (keeps its 80 registers: the group also issues conv III) */
#else
        reg_dec<64>();
#endif
        uint32_t kg = 0;                                               /* global pair counter */
        const bool lead = (warp & 3) == ((SRCNN_LEADS >> 12) & 15);
#if SRCNN_PI
        uint32_t sgc = 0;                                              /* global conv III set counter */
        const uint32_t db3 = (uint32_t)umma_desc(sW + W_B3, 128, 256);
        constexpr uint32_t dh3 = umma_desc_hi(256);
#endif
        if (SRCNN_MERGE_BAR && span.nunits > 0) { warp_arrive(BAR(D2_EMPTY + 0), lane); warp_arrive(BAR(D2_EMPTY + 1), lane); }   /* D2 starts out drained */
        for (int u = 0; u < span.nunits; u++) {
            const Unit un = unit_of(p, span, u);
            #pragma unroll 1
            for (int k = 0; k < un.npairs + 2; k++, kg++) {
                const uint32_t ph = kg & 1, js = kg & 3;
                const uint32_t a3 = lane_base + A3_COL + js * (2 * A3_SLOT);
#if SRCNN_E2_DECOUPLE
                /* both rows of the pair leave D2 (and release it to the conv II issuer) BEFORE the wait for the A3 ring slot:
                 * a late conv III set then stalls only this group, not conv II two rows ahead */
                uint32_t h[32];                                                /* the pair's two A3 slots are adjacent columns */
                #pragma unroll
                for (uint32_t b = 0; b < 2; b++) {
                    uint32_t v[32];
                    WAITS(7, 0, BAR(M2_DONE + b), ph);
                    tc_fence_after();
                    tmem_ld32(lane_base + D2_COL + b * D2_STR, v);
                    tmem_wait_ld();
                    tc_fence_before();
                    warp_arrive(BAR(D2_EMPTY + b), lane);
                    if (SRCNN_BIAS2_EPI) {
                        #pragma unroll
                        for (int m = 0; m < 32; m++) v[m] = __float_as_uint(__fadd_rn(__uint_as_float(v[m]), cBias2[m]));
                    }
                    #pragma unroll
                    for (int m = 0; m < 16; m++) h[16 * b + m] = cvt_relu_h2(v[2 * m], v[2 * m + 1]);
                }
                WAITP(1, BAR(A3_EMPTY) + 8u * js, ((kg >> 2) & 1) ^ 1);       /* every set that read this pair slot is done */
                tc_fence_after();
                static_assert(A3_SLOT == 16, "one 32-column store covers the pair");
                tmem_st32(a3, h);
#else
                #pragma unroll
                for (uint32_t b = 0; b < 2; b++) {
                    uint32_t v[32];
                    GWAITS(lead, NB_E2, 7, 0, BAR(M2_DONE + b), ph);
                    tc_fence_after();
                    TL(11, 2 * kg + b, ((warp & 3) | lane) == 0);
                    tmem_ld32(lane_base + D2_COL + b * D2_STR, v);
                    tmem_wait_ld();
                    tc_fence_before();
                    TL(12, 2 * kg + b, ((warp & 3) | lane) == 0);
                    warp_arrive(BAR(D2_EMPTY + b), lane);
                    if (SRCNN_BIAS2_EPI == 1) {
                        #pragma unroll
                        for (int m = 0; m < 32; m++) v[m] = __float_as_uint(__fadd_rn(__uint_as_float(v[m]), cBias2[m]));
                    }
                    #pragma unroll
                    for (int m = 0; m < 16; m++) v[m] = cvt_relu_h2(v[2 * m], v[2 * m + 1]);
                    if (b == 0) {
                        GWAITP(lead, NB_E2, 1, BAR(A3_EMPTY) + 8u * js, ((kg >> 2) & 1) ^ 1);   /* every set that read this pair slot is done */
                        tc_fence_after();
                    }
                    tmem_st16(a3 + b * A3_SLOT, v);
                    TL(13, 2 * kg + b, ((warp & 3) | lane) == 0);
                }
#endif
                tmem_wait_st();
                tc_fence_before();
                TL(21, 2 * kg, ((warp & 3) | lane) == 0);
#if SRCNN_PI
                asm volatile("bar.sync 6, 128;" ::: "memory");       /* the pair is complete in all four lane quarters */
                if (warp == W_E2 && k >= 2) {
                    /* conv III, dy on the tensor core: the set of output rows (r0, r0+1), r0 = ra + 2 (k-2), reads the conv II
                     * pairs k-2, k-1, k of this unit:
                     *   D3[px][5*o + dx] = sum_{rho<6} sum_ch A3[row clamp(r0-2+rho)][px][ch] * W3[ch][dy = rho-o-2][dx]
                     * the row clamp of the reference (src/libsrcnn.cpp:461-491) is a choice of ring slot per K-step */
                    const int pp = k - 2;
                    const uint32_t kg0 = kg - (uint32_t)k;                 /* pair counter at the start of this unit */
                    WAITS(9, 1, BAR(D3_EMPTY + (sgc & 1)), ((sgc >> 1) & 1) ^ 1);
                    tc_fence_after();
                    if (elect_one()) {
                        const int r0 = un.ra + 2 * pp;
                        const uint32_t d3 = tmem + D3_COL + (sgc & 1) * D3_STR;
                        if (r0 >= 2 && r0 + 3 < p.dh) {
                            const uint32_t kp = kg0 + (uint32_t)pp;
                            #pragma unroll
                            for (int j = 0; j < 3; j++) {
                                const uint32_t pa = tmem + A3_COL + ((kp + j) & 3u) * (2 * A3_SLOT);
                                #pragma unroll
                                for (int h = 0; h < 4; h++)
                                    umma_ts(d3, pa + 8 * h, db3 + (uint32_t)((4 * j + h) * (B3_TILE >> 4)), dh3, umma_idesc(N3), (j | h) > 0);
                            }
                        } else {
                            #pragma unroll
                            for (int rho = 0; rho < 6; rho++) {
                                int q = r0 - 2 + rho;
                                q = q < 0 ? 0 : (q > p.dh - 1 ? p.dh - 1 : q);
                                const uint32_t slot = (2u * kg0 + (uint32_t)(q - (un.ra - 2))) & 7u;
                                const uint32_t a = tmem + A3_COL + slot * A3_SLOT;
                                umma_ts(d3, a,     db3 + (uint32_t)((2 * rho) * (B3_TILE >> 4)),     dh3, umma_idesc(N3), rho > 0);
                                umma_ts(d3, a + 8, db3 + (uint32_t)((2 * rho + 1) * (B3_TILE >> 4)), dh3, umma_idesc(N3), 1);
                            }
                        }
                        umma_commit(BAR(M3_DONE + (sgc & 1)));
                        /* pair kg0+pp is dead after this set; the last set of a unit retires the remaining two */
                        umma_commit(BAR(A3_EMPTY) + 8u * ((kg0 + (uint32_t)pp) & 3));
                        if (pp == un.npairs - 1) {
                            umma_commit(BAR(A3_EMPTY) + 8u * ((kg0 + (uint32_t)pp + 1) & 3));
                            umma_commit(BAR(A3_EMPTY) + 8u * ((kg0 + (uint32_t)pp + 2) & 3));
                        }
                    }
                    __syncwarp();
                    sgc++;
                }
#else
                warp_arrive(BAR(A3_FULL) + 8u * js, lane);
#endif
            }
        }
    } else {
        /* ===================== conv III gather: dx across lanes, bias, clamp, store (two rows per set) ======== */
        reg_dec<64>();
        const int t = (warp - W_E3) * 32 + lane;
        const bool lead = (warp & 3) == ((SRCNN_LEADS >> 16) & 15);
        uint32_t sg = 0;
        for (int u = 0; u < span.nunits; u++) {
            const Unit un = unit_of(p, span, u);
            const int c0 = un.band * kBandCols - 2, c = c0 + t;
            const bool writer = (t >= 2 && t < 126 && c < p.dw);
            uint32_t pa[5];                     /* shared address of the dx term: lane of column clamp(c + dx) -- conv III
                                                   clamps its input plane -- in exchange buffer 0 */
            #pragma unroll
            for (int d = 0; d < 5; d++) {
                int cc = c + d - 2;
                cc = cc < 0 ? 0 : (cc > p.dw - 1 ? p.dw - 1 : cc);
                cc -= c0;
                pa[d] = sbase + S_P + 4u * (uint32_t)(d * 128 + (cc < 0 ? 0 : (cc > 127 ? 127 : cc)));
            }
            const uint32_t pst = sbase + S_P + 4u * (uint32_t)t;
            /* RGB (d = 3) rows are staged in shared memory and stored as words when the band's row segment is word-aligned */
            const bool rgb_staged = FUSE && p.rgb && p.d == 3 && (p.dw & 3) == 0 && ((uintptr_t)p.rgb & 3u) == 0;
            const int nvalid = min(kBandCols, p.dw - un.band * kBandCols), nwords = (nvalid * 3) >> 2;
            float ncb0 = 0.f, ncr0 = 0.f, nal0 = 0.f, ncb1 = 0.f, ncr1 = 0.f, nal1 = 0.f;
            auto load_chroma = [&](int r) {
                if (!(FUSE && writer && p.rgb)) return;
                const size_t ix = (size_t)(r - p.row0) * p.dw + c;
                ncb0 = __ldg(p.cb + ix); ncr0 = __ldg(p.cr + ix);
                if (p.d == 4) nal0 = __ldg(p.al + ix);
                if (r + 1 < un.rb) {
                    ncb1 = __ldg(p.cb + ix + p.dw); ncr1 = __ldg(p.cr + ix + p.dw);
                    if (p.d == 4) nal1 = __ldg(p.al + ix + p.dw);
                }
            };
            load_chroma(un.ra);
            float* odst = (!FUSE || p.out) ? p.out + ((size_t)un.frame * p.out_fstride + (size_t)(un.ra - p.row0) * p.dw + c) : nullptr;   /* two rows further per set */
            const size_t ostep = 2 * (size_t)p.dw;
            #pragma unroll 1
            for (int pp = 0; pp < un.npairs; pp++, sg++, odst += ostep) {
                uint32_t z[16];
                const int r = un.ra + 2 * pp;
                const bool two = r + 1 < un.rb;
                const size_t pix = FUSE ? (size_t)(r - p.row0) * p.dw + c : 0;
                /* chroma of this pair was requested one set ago; request the next pair's now (a full set period to land) */
                const float cb0 = ncb0, cr0 = ncr0, al0 = nal0, cb1 = ncb1, cr1 = ncr1, al1 = nal1;
                if (pp + 1 < un.npairs) load_chroma(r + 2);
                GWAITS(lead, NB_G, 10, 0, BAR(M3_DONE + (sg & 1)), (sg >> 1) & 1);
                tc_fence_after();
                TL(17, 2 * sg + 4, t == 0);
                tmem_ld16(lane_base + D3_COL + (sg & 1) * D3_STR, z);
                tmem_wait_ld();
                tc_fence_before();
                TL(18, 2 * sg + 4, t == 0);
                warp_arrive(BAR(D3_EMPTY + (sg & 1)), lane);
                const uint32_t poff = (sg & 1) * (10 * 128 * 4);            /* exchange buffer of this set */
                const uint32_t pdst = pst + poff;
                #pragma unroll
                for (int m = 0; m < 10; m++)
                    asm volatile("st.shared.b32 [%0], %1;" :: "r"(pdst + 512u * m), "r"(z[m]) : "memory");
                asm volatile("bar.sync 1, 128;" ::: "memory");
                TL(19, 2 * sg + 4, t == 0);
                if (writer) {
                    float o0 = cBias3, o1 = cBias3;
                    #pragma unroll
                    for (int d = 0; d < 5; d++) {
                        float a0, a1;
                        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a0) : "r"(pa[d] + poff) : "memory");
                        asm volatile("ld.shared.f32 %0, [%1+2560];" : "=f"(a1) : "r"(pa[d] + poff) : "memory");
                        o0 += a0; o1 += a1;
                    }
                    o0 = fminf(fmaxf(o0, 0.f), 255.f);
                    o1 = fminf(fmaxf(o1, 0.f), 255.f);
                    if (!FUSE || p.out) {
                        odst[0] = o0;
                        if (two) odst[p.dw] = o1;
                    }
                    if (FUSE && p.rgb) {
                        /* memcpy conv III -> Y, convertImgF32XtoImgU8, conv plane (src/libsrcnn.cpp:855-861, 889-915) */
                        auto emit = [&](float y, float cb, float cr, float al, size_t ix, unsigned char* stage) {
                            uint32_t R, G, B;
                            rgb_of(y, cb, cr, R, G, B);
                            if (p.d == 4) {
                                reinterpret_cast<uint32_t*>(p.rgb)[ix] = R | (G << 8) | (B << 16) | (to_u8(al) << 24);
                            } else if (stage) {
                                stage[0] = (unsigned char)R; stage[1] = (unsigned char)G; stage[2] = (unsigned char)B;
                            } else {
                                unsigned char* q = p.rgb + ix * 3;
                                q[0] = (unsigned char)R; q[1] = (unsigned char)G; q[2] = (unsigned char)B;
                            }
                            if (p.conv8) p.conv8[ix] = (unsigned char)__float2uint_rz(y);
                        };
                        unsigned char* stg = rgb_staged ? sOut + (sg & 1) * 768 + (t - 2) * 3 : nullptr;
                        emit(o0, cb0, cr0, al0, pix, stg);
                        if (two) emit(o1, cb1, cr1, al1, pix + p.dw, stg ? stg + 384 : nullptr);
                    }
                }
                if (rgb_staged) {
                    /* RGB rows leave as coalesced 32-bit words: 124 px x 3 B = 93 words per row and band */
                    asm volatile("bar.sync 2, 128;" ::: "memory");
                    if (t < nwords) {
                        const uint32_t* w0 = reinterpret_cast<const uint32_t*>(sOut + (sg & 1) * 768);
                        uint32_t* g0 = reinterpret_cast<uint32_t*>(p.rgb + ((size_t)(r - p.row0) * p.dw + (c0 + 2)) * 3);
                        g0[t] = w0[t];
                        if (two) reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(g0) + (size_t)p.dw * 3)[t] = w0[96 + t];
                    }
                }
                TL(20, 2 * sg + 4, t == 0);
                presleep();
            }
        }
    }

#ifdef SRCNN_PROF
    if (blockIdx.x == 0 && lane == 0) {
        p.prof[warp * 4 + 0] = clock64() - prof_t0; p.prof[warp * 4 + 1] = prof_acc[0]; p.prof[warp * 4 + 2] = prof_acc[1];
    }
#endif
    tc_fence_before();
    __syncthreads();
    if (warp == (SRCNN_ONE_LANE ? 0 : W_MMA2)) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}

/* element (n, k) of an N x K K-major no-swizzle operand image: 8x8 core matrices of 128 B, K-adjacent ones contiguous */
inline size_t canon_off(int n, int k, int K) { return (size_t)(n / 8) * ((K / 8) * 128) + (size_t)(k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

} // namespace

struct ConvTcState {
    unsigned char* d_wimg = nullptr;
    EncodeTiledFn encode = nullptr;
    std::string err;
};

int conv_tc_create(ConvTcState** out)
{
    ConvTcState* st = new ConvTcState;
    std::vector<unsigned char> img(WIMG_BYTES, 0);
    auto put = [&](size_t base, int n, int k, int K, float v) {
        const __half h = __float2half_rn(v);
        memcpy(img.data() + base + canon_off(n, k, K), &h, 2);
    };
    /* conv I, image R (rotation): the A1 operand is a ring of 9 row slots, k = 10*s + j for slot s; for an output row with
     * n mod 9 == R the window row i (Y row q-4+i) sits in slot (R + i) mod 9.  k = 90 / 91 carry the bias (hi / lo). */
    for (int R = 0; R < NROT; R++)
        for (int n = 0; n < N1; n++) {
            const size_t base = (size_t)R * B1_BYTES;
            for (int i = 0; i < 9; i++)
                for (int j = 0; j < 9; j++) put(base, n, 10 * ((R + i) % 9) + j, K1, SRCNN_W1[i * 9 + j][n]);
            const float b = SRCNN_B1[n];
            const float hi = __half2float(__float2half_rn(b));
            put(base, n, 90, K1, hi);
            put(base, n, 91, K1, b - hi);
        }
    for (int n = 0; n < N2; n++)
        for (int k = 0; k < K2; k++) put(W_B2, n, k, K2, SRCNN_W2[k][n]);
    /* conv II bias as a 32 x 16 operand for the constant-1 K-step (A = A1 columns 40..47): k = 10 hi part, k = 11 lo part */
    for (int n = 0; n < N2; n++) {
        const float b = SRCNN_B2[n];
        const float hi = __half2float(__float2half_rn(b));
        put(W_B2B, n, KBIAS, 16, hi);
        put(W_B2B, n, KBIAS + 1, 16, b - hi);
    }
    /* conv III, one 16 x 16 tile per K-step (rho, h): column n = 5*o + dx of output row r0+o takes conv II row r0-2+rho
     * with the tap dy+2 = rho - o (zero outside 0..4); k = channel 16*h + kk.  SRCNN_W3[tap = (dy+2)*5 + (dx+2)][ch]. */
    for (int rho = 0; rho < 6; rho++)
        for (int h = 0; h < 2; h++)
            for (int o = 0; o < 2; o++) {
                const int dyi = rho - o;
                if (dyi < 0 || dyi > 4) continue;
                for (int dx = 0; dx < 5; dx++)
                    for (int kk = 0; kk < 16; kk++)
                        put(W_B3 + (size_t)(2 * rho + h) * B3_TILE, 5 * o + dx, kk, 16, SRCNN_W3[dyi * 5 + dx][16 * h + kk]);
            }

    cudaError_t e = cudaMalloc(&st->d_wimg, WIMG_BYTES);
    if (e == cudaSuccess) e = cudaMemcpy(st->d_wimg, img.data(), WIMG_BYTES, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(cBias3, &SRCNN_B3, sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(cBias2, SRCNN_B2, sizeof(cBias2));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess) {
        void* fn = nullptr; cudaDriverEntryPointQueryResult q;
        e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (e == cudaSuccess && (q != cudaDriverEntryPointSuccess || !fn)) e = cudaErrorNotSupported;
        st->encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    if (e != cudaSuccess) { if (st->d_wimg) cudaFree(st->d_wimg); delete st; *out = nullptr; return (int)e; }
    *out = st;
    return 0;
}

void conv_tc_destroy(ConvTcState* st)
{
    if (!st) return;
    if (st->d_wimg) cudaFree(st->d_wimg);
    delete st;
}

const char* conv_tc_error(const ConvTcState* st) { return st ? st->err.c_str() : "no state"; }

int launch_conv_tc(ConvTcState* st, const __half* y16, Y16Geom g, unsigned dw, unsigned dh,
                   float* out, unsigned row0, unsigned row1, int num_sms, cudaStream_t s, const ConvTcColor* col,
                   unsigned nframes, size_t out_fstride)
{
    TcParams p;
    if (nframes == 0 || (nframes > 1 && col)) { st->err = "frame groups: no fused colour-out"; return 1; }
    p.nframes = (int)nframes; p.out_fstride = out_fstride;
    p.cb = col ? col->cb : nullptr; p.cr = col ? col->cr : nullptr; p.al = col ? col->al : nullptr;
    p.rgb = col ? col->rgb : nullptr; p.conv8 = col ? col->conv8 : nullptr; p.d = col ? (int)col->d : 3;
    /* the RGBA store is one 32-bit word per pixel */
    if (p.rgb && p.d == 4 && ((uintptr_t)p.rgb & 3u)) { st->err = "fused colour-out: RGBA output must be 4-byte aligned"; return 1; }
    if (!out && !p.rgb) { st->err = "no output requested"; return 1; }
    p.pitch = g.pitch; p.y16_rows = g.rows;
    p.dw = (int)dw; p.dh = (int)dh; p.out = out; p.row0 = (int)row0; p.row1 = (int)row1;
    p.bands = g.bands;
    const int rows = (int)(row1 - row0);
    p.ppb = (rows + 1) / 2;
    p.wimg = st->d_wimg;

    CUtensorMap map;
    const cuuint64_t gdim[2] = { (cuuint64_t)g.pitch, (cuuint64_t)g.rows * nframes };
    const cuuint64_t gstr[1] = { (cuuint64_t)g.pitch * 2 };
    const cuuint32_t box[2] = { BOX_W, CH_ROWS };
    const cuuint32_t estr[2] = { 1, 1 };
    const CUresult cr = st->encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(y16), gdim, gstr, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { st->err = "cuTensorMapEncodeTiled failed (" + std::to_string((int)cr) + ")"; return 1; }

    /* one persistent CTA per SM; tiny inputs: at least ~16 row pairs per CTA so that the per-unit costs stay amortised */
    const long long T = (long long)p.nframes * p.bands * p.ppb, want = (T + 15) / 16;
    const int grid = (int)(want < 1 ? 1 : (want < num_sms ? want : num_sms));
#ifdef SRCNN_PROF
    static long long* d_prof = nullptr;
    if (!d_prof) cudaMalloc(&d_prof, NWARPS * 4 * 8);
    p.prof = d_prof;
#endif
#ifdef SRCNN_TL
    static long long* d_tl = nullptr;
    if (!d_tl) cudaMalloc(&d_tl, TL_N * TL_W * 8);
    cudaMemsetAsync(d_tl, 0, TL_N * TL_W * 8, s);
    p.prof = d_tl;
#endif
    if (p.rgb) conv_tc_kernel<true><<<grid, NWARPS * 32, SMEM_BYTES, s>>>(p, map);
    else       launch_pdl(conv_tc_kernel<false>, dim3(grid), dim3(NWARPS * 32), SMEM_BYTES, s, p, map);
#ifdef SRCNN_PROF
    {
        cudaStreamSynchronize(s);
        long long h[NWARPS * 4];
        cudaMemcpy(h, d_prof, sizeof(h), cudaMemcpyDeviceToHost);
        static const char* role[6] = { "gather", "E2", "E1.even", "E1.odd", "im2col", "ctl" };
        static const char* ctl[4] = { "TMA", "MMA1", "MMA2", "MMA3" };
        for (int w = 0; w < NWARPS; w += (w >= W_CTL ? 1 : 4))
            fprintf(stderr, "PROF %-8s total %8lld  wait0 %5.1f%%  wait1 %5.1f%%\n", w >= W_CTL ? ctl[w - W_CTL] : role[w / 4], h[w * 4],
                    100.0 * h[w * 4 + 1] / h[w * 4], 100.0 * h[w * 4 + 2] / h[w * 4]);
    }
#endif
#ifdef SRCNN_TL
    if (const char* path = getenv("SRCNN_TL_DUMP")) {
        /* every launch overwrites the dump: the file holds the last (warm) launch */
        cudaStreamSynchronize(s);
        std::vector<long long> h(TL_N * TL_W);
        cudaMemcpy(h.data(), d_tl, h.size() * 8, cudaMemcpyDeviceToHost);
        if (FILE* f = fopen(path, "w")) {
            for (int r = 0; r < TL_N; r++) {
                fprintf(f, "%d", TL_R0 + r);
                for (int e = 0; e < TL_W; e++) fprintf(f, " %lld", h[r * TL_W + e]);
                fprintf(f, "\n");
            }
            fclose(f);
        }
    }
#endif
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { st->err = cudaGetErrorString(e); return 1; }
    return 0;
}

} // namespace srcnn
