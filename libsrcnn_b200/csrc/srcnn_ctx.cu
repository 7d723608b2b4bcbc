/* Context, pass orchestration and the extern "C" layer (include/srcnn_b200.h).
 *
 * One pass = reference libsrcnn::doSRCNN (src/libsrcnn.cpp:628-923) as a launch sequence on one stream:
 *   colour-in -> resize pass 1 -> resize pass 2 -> conv-input pack -> fused conv I/II/III -> colour-out
 * restricted to the output rows [row0,row1) of a strip (the whole image for the ordinary call) plus the 6-row
 * halo of interpolated Y the convolutions need.  No CPU arithmetic on image data anywhere in this file.
 */
#include "../../include/srcnn_b200.h"
#include "kernels.h"
#include "contrib.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <map>
#include <string>
#include <tuple>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

using namespace srcnn;

namespace {

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        const size_t want = bytes + bytes / 8 + 256;          /* grow-only with slack */
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct PinBuf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

/* host-side memcpy spread over a few persistent worker threads (pinned staging <-> the caller's pageable rows), and
 * first-touch of a fresh result buffer in the background: a buffer that just came from new[] has no pages yet, and the kernel
 * zeroes every page on its first write -- 0.7 ms for a 25 MB result on one thread (tools/pageable_probe.py), paid here by the
 * workers while the GPU is still busy with the first sub-strips */
class HostCopier {
public:
    ~HostCopier() { stop(); }
    void start(unsigned workers)
    {
        if (!th_.empty() || workers == 0) return;
        for (unsigned i = 0; i < workers; i++) th_.emplace_back([this, i] { loop(i); });
    }
    void stop()
    {
        wait();
        { std::lock_guard<std::mutex> g(m_); quit_ = true; }
        cv_.notify_all();
        for (auto& t : th_) t.join();
        th_.clear(); quit_ = false;
    }
    void copy(unsigned char* dst, const unsigned char* src, size_t n)
    {
        wait();                                                                 /* (a background touch may still be running) */
        const size_t parts = th_.size() + 1;
        if (parts == 1 || n < (size_t)256 << 10) { memcpy(dst, src, n); return; }
        const size_t slice = ((n + parts - 1) / parts + 4095) / 4096 * 4096;
        post(dst, src, n, slice, 1);
        memcpy(dst, src, std::min(slice, n));                                   /* the caller takes slice 0 */
        wait();
    }
    /* write one byte into every 4 KB of [dst, dst + n) on the worker threads; returns at once, wait() joins */
    void touch_async(unsigned char* dst, size_t n)
    {
        wait();
        if (th_.empty() || n < (size_t)1 << 20) return;
        const size_t slice = ((n + th_.size() - 1) / th_.size() + 4095) / 4096 * 4096;
        post(dst, nullptr, n, slice, 0);
    }
    void wait()
    {
        std::unique_lock<std::mutex> l(m_);
        done_.wait(l, [this] { return pending_ == 0; });
    }
    bool idle() { std::lock_guard<std::mutex> g(m_); return pending_ == 0; }
private:
    void post(unsigned char* dst, const unsigned char* src, size_t n, size_t slice, unsigned first)
    {
        {
            std::lock_guard<std::mutex> g(m_);
            dst_ = dst; src_ = src; n_ = n; slice_ = slice; first_ = first; pending_ = (unsigned)th_.size(); gen_++;
            hint_.store(gen_, std::memory_order_release);
        }
        cv_.notify_all();
    }
    void loop(unsigned id)
    {
        unsigned long seen = 0;
        for (;;) {
            unsigned char* d; const unsigned char* s; size_t n, sl; unsigned first;
            /* sub-strips of one call arrive ~100 us apart: poll for that long before going to sleep, so that the next job does
             * not pay a futex wake-up (tens of microseconds of jitter per sub-strip on the drop-in path) */
            for (int spin = 0; spin < 20000 && hint_.load(std::memory_order_acquire) == seen; spin++) cpu_relax();
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [&] { return quit_ || gen_ != seen; });
                if (quit_) return;
                seen = gen_; d = dst_; s = src_; n = n_; sl = slice_; first = first_;
            }
            const size_t a = (size_t)(id + first) * sl;
            if (a < n) {
                const size_t len = std::min(sl, n - a);
                if (s) memcpy(d + a, s + a, len);
                else for (size_t o = 0; o < len; o += 4096) static_cast<volatile unsigned char*>(d + a)[o] = 0;
            }
            { std::lock_guard<std::mutex> g(m_); if (--pending_ == 0) done_.notify_all(); }
        }
    }
    std::vector<std::thread> th_;
    std::mutex m_; std::condition_variable cv_, done_;
    unsigned char* dst_ = nullptr; const unsigned char* src_ = nullptr; size_t n_ = 0, slice_ = 0;
    unsigned pending_ = 0, first_ = 0; unsigned long gen_ = 0; bool quit_ = false;
    std::atomic<unsigned long> hint_{0};      /* copy of gen_ the workers may poll without the lock */
    static void cpu_relax()
    {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#elif defined(__aarch64__)
        asm volatile("yield");
#endif
    }
};

constexpr unsigned kMaxPipeStrips = 32, kStageSlots = 3;

struct TableDev {
    AxisTable host;
    void* dev = nullptr;
    AxisTableDev view{};
    bool quads_ok = false;      /* compact form: 4 adjacent destination samples read <= 4 adjacent source samples (color.cu) */
    bool pairs_ok = false;      /* compact form: the windows of destination samples 2j, 2j+1 start 0 or 1 apart (prep.cu: prep2_kernel) */
};

/* geometry of one pass restricted to a row strip */
struct Plan {
    unsigned w, h, d, ow, oh;
    unsigned r0, r1;        /* output rows produced                                   */
    unsigned e0, e1;        /* rows of the resized planes computed (strip + conv halo) */
    unsigned s0, s1;        /* source rows read                                        */
    int fy, fc;             /* filter of plane 0 / of the other planes                 */
    bool hfirst;            /* ow <= w: horizontal pass first (src/frawscale.cpp:195)  */
    bool need_h, need_v;
};

} // namespace

struct srcnn_b200_ctx {
    int device = 0, num_sms = 0;
    cudaStream_t stream = nullptr, stream_in = nullptr, stream_out = nullptr;
    DevBuf src8[2], planes_src, tmp, planes_dst, y16, yprime, out8[2], conv8[2], chain[2], chroma_src;
    PinBuf pin_in[2], pin_out[2], pin_conv[2];
    std::map<std::tuple<int, unsigned, unsigned>, TableDev*> tables;
    ConvTcState* tc = nullptr;
    int conv_mode = SRCNN_B200_CONV_TENSOR;
    bool chain_tensor = false;      /* step-scale chains: tensor kernel also for the intermediate passes */
    bool timing = false;
    cudaEvent_t ev[8] = {};
    int ev_alias[8] = { 0, 1, 2, 3, 4, 5, 6, 7 };   /* event that stands for stage boundary i (Stamp) */
    cudaEvent_t ev_in[2] = {}, ev_k[2] = {}, ev_out[2] = {};
    /* asynchronous host passes (srcnn_b200_pass_host_async): call k uses device buffer set k & 1; ev_ak / ev_ao mark the end of
     * the kernels / of the downloads of the last call on that set */
    cudaEvent_t ev_ak[2] = {}, ev_ao[2] = {};
    bool async_used[2] = { false, false };
    unsigned long long async_seq = 0;
    int async_rc = 0;
    float ms[7] = {};
    unsigned long long launches = 0;
    std::string err;
    /* host pipeline (srcnn_b200_pass_host_rows) */
    size_t strip_bytes = 0;                   /* result bytes per pipelined sub-strip; 0 = by destination: 4 MB into pinned
                                                 memory, 8 MB into pageable memory (measured, tools/e2e_probe.py); SRCNN_B200_STRIP_MB */
    bool stage_pageable = true;               /* pageable results go through pinned staging (SRCNN_B200_STAGE=0: off) */
    std::vector<cudaEvent_t> pipe_ev;         /* 3 per sub-strip: uploaded, computed, downloaded                      */
    HostCopier copier, toucher;
    bool fuse_front = true;                   /* one-kernel front end (SRCNN_B200_FUSE=0: the separate kernels)       */
    int keep_planes = 0;                      /* tensor mode: 1 = also write the FP32 interpolated Y and the output-size chroma planes
                                                 (getters; first-generation front end), 2 = the FP32 Y plane only, product route kept */
    /* CUDA graphs (srcnn_b200_set_graphs): a device-resident pass that is repeated with the same geometry and pointers is captured
     * once (prep -> conv -> colour-out with their programmatic-dependent-launch edges) and replayed with one cudaGraphLaunch */
    bool graphs = false;
    struct GraphEntry { std::vector<unsigned long long> key; int seen = 0; cudaGraphExec_t exec = nullptr; };
    std::vector<GraphEntry> graph_cache;
    unsigned long long graph_replays = 0;
    bool ramp = true;                         /* host pipeline: small first sub-strips (SRCNN_B200_RAMP=0: uniform cuts) */
    bool prefault = true;                     /* pageable results: first touch by the worker threads while the GPU works (SRCNN_B200_PREFAULT=0) */
    bool stage_in = true;                     /* pageable SOURCE rows go through pinned staging too (SRCNN_B200_STAGE_IN=0: off) */
    bool stage_all = true;                    /* results <= 96 MB: one staging slot per sub-strip (SRCNN_B200_STAGE_ALL=0: three slots) */
    bool trace = false;                       /* SRCNN_B200_TRACE=1: host-side timeline of every host pass on stderr */
    bool prep2 = true;                        /* second-generation front end on the x2 hand-over route (SRCNN_B200_PREP2=0: off) */
    bool last_has_y32 = false;
    bool last_has_chroma = true;              /* output-size Cb / Cr / A planes of the last pass exist (getters 1-3)   */
    bool group_frames = true;                 /* batches: one conv launch per group of small frames (SRCNN_B200_GROUP=0) */
    bool chroma_up = true;                    /* x2: chroma resampled inside colour-out (SRCNN_B200_CHROMA_UP=0: off)  */
    unsigned last_group = 1;                  /* frames covered by the last pass / group (and by its timings)          */
    bool last_has_yp = true;                  /* the FP32 network output Y' of the last pass exists (plane getter 4)  */
    bool fuse_color = false;                  /* tensor mode: colour-out inside the conv kernel's store stage (SRCNN_B200_FUSE_COLOR=1).
                                               * Off by default: measured 0.202 ms fused vs 0.155 + 0.023 ms separate per 4K frame --
                                               * the extra work on the conv III gather warps lengthens the kernel's critical loop. */
    /* last pass, for the plane getters */
    Plan last{}; bool have_last = false; size_t last_plane_stride = 0;
};

namespace {

int fail_cuda(srcnn_b200_ctx* c, cudaError_t e, const char* what)
{
    char buf[256];
    snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    c->err = buf;
    return SRCNN_B200_ECUDA;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail_cuda(ctx, e_, #call); } while (0)

bool host_is_pinned(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

cudaError_t ensure_pipe_events(srcnn_b200_ctx* c, unsigned nstrips)
{
    while (c->pipe_ev.size() < (size_t)3 * nstrips) {
        cudaEvent_t e;
        cudaError_t r = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        if (r != cudaSuccess) return r;
        c->pipe_ev.push_back(e);
    }
    return cudaSuccess;
}

int get_table(srcnn_b200_ctx* ctx, int filter, unsigned dst, unsigned src, TableDev** out)
{
    auto key = std::make_tuple(filter, dst, src);
    auto it = ctx->tables.find(key);
    if (it != ctx->tables.end()) { *out = it->second; return 0; }
    TableDev* t = new TableDev;
    build_axis_table(filter, dst, src, t->host);
    /* device image: int arrays left | count | dleft | cleft, then double arrays w | dw | cw */
    const AxisTable& h = t->host;
    const size_t ints = (size_t)dst * sizeof(int);
    const size_t ints_al = (4 * ints + 15) / 16 * 16;
    const size_t wbytes = h.w.size() * sizeof(double), cbytes = h.cw.size() * sizeof(double);
    const size_t pbytes = (h.compact && h.cwin == 2) ? (size_t)dst * 16 : 0;      /* pack2 (color.cu): every two-tap table */
    const size_t poff = (ints_al + 2 * wbytes + cbytes + 15) / 16 * 16;
    /* cwt: the compact weights transposed, [tap][dst] -- adjacent destination samples are adjacent in memory, so a warp reading
     * one tap of its 64 columns touches 512 contiguous bytes instead of 32 scattered 64-byte-strided words (prep.cu: prep2_kernel) */
    const size_t toff = (poff + pbytes + 15) / 16 * 16, tbytes = h.compact ? cbytes : 0;
    const size_t bytes = toff + tbytes;
    std::vector<unsigned char> img(bytes, 0);
    int* L = reinterpret_cast<int*>(img.data());
    int* C = L + dst;
    for (unsigned u = 0; u < dst; u++) { L[u] = h.left[u]; C[u] = h.right[u] - h.left[u] + 1; }
    memcpy(img.data() + ints_al, h.w.data(), wbytes);
    if (h.dense) {
        memcpy(C + dst, h.dleft.data(), ints);
        memcpy(img.data() + ints_al + wbytes, h.dw.data(), wbytes);
    }
    if (h.compact) {
        memcpy(C + 2 * dst, h.cleft.data(), ints);
        memcpy(img.data() + ints_al + 2 * wbytes, h.cw.data(), cbytes);
    }
    if (pbytes) {
        int* P = reinterpret_cast<int*>(img.data() + poff);
        for (unsigned u = 0; u < dst; u++) {
            const float w0 = (float)h.cw[2 * (size_t)u], w1 = (float)h.cw[2 * (size_t)u + 1];
            uint32_t b1; memcpy(&b1, &w1, 4);
            const bool swap = (b1 & 0x7FFFFFu) != 0;             /* tap 1 is the general one -> A = tap 1, B = tap 0 */
            const float wA = swap ? w1 : w0, wB = swap ? w0 : w1;
            P[4 * u] = h.cleft[u] + (swap ? 1 : 0); memcpy(&P[4 * u + 1], &wA, 4); memcpy(&P[4 * u + 2], &wB, 4);
            P[4 * u + 3] = h.cleft[u] + (swap ? 0 : 1);
        }
    }
    if (tbytes) {
        double* T = reinterpret_cast<double*>(img.data() + toff);
        for (unsigned u = 0; u < dst; u++)
            for (int i = 0; i < h.cwin; i++) T[(size_t)i * dst + u] = h.cw[(size_t)u * h.cwin + i];
    }
    cudaError_t e = cudaMalloc(&t->dev, bytes);
    if (e != cudaSuccess) { delete t; return fail_cuda(ctx, e, "cudaMalloc(table)"); }
    e = cudaMemcpyAsync(t->dev, img.data(), bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);      /* img is a temporary */
    if (e != cudaSuccess) { cudaFree(t->dev); delete t; return fail_cuda(ctx, e, "upload(table)"); }
    const unsigned char* base = static_cast<unsigned char*>(t->dev);
    t->view.left  = reinterpret_cast<const int*>(base);
    t->view.count = t->view.left + dst;
    t->view.w     = reinterpret_cast<const double*>(base + ints_al);
    t->view.win   = h.win;
    t->view.dleft = h.dense ? t->view.left + 2 * dst : nullptr;
    t->view.dw    = h.dense ? reinterpret_cast<const double*>(base + ints_al + wbytes) : nullptr;
    t->view.cleft = h.compact ? t->view.left + 3 * dst : nullptr;
    t->view.cw    = h.compact ? reinterpret_cast<const double*>(base + ints_al + 2 * wbytes) : nullptr;
    t->view.cwin  = h.cwin;
    t->view.cwt   = tbytes ? reinterpret_cast<const double*>(base + toff) : nullptr;
    t->view.exact29 = h.exact29 ? 1 : 0;
    t->view.pow2pair = h.pow2pair ? 1 : 0;
    t->view.pack2 = pbytes ? base + poff : nullptr;
    t->quads_ok = h.compact && h.cwin == 2 && color_out_up_quads_ok(h.cleft.data(), dst);
    t->pairs_ok = h.compact && prep2_pairs_ok(h.cleft.data(), dst);
    ctx->tables[key] = t;
    *out = t;
    return 0;
}

/* source rows read by the vertical pass for destination rows [e0,e1): the dense windows when the tables have them */
void source_rows(const AxisTable& a, const AxisTable& b, unsigned e0, unsigned e1, unsigned& s0, unsigned& s1)
{
    int lo = INT32_MAX, hi = -1;
    for (const AxisTable* t : { &a, &b })
        for (unsigned r = e0; r < e1; r++) {
            lo = std::min(lo, t->left[r]); hi = std::max(hi, t->right[r]);
            if (t->dense) { lo = std::min(lo, t->dleft[r]); hi = std::max(hi, t->dleft[r] + t->win - 1); }
            if (t->compact) { lo = std::min(lo, t->cleft[r]); hi = std::max(hi, t->cleft[r] + t->cwin - 1); }
        }
    s0 = (unsigned)lo; s1 = (unsigned)hi + 1;
}

int make_plan(unsigned w, unsigned h, unsigned d, float multiply, int filter, unsigned row0, unsigned row1, Plan& p)
{
    if (!w || !h || (d != 3 && d != 4)) return SRCNN_B200_EINVAL;
    if ((float)w * multiply <= 0.f || (float)h * multiply <= 0.f) return SRCNN_B200_ESIZE;
    p.w = w; p.h = h; p.d = d;
    srcnn_b200_output_dims(w, h, multiply, &p.ow, &p.oh);
    if (!p.ow || !p.oh) return SRCNN_B200_ESIZE;
    if (row1 > p.oh || row0 >= row1) return SRCNN_B200_EINVAL;
    if (p.oh > 65535u || h > 65535u) return SRCNN_B200_EINVAL;      /* rows map to gridDim.y */
    p.r0 = row0; p.r1 = row1;
    p.e0 = row0 > (unsigned)kHaloRows ? row0 - kHaloRows : 0;
    p.e1 = std::min(p.oh, row1 + kHaloRows);
    p.fy = (filter >= kNearest && filter <= kBspline) ? filter : kNearest;   /* src/libsrcnn.cpp:677-699 */
    p.fc = (filter == kNearest) ? kNearest : kBilinear;                      /* src/libsrcnn.cpp:703-713 */
    p.hfirst = (p.ow <= w);
    p.need_h = (p.ow != w); p.need_v = (p.oh != h);
    return 0;
}

/* per-kernel CUDA events of one pass (only while timing is on).  skip(n) marks n stage boundaries that coincide with the
 * previous one (stages folded into one kernel): no event is recorded for them -- every record between two kernels costs
 * GPU idle time -- and collect_timing reports 0 for those stages. */
struct Stamp {
    srcnn_b200_ctx* c; int i = 0;
    explicit Stamp(srcnn_b200_ctx* c_) : c(c_) { c->ev_alias[0] = 0; mark(); }
    void mark() { if (c->timing && i < 8) { cudaEventRecord(c->ev[i], c->stream); c->ev_alias[i] = i; } i++; }
    void skip(int n) { for (; n > 0 && i < 8; n--, i++) c->ev_alias[i] = c->ev_alias[i - 1]; }
};

/* d_src_base points at source row `src_base_row` (device; kWholeImage = row 0, kStripRows = row p.s0);
 * d_out/d_conv at output row p.r0 (device).  Asynchronous on ctx->stream. */
constexpr long kStripRows = -1;
int run_pass(srcnn_b200_ctx* ctx, Plan& p, const uint8_t* d_src_base, long src_base_row,
             uint8_t* d_out, uint8_t* d_conv)
{
    TableDev *vy = nullptr, *vc = nullptr, *hy = nullptr, *hc = nullptr;
    int rc;
    if (p.need_v) {
        if ((rc = get_table(ctx, p.fy, p.oh, p.h, &vy))) return rc;
        if ((rc = get_table(ctx, p.fc, p.oh, p.h, &vc))) return rc;
        source_rows(vy->host, vc->host, p.e0, p.e1, p.s0, p.s1);
    } else { p.s0 = p.e0; p.s1 = p.e1; }
    if (p.need_h) {
        if ((rc = get_table(ctx, p.fy, p.ow, p.w, &hy))) return rc;
        if ((rc = get_table(ctx, p.fc, p.ow, p.w, &hc))) return rc;
    }
    const uint8_t* d_src = src_base_row == kStripRows ? d_src_base
                                                      : d_src_base + ((size_t)p.s0 - (size_t)src_base_row) * p.w * p.d;

    const unsigned srows = p.s1 - p.s0, erows = p.e1 - p.e0, orows = p.r1 - p.r0;
    auto al = [](size_t n) { return (n + 63) / 64 * 64; };
    const size_t ps_src = al((size_t)srows * p.w), ps_dst = al((size_t)erows * p.ow);
    const size_t ps_tmp = al(p.hfirst ? (size_t)srows * p.ow : (size_t)erows * p.w);
    const bool tensor = ctx->conv_mode == SRCNN_B200_CONV_TENSOR;
    CK(ctx->planes_dst.ensure(ps_dst * p.d * sizeof(float)));
    CK(ctx->yprime.ensure((size_t)orows * p.ow * sizeof(float)));
    float* P_dst = static_cast<float*>(ctx->planes_dst.p);
    float* Yp    = static_cast<float*>(ctx->yprime.p);
    const Y16Geom g = y16_geometry(p.ow, orows);
    if (tensor) CK(ctx->y16.ensure((size_t)g.pitch * g.rows * sizeof(__half)));
    cudaStream_t s = ctx->stream;

    Stamp st(ctx);
    /* 1.-4. front end.  Preferred: ONE kernel (colour-in + both resampling passes + conv-input pack, prep.cu) */
    const bool y32 = !tensor || ctx->keep_planes != 0;
    bool fused = false, chroma_up = false;
    const size_t ps_csrc = al((size_t)srows * p.w);
    if (ctx->fuse_front && !p.hfirst && p.need_h && p.need_v && prep_supported(vy->view, vc->view, hy->view, hc->view)) {
        /* x2-type magnification: Cb / Cr / A stay at source resolution and are resampled inside colour-out (the output-size
         * chroma planes then never exist -- unless the plane getters were asked to keep them) */
        chroma_up = ctx->chroma_up && ctx->keep_planes != 1 && !(tensor && ctx->fuse_color) && prep_chroma_handover(vc->view, hc->view, tensor) && (p.ow % 4) == 0 &&
                    hc->quads_ok && ((uintptr_t)d_out % (p.d == 4 ? 16 : 4)) == 0 &&
                    (!d_conv || ((uintptr_t)d_conv % 4) == 0);
        if (chroma_up) CK(ctx->chroma_src.ensure(ps_csrc * (p.d - 1) * sizeof(float)));
        int lp = 0;
        if (chroma_up && ctx->prep2)
            /* product route: second-generation kernel; one double FMA per tap in tensor mode, the reference's separately rounded
             * multiply and add in bit-exact mode */
            lp = launch_prep2(d_src, p.w, p.d, p.s0, p.s1 - p.s0, p.ow, p.oh, p.e0, p.e1, y32 ? P_dst : nullptr, tensor,
                              tensor ? static_cast<__half*>(ctx->y16.p) : nullptr, g, p.r0,
                              vy->view, vc->view, hy->view, hc->view, vy->host.cleft.data(), vc->host.cleft.data(),
                              hy->host.cleft.data(), hc->host.cleft.data(), hy->pairs_ok, s,
                              static_cast<float*>(ctx->chroma_src.p), ps_csrc);
        /* (the first-generation kernel hands chroma over only for the exact, power-of-two tables) */
        if (lp == 0 && chroma_up && !prep_chroma_handover(vc->view, hc->view, false)) chroma_up = false;
        if (lp == 0)
          lp = launch_prep(d_src, p.w, p.d, p.s0, p.ow, p.oh, p.e0, p.e1, P_dst, ps_dst, y32,
                                   tensor ? static_cast<__half*>(ctx->y16.p) : nullptr, g, p.r0,
                                   vy->view, vc->view, hy->view, hc->view, vy->host.cleft.data(), vc->host.cleft.data(),
                                   hy->host.cleft.data(), hc->host.cleft.data(), s,
                                   chroma_up ? static_cast<float*>(ctx->chroma_src.p) : nullptr, ps_csrc);
        if (lp < 0) return fail_cuda(ctx, cudaGetLastError(), "prep_kernel launch");
        fused = lp != 0;
        if (!fused) chroma_up = false;
    }
    ctx->last_has_chroma = !chroma_up;
    if (fused) {
        ctx->launches++;
        st.mark(); st.skip(3);
        ctx->last_has_y32 = y32;
    } else {
        CK(ctx->planes_src.ensure(ps_src * p.d * sizeof(float)));
        if (p.need_h && p.need_v) CK(ctx->tmp.ensure(ps_tmp * p.d * sizeof(float)));
        float* P_src = static_cast<float*>(ctx->planes_src.p);
        float* P_tmp = static_cast<float*>(ctx->tmp.p);
        ctx->last_has_y32 = true;
        /* 1. colour-in on the source rows the strip needs */
        launch_color_in(d_src, (size_t)srows * p.w, p.d, P_src, ps_src, s); ctx->launches++;
        st.mark();
        /* 2./3. separable resize; order as in FRAWResizeEngine::scale (src/frawscale.cpp:195-278) */
        if (!p.need_h && !p.need_v) {
            /* equal size: plain copy of every plane (the reference copies only half of it, a bug we do not keep) */
            for (unsigned pl = 0; pl < p.d; pl++)
                CK(cudaMemcpyAsync(P_dst + pl * ps_dst, P_src + pl * ps_src, (size_t)erows * p.ow * sizeof(float),
                                   cudaMemcpyDeviceToDevice, s));
            st.mark(); st.mark();
        } else if (p.hfirst) {
            if (p.need_h) {
                float* o = p.need_v ? P_tmp : P_dst; const size_t os = p.need_v ? ps_tmp : ps_dst;
                launch_resize_horizontal(P_src, ps_src, p.w, o, os, p.ow, srows, p.d, hy->view, hc->view, s); ctx->launches += 2;
            }
            st.mark();
            if (p.need_v) {
                const float* i = p.need_h ? P_tmp : P_src; const size_t is = p.need_h ? ps_tmp : ps_src;
                launch_resize_vertical(i, is, p.ow, p.s0, P_dst, ps_dst, p.e0, erows, p.d, vy->view, vc->view, s); ctx->launches += 2;
            }
            st.mark();
        } else {
            if (p.need_v) {
                float* o = p.need_h ? P_tmp : P_dst; const size_t os = p.need_h ? ps_tmp : ps_dst;
                launch_resize_vertical(P_src, ps_src, p.w, p.s0, o, os, p.e0, erows, p.d, vy->view, vc->view, s); ctx->launches += 2;
            }
            st.mark();
            if (p.need_h) {
                const float* i = p.need_v ? P_tmp : P_src; const size_t is = p.need_v ? ps_tmp : ps_src;
                launch_resize_horizontal(i, is, p.w, P_dst, ps_dst, p.ow, erows, p.d, hy->view, hc->view, s); ctx->launches += 2;
            }
            st.mark();
        }
        /* 4. conv-input pack */
        if (tensor) { launch_pack_y16(P_dst, p.ow, p.oh, p.e0, static_cast<__half*>(ctx->y16.p), g, p.r0, s); ctx->launches++; }
        st.mark();
    }
    /* 5. fused conv I+II+III on the Y plane; in tensor mode the kernel's store stage also does 6. colour-out (+ conv
     * plane) for the strip rows, so Y' never travels through HBM (it is written only for the plane getters) */
    const size_t off = (size_t)(p.r0 - p.e0) * p.ow;
    const bool fuse_out = tensor && ctx->fuse_color && (p.d != 4 || ((uintptr_t)d_out & 3u) == 0);
    if (tensor) {
        ConvTcColor col;
        if (fuse_out) {
            col.cb = P_dst + ps_dst + off; col.cr = P_dst + 2 * ps_dst + off;
            col.al = p.d == 4 ? P_dst + 3 * ps_dst + off : nullptr;
            col.rgb = d_out; col.conv8 = d_conv; col.d = p.d;
        }
        ctx->last_has_yp = !fuse_out || ctx->keep_planes != 0;
        if (launch_conv_tc(ctx->tc, static_cast<const __half*>(ctx->y16.p), g, p.ow, p.oh, ctx->last_has_yp ? Yp : nullptr,
                           p.r0, p.r1, ctx->num_sms, s, fuse_out ? &col : nullptr)) {
            ctx->err = std::string("conv_tc: ") + conv_tc_error(ctx->tc);
            return SRCNN_B200_ECUDA;
        }
        ctx->launches++;
    } else {
        if (const int ce = launch_conv_exact(P_dst, p.ow, p.oh, p.e0, erows, Yp, p.r0, p.r1, s)) {
            ctx->err = std::string("conv_exact set-up: ") + cudaGetErrorString((cudaError_t)ce);
            return SRCNN_B200_ECUDA;
        }
        ctx->launches++;
        ctx->last_has_yp = true;
    }
    st.mark();
    /* 6. colour-out (+ conv plane) for the strip rows, unless the conv kernel already stored them */
    if (!fuse_out && chroma_up) {
        if (!launch_color_out_up(Yp, static_cast<const float*>(ctx->chroma_src.p), ps_csrc, p.w, p.s0, vc->view, hc->view,
                                 p.ow, p.r0, p.r1, p.d, d_out, d_conv, s, ctx->num_sms)) {
            ctx->err = "colour-out (x2 chroma path): geometry not covered";
            return SRCNN_B200_ECUDA;
        }
        ctx->launches++;
    } else if (!fuse_out) {
        launch_color_out(Yp, P_dst + ps_dst + off, P_dst + 2 * ps_dst + off, p.d == 4 ? P_dst + 3 * ps_dst + off : nullptr,
                         (size_t)orows * p.ow, p.d, d_out, d_conv, s); ctx->launches++;
    }
    st.mark();
    CK(cudaGetLastError());
    ctx->last = p; ctx->have_last = true; ctx->last_plane_stride = ps_dst; ctx->last_group = 1;
    return 0;
}

int collect_timing(srcnn_b200_ctx* ctx)
{
    if (!ctx->timing) return 0;
    for (int i = 0; i < 6; i++) {
        float v = 0.f;
        const int a = ctx->ev_alias[i], b = ctx->ev_alias[i + 1];
        if (a != b && cudaEventElapsedTime(&v, ctx->ev[a], ctx->ev[b]) != cudaSuccess) v = -1.f;
        ctx->ms[i] = v;
    }
    float v = 0.f;
    if (cudaEventElapsedTime(&v, ctx->ev[0], ctx->ev[6]) != cudaSuccess) v = -1.f;
    ctx->ms[6] = v;
    return 0;
}


/* ---- frame groups (same geometry): G front-end launches, ONE fused-conv launch over all G frames, G colour-out
 * launches.  Small frames do not fill the persistent conv kernel (a 1024 x 1024 result is ~31 row pairs per CTA, a third of
 * them pipeline fill and halo); a group amortises that.  Only the product route is grouped: tensor-core conv, fused
 * front end, chroma resampled inside colour-out; anything else runs frame by frame through run_pass. */
unsigned group_size(const srcnn_b200_ctx* ctx, const Plan& p, unsigned n)
{
    if (ctx->conv_mode != SRCNN_B200_CONV_TENSOR || !ctx->fuse_front || !ctx->chroma_up || ctx->keep_planes != 0 ||
        ctx->fuse_color || !ctx->group_frames) return 1;
    if (p.hfirst || !p.need_h || !p.need_v || (p.ow % 4) || p.r0 != 0 || p.r1 != p.oh) return 1;
    const size_t px = (size_t)p.ow * p.oh;
    const size_t g = ((size_t)8 << 20) / (px ? px : 1);                  /* groups of up to ~8 Mpx of output */
    return (unsigned)std::max<size_t>(1, std::min<size_t>(std::min<size_t>(g, 16), n));
}

/* d_src: G frames of w*h*d bytes; d_out: G frames of ow*oh*d; d_conv: G frames of ow*oh or NULL.  Returns 1 if the geometry
 * is not covered by the grouped route (nothing launched; the caller falls back to run_pass per frame). */
int run_group(srcnn_b200_ctx* ctx, Plan& p, unsigned G, const uint8_t* d_src, uint8_t* d_out, uint8_t* d_conv)
{
    TableDev *vy = nullptr, *vc = nullptr, *hy = nullptr, *hc = nullptr;
    int rc;
    if ((rc = get_table(ctx, p.fy, p.oh, p.h, &vy))) return rc;
    if ((rc = get_table(ctx, p.fc, p.oh, p.h, &vc))) return rc;
    if ((rc = get_table(ctx, p.fy, p.ow, p.w, &hy))) return rc;
    if ((rc = get_table(ctx, p.fc, p.ow, p.w, &hc))) return rc;
    source_rows(vy->host, vc->host, p.e0, p.e1, p.s0, p.s1);
    const size_t ib = (size_t)p.w * p.h * p.d, ob = (size_t)p.ow * p.oh * p.d, cb = (size_t)p.ow * p.oh;
    if (!prep_supported(vy->view, vc->view, hy->view, hc->view) || !prep_chroma_handover(vc->view, hc->view, true) || !hc->quads_ok ||
        ((uintptr_t)d_out % 16) || (ob % 16) || (d_conv && (((uintptr_t)d_conv % 4) || (cb % 4)))) return 1;
    auto al = [](size_t n) { return (n + 63) / 64 * 64; };
    const unsigned srows = p.s1 - p.s0;
    const size_t ps_csrc = al((size_t)srows * p.w), csrc_f = ps_csrc * (p.d - 1);
    const Y16Geom g = y16_geometry(p.ow, p.oh);
    const size_t y16_f = (size_t)g.pitch * g.rows, yp_f = al((size_t)p.ow * p.oh);
    CK(ctx->y16.ensure(y16_f * G * sizeof(__half)));
    CK(ctx->chroma_src.ensure(csrc_f * G * sizeof(float)));
    CK(ctx->yprime.ensure(yp_f * G * sizeof(float)));
    __half* Y16 = static_cast<__half*>(ctx->y16.p);
    float* CS = static_cast<float*>(ctx->chroma_src.p);
    float* Yp = static_cast<float*>(ctx->yprime.p);
    cudaStream_t s = ctx->stream;
    Stamp st(ctx);                                                  /* timings then describe the whole group */
    int lp = 0;
    if (ctx->prep2)
        lp = launch_prep2(d_src + (size_t)p.s0 * p.w * p.d, p.w, p.d, p.s0, p.s1 - p.s0, p.ow, p.oh, p.e0, p.e1, nullptr, true, Y16, g, p.r0,
                          vy->view, vc->view, hy->view, hc->view, vy->host.cleft.data(), vc->host.cleft.data(),
                          hy->host.cleft.data(), hc->host.cleft.data(), hy->pairs_ok, s, CS, ps_csrc, G, ib, y16_f, csrc_f);
    if (lp == 0 && !prep_chroma_handover(vc->view, hc->view, false)) return 1;      /* (approximate hand-over: second-generation kernel only) */
    if (lp == 0)
      lp = launch_prep(d_src + (size_t)p.s0 * p.w * p.d, p.w, p.d, p.s0, p.ow, p.oh, p.e0, p.e1, nullptr, 0, false,
                               Y16, g, p.r0, vy->view, vc->view, hy->view, hc->view, vy->host.cleft.data(),
                               vc->host.cleft.data(), hy->host.cleft.data(), hc->host.cleft.data(), s, CS, ps_csrc,
                               G, ib, y16_f, csrc_f);
    if (lp < 0) return fail_cuda(ctx, cudaGetLastError(), "prep_kernel launch");
    if (!lp) return 1;                                              /* geometry not covered: nothing launched */
    ctx->launches++;
    st.mark(); st.skip(3);
    if (launch_conv_tc(ctx->tc, Y16, g, p.ow, p.oh, Yp, p.r0, p.r1, ctx->num_sms, s, nullptr, G, yp_f)) {
        ctx->err = std::string("conv_tc: ") + conv_tc_error(ctx->tc);
        return SRCNN_B200_ECUDA;
    }
    ctx->launches++;
    st.mark();
    if (!launch_color_out_up(Yp, CS, ps_csrc, p.w, p.s0, vc->view, hc->view, p.ow, p.r0, p.r1, p.d, d_out, d_conv, s,
                             ctx->num_sms, G, yp_f, csrc_f, ob, cb)) {
        ctx->err = "frame group: colour-out geometry not covered"; return SRCNN_B200_ECUDA;
    }
    ctx->launches++;
    st.mark();
    CK(cudaGetLastError());
    ctx->last = p; ctx->have_last = false;                          /* (the plane getters describe single passes only) */
    ctx->last_group = G;
    return 0;
}

} // namespace

/* ================================================================ extern "C" ================================ */

extern "C" {

int srcnn_b200_abi_version(void) { return SRCNN_B200_ABI_VERSION; }

int srcnn_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int srcnn_b200_create(int device, srcnn_b200_ctx** out)
{
    if (!out) return SRCNN_B200_EINVAL;
    *out = nullptr;
    int n = srcnn_b200_device_count();
    if (n <= 0 || device < 0 || device >= n) return SRCNN_B200_ENODEV;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return SRCNN_B200_ENODEV; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { cudaGetLastError(); return SRCNN_B200_ENODEV; }
    if (prop.major != 10) {     /* the kernels are sm_100a only: no other architecture, no CPU fallback */
        fprintf(stderr, "libsrcnn-b200: device %d is sm_%d%d, this build needs sm_100 (B200)\n", device, prop.major, prop.minor);
        return SRCNN_B200_ENODEV;
    }
    srcnn_b200_ctx* ctx = new srcnn_b200_ctx;
    ctx->device = device; ctx->num_sms = prop.multiProcessorCount;
    if (const char* e = getenv("SRCNN_B200_CHAIN")) ctx->chain_tensor = (strcmp(e, "tensor") == 0);
    if (const char* e = getenv("SRCNN_B200_STRIP_MB")) { const double mb = atof(e); if (mb >= 0.25) ctx->strip_bytes = (size_t)(mb * 1048576.0); }
    if (const char* e = getenv("SRCNN_B200_STAGE")) ctx->stage_pageable = atoi(e) != 0;
    if (const char* e = getenv("SRCNN_B200_FUSE")) ctx->fuse_front = atoi(e) != 0;
    if (const char* e = getenv("SRCNN_B200_FUSE_COLOR")) ctx->fuse_color = atoi(e) != 0;
    if (const char* e = getenv("SRCNN_B200_CHROMA_UP")) ctx->chroma_up = atoi(e) != 0;
    if (const char* e = getenv("SRCNN_B200_GROUP")) ctx->group_frames = atoi(e) != 0;
    if (const char* e = getenv("SRCNN_B200_PREP2")) ctx->prep2 = atoi(e) != 0;
    if (const char* e = getenv("SRCNN_B200_STAGE_IN")) ctx->stage_in = atoi(e) != 0;
    if (const char* e = getenv("SRCNN_B200_PREFAULT")) ctx->prefault = atoi(e) != 0;
    if (const char* e = getenv("SRCNN_B200_GRAPH")) ctx->graphs = atoi(e) != 0;
    if (const char* e = getenv("SRCNN_B200_RAMP")) ctx->ramp = atoi(e) != 0;
    if (const char* e = getenv("SRCNN_B200_STAGE_ALL")) ctx->stage_all = atoi(e) != 0;
    if (const char* e = getenv("SRCNN_B200_TRACE")) ctx->trace = atoi(e) != 0;
    {
        unsigned nt = std::min(6u, std::max(1u, std::thread::hardware_concurrency() / 2));
        if (const char* e = getenv("SRCNN_B200_COPY_THREADS")) nt = (unsigned)std::max(1, atoi(e));
        ctx->copier.start(nt - 1);
        ctx->toucher.start(std::min(3u, std::max(1u, nt / 2)));
    }
    bool ok = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->stream_in, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->stream_out, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; ok && i < 8; i++) ok = cudaEventCreate(&ctx->ev[i]) == cudaSuccess;
    for (int i = 0; ok && i < 2; i++)
        ok = cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ev_k[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ev_out[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ev_ak[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ev_ao[i], cudaEventDisableTiming) == cudaSuccess;
    if (ok) ok = conv_exact_init() == 0;
    if (ok) ok = conv_tc_create(&ctx->tc) == 0;
    if (!ok) {
        fprintf(stderr, "libsrcnn-b200: context creation failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        srcnn_b200_destroy(ctx);
        return SRCNN_B200_ENODEV;
    }
    *out = ctx;
    return 0;
}

void srcnn_b200_destroy(srcnn_b200_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto& g : ctx->graph_cache) if (g.exec) cudaGraphExecDestroy(g.exec);
    for (auto& kv : ctx->tables) { cudaFree(kv.second->dev); delete kv.second; }
    if (ctx->tc) conv_tc_destroy(ctx->tc);
    for (int i = 0; i < 2; i++) {
        ctx->src8[i].release(); ctx->out8[i].release(); ctx->conv8[i].release(); ctx->chain[i].release();
        ctx->pin_in[i].release(); ctx->pin_out[i].release(); ctx->pin_conv[i].release();
        if (ctx->ev_in[i]) cudaEventDestroy(ctx->ev_in[i]);
        if (ctx->ev_k[i]) cudaEventDestroy(ctx->ev_k[i]);
        if (ctx->ev_out[i]) cudaEventDestroy(ctx->ev_out[i]);
        if (ctx->ev_ak[i]) cudaEventDestroy(ctx->ev_ak[i]);
        if (ctx->ev_ao[i]) cudaEventDestroy(ctx->ev_ao[i]);
    }
    ctx->chroma_src.release();
    ctx->planes_src.release(); ctx->tmp.release(); ctx->planes_dst.release(); ctx->y16.release(); ctx->yprime.release();
    for (int i = 0; i < 8; i++) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    for (cudaEvent_t e : ctx->pipe_ev) cudaEventDestroy(e);
    ctx->copier.stop(); ctx->toucher.stop();
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->stream_in) cudaStreamDestroy(ctx->stream_in);
    if (ctx->stream_out) cudaStreamDestroy(ctx->stream_out);
    delete ctx;
}

const char* srcnn_b200_last_error(const srcnn_b200_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context"; }

int srcnn_b200_set_conv_mode(srcnn_b200_ctx* ctx, int mode)
{
    if (!ctx || (mode != SRCNN_B200_CONV_TENSOR && mode != SRCNN_B200_CONV_EXACT)) return SRCNN_B200_EINVAL;
    ctx->conv_mode = mode; return 0;
}
int srcnn_b200_get_conv_mode(const srcnn_b200_ctx* ctx) { return ctx ? ctx->conv_mode : SRCNN_B200_EINVAL; }

void srcnn_b200_output_dims(unsigned w, unsigned h, float multiply, unsigned* ow, unsigned* oh)
{
    /* unsigned = unsigned * float: evaluated in float, truncated (src/libsrcnn.cpp:662-663) */
    const float fw = (float)w * multiply, fh = (float)h * multiply;
    if (ow) *ow = fw > 0.f ? (unsigned)fw : 0u;
    if (oh) *oh = fh > 0.f ? (unsigned)fh : 0u;
}

/* step-scale schedule (src/libsrcnn.cpp:980-1061): factors[] of the passes that actually run */
/* *conv_pass: only the pass with cnt + 1 == repeat receives the caller's convbuff (:1024-1028); when the loop breaks before it
 * (remaining factor exactly 1, e.g. x8 = x2 x2 x2 then x1) the reference never produces a convolution plane */
static int step_schedule(unsigned w, float multiply, float* factors, int maxn, int* conv_pass = nullptr)
{
    const float lf = fmodf(multiply, 2.f);
    int repeat = (int)(multiply / 2.f);
    if (lf > 0.f) repeat++;
    unsigned sw = w; int n = 0;
    if (conv_pass) *conv_pass = 0;
    for (int cnt = 0; cnt < repeat && n < maxn; cnt++) {
        float m = 2.0f;
        if (cnt + 1 == repeat) {
            m = ((float)w * multiply) / (float)sw;
            if (m == 0.f || m == 1.0f) break;
            if (conv_pass) *conv_pass = 1;
        }
        factors[n++] = m;
        if (repeat > 1) sw = (unsigned)((float)sw * m);
    }
    return n;
}

int srcnn_b200_step_factors(unsigned w, float multiply, float* factors, int maxn)
{
    if (!factors || maxn <= 0) return SRCNN_B200_EINVAL;
    return step_schedule(w, multiply, factors, maxn);
}

int srcnn_b200_process_has_conv(unsigned w, float multiply, int stepscale)
{
    if (!stepscale) return 1;
    float f[64]; int cp = 0;
    step_schedule(w, multiply, f, 64, &cp);
    return cp;
}

int srcnn_b200_process_dims(unsigned w, unsigned h, float multiply, int stepscale, unsigned* ow, unsigned* oh)
{
    unsigned cw = w, ch = h;
    if (!stepscale) { srcnn_b200_output_dims(w, h, multiply, &cw, &ch); if (ow) *ow = cw; if (oh) *oh = ch; return 1; }
    float f[64];
    const int n = step_schedule(w, multiply, f, 64);
    for (int i = 0; i < n; i++) {
        /* output of a pass from doSRCNN's own size rule; the loop's running sw/sh (:1047-1051) use the same
         * float product when more than one pass runs */
        unsigned nw, nh; srcnn_b200_output_dims(cw, ch, f[i], &nw, &nh); cw = nw; ch = nh;
    }
    if (ow) *ow = n ? cw : 0; if (oh) *oh = n ? ch : 0;
    return n;
}

int srcnn_b200_strip_geometry(unsigned w, unsigned h, unsigned d, float multiply, int filter,
                              unsigned row0, unsigned row1, unsigned* halo_row0, unsigned* halo_row1,
                              unsigned* src_row0, unsigned* src_row1)
{
    Plan p; int rc = make_plan(w, h, d, multiply, filter, row0, row1, p);
    if (rc) return rc;
    if (p.need_v) {
        AxisTable vy, vc;
        build_axis_table(p.fy, p.oh, p.h, vy);
        build_axis_table(p.fc, p.oh, p.h, vc);
        source_rows(vy, vc, p.e0, p.e1, p.s0, p.s1);
    } else { p.s0 = p.e0; p.s1 = p.e1; }
    if (halo_row0) *halo_row0 = p.e0; if (halo_row1) *halo_row1 = p.e1;
    if (src_row0) *src_row0 = p.s0; if (src_row1) *src_row1 = p.s1;
    return 0;
}

int srcnn_b200_pass_device(srcnn_b200_ctx* ctx, const void* d_src, unsigned w, unsigned h, unsigned d,
                           float multiply, int filter, unsigned row0, unsigned row1, void* d_out, void* d_conv)
{
    if (!ctx || !d_src || !d_out) return SRCNN_B200_EINVAL;
    Plan p; int rc = make_plan(w, h, d, multiply, filter, row0, row1, p);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    const uint8_t* src8 = static_cast<const uint8_t*>(d_src);
    uint8_t *out8 = static_cast<uint8_t*>(d_out), *conv8 = static_cast<uint8_t*>(d_conv);
    if (!ctx->graphs || ctx->timing) return run_pass(ctx, p, src8, 0, out8, conv8);

    /* everything the launch sequence depends on: the call's arguments, the mode switches, and the context's scratch buffers
     * (a grown buffer has a new address: the entry then no longer matches and the pass is captured again) */
    uint32_t mbits; memcpy(&mbits, &multiply, 4);
    auto key_now = [&] {
        return std::vector<unsigned long long>{
            (unsigned long long)(uintptr_t)d_src, (unsigned long long)(uintptr_t)d_out, (unsigned long long)(uintptr_t)d_conv,
            ((unsigned long long)w << 32) | h, ((unsigned long long)d << 32) | mbits, ((unsigned long long)row0 << 32) | row1,
            ((unsigned long long)(unsigned)filter << 32) | ((unsigned)ctx->conv_mode << 8) | ((unsigned)ctx->keep_planes << 4) |
                ((unsigned)ctx->fuse_front << 3) | ((unsigned)ctx->chroma_up << 2) | ((unsigned)ctx->prep2 << 1) | (unsigned)ctx->fuse_color,
            (unsigned long long)(uintptr_t)ctx->y16.p, (unsigned long long)(uintptr_t)ctx->yprime.p,
            (unsigned long long)(uintptr_t)ctx->chroma_src.p, (unsigned long long)(uintptr_t)ctx->planes_dst.p,
            (unsigned long long)(uintptr_t)ctx->planes_src.p, (unsigned long long)(uintptr_t)ctx->tmp.p };
    };
    std::vector<unsigned long long> key = key_now();
    srcnn_b200_ctx::GraphEntry* e = nullptr;
    for (auto& g : ctx->graph_cache) if (g.key == key) { e = &g; break; }
    if (e && e->exec) {
        CK(cudaGraphLaunch(e->exec, ctx->stream));
        ctx->graph_replays++; ctx->launches += 3;
        ctx->last = p; ctx->have_last = true; ctx->last_group = 1;
        return 0;
    }
    if (!e) {
        /* first sighting: run it plainly (tables are built and uploaded, buffers grow -- none of that may happen inside a capture) */
        rc = run_pass(ctx, p, src8, 0, out8, conv8);
        if (rc) return rc;
        if (ctx->graph_cache.size() >= 16) {                                  /* small cache, oldest entry goes */
            if (ctx->graph_cache.front().exec) cudaGraphExecDestroy(ctx->graph_cache.front().exec);
            ctx->graph_cache.erase(ctx->graph_cache.begin());
        }
        srcnn_b200_ctx::GraphEntry ne; ne.key = key_now(); ne.seen = 1;       /* (the key AFTER the run: buffers may have grown) */
        ctx->graph_cache.push_back(ne);
        return 0;
    }
    /* second sighting: capture the launch sequence, instantiate, replay */
    cudaGraph_t graph = nullptr;
    CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    rc = run_pass(ctx, p, src8, 0, out8, conv8);
    cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
    if (rc || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        e->seen = 0; e->key.clear();                                          /* not capturable: never try this entry again */
        ctx->graphs = false;
        return rc ? rc : run_pass(ctx, p, src8, 0, out8, conv8);
    }
    ce = cudaGraphInstantiate(&e->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) { e->exec = nullptr; cudaGetLastError(); ctx->graphs = false; return run_pass(ctx, p, src8, 0, out8, conv8); }
    CK(cudaGraphLaunch(e->exec, ctx->stream));
    ctx->graph_replays++;
    return 0;
}

int srcnn_b200_set_graphs(srcnn_b200_ctx* ctx, int on)
{
    if (!ctx) return SRCNN_B200_EINVAL;
    ctx->graphs = on != 0;
    return 0;
}
unsigned long long srcnn_b200_graph_replays(const srcnn_b200_ctx* ctx) { return ctx ? ctx->graph_replays : 0; }

int srcnn_b200_sync(srcnn_b200_ctx* ctx)
{
    if (!ctx) return SRCNN_B200_EINVAL;
    CK(cudaStreamSynchronize(ctx->stream));
    collect_timing(ctx);
    return 0;
}
void* srcnn_b200_stream(srcnn_b200_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

static inline size_t conv_bytes_of(unsigned rows, unsigned ow, const void* conv) { return conv ? (size_t)rows * ow : 0; }

/* An error in the middle of a host pipeline leaves kernels and asynchronous copies of earlier sub-strips in flight -- to the
 * caller's buffers (which it is about to free) and to the context's staging slots (which the next call reuses).  Drain all
 * three streams before the error is reported, and clear the sticky error state. */
static void drain_streams(srcnn_b200_ctx* ctx)
{
    ctx->toucher.wait();
    cudaStreamSynchronize(ctx->stream_in);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->stream_out);
    cudaGetLastError();
}

static int pass_host_rows_impl(srcnn_b200_ctx* ctx, const uint8_t* src, unsigned w, unsigned h, unsigned d,
                               float multiply, int filter, unsigned row0, unsigned row1,
                               uint8_t* out_rows, uint8_t* conv_rows, bool async = false)
{
    if (!ctx || !src || !out_rows) return SRCNN_B200_EINVAL;
    const std::chrono::steady_clock::time_point t_entry = std::chrono::steady_clock::now();
    Plan p; int rc = make_plan(w, h, d, multiply, filter, row0, row1, p);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    TableDev *vy = nullptr, *vc = nullptr;
    if (p.need_v) {
        if ((rc = get_table(ctx, p.fy, p.oh, p.h, &vy))) return rc;
        if ((rc = get_table(ctx, p.fc, p.oh, p.h, &vc))) return rc;
    }
    /* The call is cut into row sub-strips that flow through a three-stage pipeline
     *     upload source rows (stream_in) | kernels (stream) | download result rows (stream_out)
     * so the PCIe transfers of one sub-strip hide behind the kernels of its neighbours.  Sub-strips are ordinary
     * row strips (6-row halo recomputed), hence bit-identical to the one-shot result.  The whole source row range
     * and the whole result live in device buffers: overlapping source rows are uploaded once, nothing is reused
     * while in flight. */
    const size_t row_bytes = (size_t)p.ow * d;
    const size_t out_bytes = (size_t)(row1 - row0) * row_bytes, conv_bytes = (size_t)(row1 - row0) * p.ow;
    /* sub-strip size: 4 MB of result; a pageable destination too large to be staged whole (below) uses 8 MB (measured,
     * tools/e2e_probe.py) */
    const bool out_pinned = host_is_pinned(out_rows);
    constexpr size_t kStageAllBytes = (size_t)96 << 20;
    const bool stage_all = ctx->stage_all && !out_pinned && out_bytes + conv_bytes_of(row1 - row0, p.ow, conv_rows) <= kStageAllBytes;
    const size_t strip_bytes = ctx->strip_bytes ? ctx->strip_bytes : (size_t)((out_pinned || stage_all) ? 4 : 8) << 20;
    unsigned nstrips = (unsigned)std::min<size_t>((out_bytes + strip_bytes / 2) / strip_bytes, kMaxPipeStrips);
    nstrips = std::max(1u, std::min(nstrips, (row1 - row0) / 64u));
    /* ramp: the first two sub-strips are a quarter and a half of the others, so the first download -- the transfers are what
     * bounds the call -- starts after ~1 MB of work instead of ~4 MB (cuts are free: every sub-strip is an ordinary row strip) */
    const bool ramp = ctx->ramp && nstrips >= 3 && nstrips < kMaxPipeStrips && (row1 - row0) / (nstrips + 1) >= 256u;
    if (ramp) nstrips++;
    auto cut = [&](unsigned i) -> unsigned {                       /* first row of sub-strip i (i == nstrips: the end) */
        if (!ramp) return row0 + (unsigned)((size_t)(row1 - row0) * i / nstrips);
        const double total = (double)nstrips - 1.25;               /* weights 0.25, 0.5, 1, 1, ... */
        const double upto = i == 0 ? 0.0 : i == 1 ? 0.25 : 0.75 + (double)(i - 2);
        const unsigned r = row0 + (unsigned)((double)(row1 - row0) * (upto / total) + 0.5);
        return i >= nstrips ? row1 : std::min(r, row1);
    };
    Plan sp[kMaxPipeStrips];
    for (unsigned i = 0; i < nstrips; i++) {
        const unsigned a = cut(i), b = cut(i + 1);
        if ((rc = make_plan(w, h, d, multiply, filter, a, b, sp[i]))) return rc;
        if (p.need_v) source_rows(vy->host, vc->host, sp[i].e0, sp[i].e1, sp[i].s0, sp[i].s1);
        else { sp[i].s0 = sp[i].e0; sp[i].s1 = sp[i].e1; }
    }
    const unsigned S0 = sp[0].s0, S1 = sp[nstrips - 1].s1;
    const size_t src_row_bytes = (size_t)w * d;
    /* asynchronous calls alternate between two device buffer sets, so call k+1 uploads while call k computes / downloads;
     * the blocking call always uses set 0 (and has drained everything before it returns) */
    const int bs = async ? (int)(ctx->async_seq & 1) : 0;
    if (async && ctx->async_used[bs]) {
        /* set `bs` was last used two calls ago: its kernels must have read the source, its downloads must have left */
        CK(cudaStreamWaitEvent(ctx->stream_in, ctx->ev_ak[bs], 0));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_ao[bs], 0));
    }
    /* (a growing buffer is re-allocated: cudaFree waits for the device, so earlier asynchronous work is not disturbed) */
    CK(ctx->src8[bs].ensure((size_t)(S1 - S0) * src_row_bytes));
    CK(ctx->out8[bs].ensure(out_bytes));
    if (conv_rows) CK(ctx->conv8[bs].ensure(conv_bytes));
    uint8_t* d_src = static_cast<uint8_t*>(ctx->src8[bs].p);
    uint8_t* d_out = static_cast<uint8_t*>(ctx->out8[bs].p);
    uint8_t* d_conv = conv_rows ? static_cast<uint8_t*>(ctx->conv8[bs].p) : nullptr;

    /* pageable destination: DMA into pinned staging slots, worker threads move the rows on (a cudaMemcpy straight
     * into pageable memory is staged by the driver in small chunks and blocks the pipeline) */
    const bool stage = !async && ctx->stage_pageable && nstrips > 1 && !out_pinned;
    /* results of up to 96 MB are staged whole (one slot per sub-strip): the GPU side of the call is then enqueued in one go
     * and never waits for the host's copies; larger results cycle through three slots */
    const unsigned nslots = (stage && stage_all) ? nstrips : kStageSlots;
    size_t slot_bytes = 0;
    if (stage) {
        for (unsigned i = 0; i < nstrips; i++)
            slot_bytes = std::max(slot_bytes, (size_t)(sp[i].r1 - sp[i].r0) * (row_bytes + (conv_rows ? p.ow : 0)));
        slot_bytes = (slot_bytes + 4095) / 4096 * 4096;
        CK(ctx->pin_out[0].ensure(slot_bytes * nslots));
    }
    /* pageable source: the rows of a sub-strip are copied into pinned memory by the worker threads and uploaded from there
     * (a cudaMemcpyAsync from pageable memory is staged by the driver, synchronously and at a third of the DMA rate) */
    /* (only when the result is staged as well: with a pinned destination the driver's own path is faster, measured) */
    const bool stage_in = stage && ctx->stage_in && !host_is_pinned(src);
    if (stage_in) CK(ctx->pin_in[0].ensure((size_t)(S1 - S0) * src_row_bytes));
    /* events of the sub-strip pipeline; an asynchronous call on buffer set 1 uses the second half of the event pool */
    const unsigned ev0 = (unsigned)bs * 3u * kMaxPipeStrips;
    CK(ensure_pipe_events(ctx, 2 * kMaxPipeStrips));

    auto drain = [&](unsigned j) -> int {      /* staging slot j % nslots -> caller's rows */
        CK(cudaEventSynchronize(ctx->pipe_ev[ev0 + 3 * j + 2]));
        ctx->toucher.wait();                   /* (the background first touch writes into the same pages) */
        const unsigned char* slot = static_cast<unsigned char*>(ctx->pin_out[0].p) + (size_t)(j % nslots) * slot_bytes;
        const size_t ob = (size_t)(sp[j].r1 - sp[j].r0) * row_bytes, off = (size_t)(sp[j].r0 - row0);
        ctx->copier.copy(out_rows + off * row_bytes, slot, ob);
        if (conv_rows) ctx->copier.copy(conv_rows + off * p.ow, slot + ob, (size_t)(sp[j].r1 - sp[j].r0) * p.ow);
        return 0;
    };

    using clk = std::chrono::steady_clock;
    const clk::time_point t_enq = clk::now();
    /* a fresh result buffer (ProcessSRCNN hands over new[] memory): first touch in the background, joined by the first copy */
    if (stage && ctx->prefault) {
        ctx->toucher.touch_async(out_rows, out_bytes);                 /* (its own threads: the copier stages source rows meanwhile) */
    }
    unsigned uploaded = S0;
    for (unsigned i = 0; i < nstrips; i++) {
        cudaEvent_t ev_in = ctx->pipe_ev[ev0 + 3 * i], ev_k = ctx->pipe_ev[ev0 + 3 * i + 1], ev_out = ctx->pipe_ev[ev0 + 3 * i + 2];
        if (sp[i].s1 > uploaded) {
            const size_t o = (size_t)(uploaded - S0) * src_row_bytes, n = (size_t)(sp[i].s1 - uploaded) * src_row_bytes;
            const uint8_t* from = src + (size_t)uploaded * src_row_bytes;
            if (stage_in) {
                uint8_t* pin = static_cast<uint8_t*>(ctx->pin_in[0].p) + o;
                ctx->copier.copy(pin, from, n);
                from = pin;
            }
            CK(cudaMemcpyAsync(d_src + o, from, n, cudaMemcpyHostToDevice, ctx->stream_in));
            uploaded = sp[i].s1;
        }
        CK(cudaEventRecord(ev_in, ctx->stream_in));
        CK(cudaStreamWaitEvent(ctx->stream, ev_in, 0));
        const size_t off = (size_t)(sp[i].r0 - row0);
        rc = run_pass(ctx, sp[i], d_src, (long)S0, d_out + off * row_bytes, d_conv ? d_conv + off * p.ow : nullptr);
        if (rc) return rc;
        CK(cudaEventRecord(ev_k, ctx->stream));
        CK(cudaStreamWaitEvent(ctx->stream_out, ev_k, 0));
        const size_t ob = (size_t)(sp[i].r1 - sp[i].r0) * row_bytes, cb = (size_t)(sp[i].r1 - sp[i].r0) * p.ow;
        if (stage) {
            if (i >= nslots && (rc = drain(i - nslots))) return rc;      /* slot free again */
            unsigned char* slot = static_cast<unsigned char*>(ctx->pin_out[0].p) + (size_t)(i % nslots) * slot_bytes;
            CK(cudaMemcpyAsync(slot, d_out + off * row_bytes, ob, cudaMemcpyDeviceToHost, ctx->stream_out));
            if (conv_rows) CK(cudaMemcpyAsync(slot + ob, d_conv + off * p.ow, cb, cudaMemcpyDeviceToHost, ctx->stream_out));
        } else {
            CK(cudaMemcpyAsync(out_rows + off * row_bytes, d_out + off * row_bytes, ob, cudaMemcpyDeviceToHost, ctx->stream_out));
            if (conv_rows) CK(cudaMemcpyAsync(conv_rows + off * p.ow, d_conv + off * p.ow, cb, cudaMemcpyDeviceToHost, ctx->stream_out));
        }
        CK(cudaEventRecord(ev_out, ctx->stream_out));
    }
    if (async) {
        /* nothing is waited for: the caller's buffers are in use until srcnn_b200_wait */
        CK(cudaEventRecord(ctx->ev_ak[bs], ctx->stream));
        CK(cudaEventRecord(ctx->ev_ao[bs], ctx->stream_out));
        ctx->async_used[bs] = true;
        ctx->async_seq++;
        return 0;
    }
    const clk::time_point t_drain = clk::now();
    if (stage)
        for (unsigned j = nstrips > nslots ? nstrips - nslots : 0; j < nstrips; j++)
            if ((rc = drain(j))) return rc;
    CK(cudaStreamSynchronize(ctx->stream_out));
    if (ctx->trace) {
        const clk::time_point t_end = clk::now();
        auto us = [](clk::time_point a, clk::time_point b) { return (double)std::chrono::duration_cast<std::chrono::nanoseconds>(b - a).count() * 1e-3; };
        fprintf(stderr, "srcnn trace: %u sub-strips, stage_out %d (%u slots) stage_in %d | set-up %.0f us, enqueue %.0f us, drain+tail %.0f us, total %.0f us\n",
                nstrips, (int)stage, nslots, (int)stage_in, us(t_entry, t_enq), us(t_enq, t_drain), us(t_drain, t_end), us(t_entry, t_end));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    collect_timing(ctx);
    return 0;                                 /* (the plane getters describe the last sub-strip) */
}

int srcnn_b200_pass_host_rows(srcnn_b200_ctx* ctx, const uint8_t* src, unsigned w, unsigned h, unsigned d,
                              float multiply, int filter, unsigned row0, unsigned row1,
                              uint8_t* out_rows, uint8_t* conv_rows)
{
    if (ctx && (ctx->async_used[0] || ctx->async_used[1])) { const int w0 = srcnn_b200_wait(ctx); if (w0) return w0; }
    const int rc = pass_host_rows_impl(ctx, src, w, h, d, multiply, filter, row0, row1, out_rows, conv_rows);
    if (rc && ctx) drain_streams(ctx);
    return rc;
}

/* Stream-ordered, non-blocking variant: uploads, kernels and downloads are only ENQUEUED; src / out / conv
 * must stay valid -- and should be pinned (srcnn_b200_host_alloc), a copy from / to pageable memory makes the call wait --
 * until srcnn_b200_wait returns.  Up to two calls are in flight per context (two device buffer sets): the upload of call k+1
 * overlaps the kernels and the download of call k. */
int srcnn_b200_pass_host_async(srcnn_b200_ctx* ctx, const uint8_t* src, unsigned w, unsigned h, unsigned d,
                               float multiply, int filter, uint8_t* out, uint8_t* conv)
{
    unsigned ow = 0, oh = 0;
    if (!ctx || !w || !h || (d != 3 && d != 4)) return SRCNN_B200_EINVAL;
    if ((float)w * multiply <= 0.f || (float)h * multiply <= 0.f) return SRCNN_B200_ESIZE;
    srcnn_b200_output_dims(w, h, multiply, &ow, &oh);
    if (!ow || !oh) return SRCNN_B200_ESIZE;
    const int rc = pass_host_rows_impl(ctx, src, w, h, d, multiply, filter, 0, oh, out, conv, true);
    if (rc) { drain_streams(ctx); ctx->async_used[0] = ctx->async_used[1] = false; }
    return rc;
}

int srcnn_b200_wait(srcnn_b200_ctx* ctx)
{
    if (!ctx) return SRCNN_B200_EINVAL;
    CK(cudaSetDevice(ctx->device));
    cudaError_t e = cudaStreamSynchronize(ctx->stream_in);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream_out);
    ctx->async_used[0] = ctx->async_used[1] = false;
    if (e != cudaSuccess) { drain_streams(ctx); return fail_cuda(ctx, e, "srcnn_b200_wait"); }
    collect_timing(ctx);
    return 0;
}

int srcnn_b200_pass_host(srcnn_b200_ctx* ctx, const uint8_t* src, unsigned w, unsigned h, unsigned d,
                         float multiply, int filter, uint8_t* out, uint8_t* conv)
{
    unsigned ow = 0, oh = 0;
    if (!w || !h || (d != 3 && d != 4)) return SRCNN_B200_EINVAL;
    if ((float)w * multiply <= 0.f || (float)h * multiply <= 0.f) return SRCNN_B200_ESIZE;
    srcnn_b200_output_dims(w, h, multiply, &ow, &oh);
    if (!ow || !oh) return SRCNN_B200_ESIZE;
    return srcnn_b200_pass_host_rows(ctx, src, w, h, d, multiply, filter, 0, oh, out, conv);
}

static int process_host_impl(srcnn_b200_ctx* ctx, const uint8_t* src, unsigned w, unsigned h, unsigned d,
                             float multiply, int filter, int stepscale, uint8_t* out, uint8_t* conv)
{
    if (!ctx || !src || !w || !h || !d) return SRCNN_B200_EINVAL;
    if (d != 3 && d != 4) return SRCNN_B200_EINVAL;
    if ((float)w * multiply <= 0.f || (float)h * multiply <= 0.f) return SRCNN_B200_ESIZE;
    if (!stepscale) return srcnn_b200_pass_host(ctx, src, w, h, d, multiply, filter, out, conv);

    float f[64]; int conv_pass = 0;
    const int n = step_schedule(w, multiply, f, 64, &conv_pass);
    if (n == 0) return SRCNN_B200_ENOTHING;
    if (!conv_pass) conv = nullptr;           /* the reference produces no convolution plane for this factor: `conv` stays untouched */
    if (n == 1) return srcnn_b200_pass_host(ctx, src, w, h, d, f[0], filter, out, conv);
    if (!out) return SRCNN_B200_EINVAL;
    /* chain: 8-bit RGB(A) intermediates stay in HBM (the reference round-trips them through host buffers,
     * src/libsrcnn.cpp:999-1051; the quantisation to 8 bit between passes is preserved) */
    CK(cudaSetDevice(ctx->device));
    unsigned cw = w, ch = h;
    CK(ctx->chain[0].ensure((size_t)w * h * d));
    CK(cudaMemcpyAsync(ctx->chain[0].p, src, (size_t)w * h * d, cudaMemcpyHostToDevice, ctx->stream));
    int cur = 0;
    for (int i = 0; i < n; i++) {
        Plan p; int rc = make_plan(cw, ch, d, f[i], filter, 0, 1, p);
        if (rc) return rc;
        rc = make_plan(cw, ch, d, f[i], filter, 0, p.oh, p);
        if (rc) return rc;
        const bool last = (i + 1 == n);
        /* An intermediate pass feeds the next one through an 8-bit image: a 1-LSB difference there would be
         * amplified by the following pass, so unless asked otherwise intermediates use the bit-exact kernel and
         * only the last pass (whose output is what the 1-LSB criterion is about) uses the tensor-core kernel. */
        const int saved_mode = ctx->conv_mode;
        if (!last && !ctx->chain_tensor) ctx->conv_mode = SRCNN_B200_CONV_EXACT;
        const size_t ob = (size_t)p.ow * p.oh * d;
        { cudaError_t e1 = ctx->chain[cur ^ 1].ensure(ob);
          if (e1 == cudaSuccess && last && conv) e1 = ctx->conv8[0].ensure((size_t)p.ow * p.oh);
          if (e1 != cudaSuccess) { ctx->conv_mode = saved_mode; return fail_cuda(ctx, e1, "chain buffers"); } }
        rc = run_pass(ctx, p, static_cast<const uint8_t*>(ctx->chain[cur].p), 0,
                      static_cast<uint8_t*>(ctx->chain[cur ^ 1].p),
                      (last && conv) ? static_cast<uint8_t*>(ctx->conv8[0].p) : nullptr);
        ctx->conv_mode = saved_mode;
        if (rc) return rc;
        cur ^= 1; cw = p.ow; ch = p.oh;
    }
    CK(cudaMemcpyAsync(out, ctx->chain[cur].p, (size_t)cw * ch * d, cudaMemcpyDeviceToHost, ctx->stream));
    if (conv) CK(cudaMemcpyAsync(conv, ctx->conv8[0].p, (size_t)cw * ch, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    collect_timing(ctx);
    return 0;
}

int srcnn_b200_process_host(srcnn_b200_ctx* ctx, const uint8_t* src, unsigned w, unsigned h, unsigned d,
                            float multiply, int filter, int stepscale, uint8_t* out, uint8_t* conv)
{
    if (ctx && (ctx->async_used[0] || ctx->async_used[1])) { const int w0 = srcnn_b200_wait(ctx); if (w0) return w0; }
    const int rc = process_host_impl(ctx, src, w, h, d, multiply, filter, stepscale, out, conv);
    if (rc && rc != SRCNN_B200_ENOTHING && ctx) drain_streams(ctx);
    return rc;
}

/* frames [f0, f0+G) of a batch, device-resident: the grouped route when it applies, run_pass per frame otherwise */
static int batch_frames_device(srcnn_b200_ctx* ctx, unsigned w, unsigned h, unsigned d, float multiply, int filter, unsigned G,
                               const uint8_t* d_src, uint8_t* d_out, uint8_t* d_conv)
{
    Plan p; int rc = make_plan(w, h, d, multiply, filter, 0, 1, p);
    if (rc) return rc;
    if ((rc = make_plan(w, h, d, multiply, filter, 0, p.oh, p))) return rc;
    const size_t ib = (size_t)w * h * d, ob = (size_t)p.ow * p.oh * d, cb = (size_t)p.ow * p.oh;
    if (G > 1) {
        rc = run_group(ctx, p, G, d_src, d_out, d_conv);
        if (rc <= 0) return rc;                                     /* done, or a real error; 1 = not covered */
    }
    for (unsigned f = 0; f < G; f++) {
        Plan q; make_plan(w, h, d, multiply, filter, 0, p.oh, q);
        if ((rc = run_pass(ctx, q, d_src + f * ib, 0, d_out + f * ob, d_conv ? d_conv + f * cb : nullptr))) return rc;
    }
    return 0;
}

int srcnn_b200_batch_device(srcnn_b200_ctx* ctx, const void* d_frames, unsigned n, unsigned w, unsigned h,
                            unsigned d, float multiply, int filter, void* d_out, void* d_conv)
{
    if (!ctx || !d_frames || !d_out) return SRCNN_B200_EINVAL;
    if (n == 0) return 0;
    Plan p0; int rc = make_plan(w, h, d, multiply, filter, 0, 1, p0);
    if (rc) return rc;
    if ((rc = make_plan(w, h, d, multiply, filter, 0, p0.oh, p0))) return rc;
    CK(cudaSetDevice(ctx->device));
    const size_t ib = (size_t)w * h * d, ob = (size_t)p0.ow * p0.oh * d, cb = (size_t)p0.ow * p0.oh;
    const unsigned G = group_size(ctx, p0, n);
    for (unsigned f0 = 0; f0 < n; f0 += G) {
        const unsigned g = std::min(G, n - f0);
        rc = batch_frames_device(ctx, w, h, d, multiply, filter, g, static_cast<const uint8_t*>(d_frames) + f0 * ib,
                                 static_cast<uint8_t*>(d_out) + f0 * ob, d_conv ? static_cast<uint8_t*>(d_conv) + f0 * cb : nullptr);
        if (rc) return rc;
    }
    return 0;
}

static int batch_host_impl(srcnn_b200_ctx* ctx, const uint8_t* frames, unsigned n, unsigned w, unsigned h,
                           unsigned d, float multiply, int filter, uint8_t* out, uint8_t* conv)
{
    if (!ctx || !frames || !out) return SRCNN_B200_EINVAL;
    if (n == 0) return 0;
    Plan p0; int rc = make_plan(w, h, d, multiply, filter, 0, 1, p0);
    if (rc) return rc;
    if ((rc = make_plan(w, h, d, multiply, filter, 0, p0.oh, p0))) return rc;
    CK(cudaSetDevice(ctx->device));
    const size_t ib = (size_t)w * h * d, ob = (size_t)p0.ow * p0.oh * d, cb = (size_t)p0.ow * p0.oh;
    /* frames per pipeline step (1 for large frames).  The host pipeline is bound by the download, and its fill / drain
     * grow with the step size: at most 4 frames per step here (16 for device-resident batches) */
    const unsigned G = std::min(4u, group_size(ctx, p0, n));
    for (int i = 0; i < 2; i++) {
        CK(ctx->src8[i].ensure(ib * G)); CK(ctx->out8[i].ensure(ob * G));
        if (conv) CK(ctx->conv8[i].ensure(cb * G));
    }
    /* three-stage software pipeline over frame groups: upload (stream_in) | kernels (stream) | download (stream_out),
     * double-buffered device input/output so group i+1 uploads and group i-1 downloads while group i computes */
    unsigned i = 0;
    for (unsigned f0 = 0; f0 < n; f0 += G, i++) {
        const unsigned g = std::min(G, n - f0);
        const int b = (int)(i & 1);
        if (i >= 2) CK(cudaStreamWaitEvent(ctx->stream_in, ctx->ev_k[b], 0));      /* src8[b] consumed by group i-2 */
        CK(cudaMemcpyAsync(ctx->src8[b].p, frames + (size_t)f0 * ib, ib * g, cudaMemcpyHostToDevice, ctx->stream_in));
        CK(cudaEventRecord(ctx->ev_in[b], ctx->stream_in));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_in[b], 0));
        if (i >= 2) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_out[b], 0));       /* out8[b] drained by group i-2 */
        rc = batch_frames_device(ctx, w, h, d, multiply, filter, g, static_cast<const uint8_t*>(ctx->src8[b].p),
                                 static_cast<uint8_t*>(ctx->out8[b].p), conv ? static_cast<uint8_t*>(ctx->conv8[b].p) : nullptr);
        if (rc) return rc;
        CK(cudaEventRecord(ctx->ev_k[b], ctx->stream));
        CK(cudaStreamWaitEvent(ctx->stream_out, ctx->ev_k[b], 0));
        CK(cudaMemcpyAsync(out + (size_t)f0 * ob, ctx->out8[b].p, ob * g, cudaMemcpyDeviceToHost, ctx->stream_out));
        if (conv) CK(cudaMemcpyAsync(conv + (size_t)f0 * cb, ctx->conv8[b].p, cb * g, cudaMemcpyDeviceToHost, ctx->stream_out));
        CK(cudaEventRecord(ctx->ev_out[b], ctx->stream_out));
    }
    CK(cudaStreamSynchronize(ctx->stream_out));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int srcnn_b200_batch_host(srcnn_b200_ctx* ctx, const uint8_t* frames, unsigned n, unsigned w, unsigned h,
                          unsigned d, float multiply, int filter, uint8_t* out, uint8_t* conv)
{
    if (ctx && (ctx->async_used[0] || ctx->async_used[1])) { const int w0 = srcnn_b200_wait(ctx); if (w0) return w0; }
    const int rc = batch_host_impl(ctx, frames, n, w, h, d, multiply, filter, out, conv);
    if (rc && ctx) drain_streams(ctx);
    return rc;
}

void* srcnn_b200_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void srcnn_b200_host_free(void* p) { if (p) cudaFreeHost(p); }

int srcnn_b200_get_plane(srcnn_b200_ctx* ctx, int which, float* host, size_t cap,
                         unsigned* row0, unsigned* rows, unsigned* width)
{
    if (!ctx || !ctx->have_last || !host || which < 0 || which > 4) return SRCNN_B200_EINVAL;
    const Plan& p = ctx->last;
    if (which == 3 && p.d != 4) return SRCNN_B200_EINVAL;
    if (which == 0 && !ctx->last_has_y32) return SRCNN_B200_EINVAL;      /* see srcnn_b200_set_keep_planes */
    const float* srcp; unsigned r0, nr;
    if (which >= 1 && which <= 3 && !ctx->last_has_chroma) return SRCNN_B200_EINVAL;   /* see srcnn_b200_set_keep_planes */
    if (which == 4 && !ctx->last_has_yp) return SRCNN_B200_EINVAL;       /* see srcnn_b200_set_keep_planes */
    if (which == 4) { srcp = static_cast<const float*>(ctx->yprime.p); r0 = p.r0; nr = p.r1 - p.r0; }
    else { srcp = static_cast<const float*>(ctx->planes_dst.p) + (size_t)which * ctx->last_plane_stride; r0 = p.e0; nr = p.e1 - p.e0; }
    const size_t n = (size_t)nr * p.ow;
    if (cap < n) return SRCNN_B200_EINVAL;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(host, srcp, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (row0) *row0 = r0; if (rows) *rows = nr; if (width) *width = p.ow;
    return 0;
}

int srcnn_b200_set_copy_threads(srcnn_b200_ctx* ctx, unsigned n)
{
    if (!ctx || n == 0 || n > 64) return SRCNN_B200_EINVAL;
    ctx->copier.stop();
    ctx->copier.start(n - 1);
    return 0;
}

int srcnn_b200_set_keep_planes(srcnn_b200_ctx* ctx, int on) { if (!ctx) return SRCNN_B200_EINVAL; ctx->keep_planes = on == 2 ? 2 : (on != 0); return 0; }
int srcnn_b200_set_timing(srcnn_b200_ctx* ctx, int on) { if (!ctx) return SRCNN_B200_EINVAL; ctx->timing = on != 0; return 0; }
int srcnn_b200_last_timings(srcnn_b200_ctx* ctx, float* ms7)
{
    if (!ctx || !ms7) return SRCNN_B200_EINVAL;
    memcpy(ms7, ctx->ms, sizeof(ctx->ms)); return 0;
}
unsigned long long srcnn_b200_launch_count(const srcnn_b200_ctx* ctx) { return ctx ? ctx->launches : 0; }
unsigned srcnn_b200_last_group(const srcnn_b200_ctx* ctx) { return ctx ? ctx->last_group : 0; }

/* ---------------------------------------------------------------- single stages (parity tests) */

int srcnn_b200_stage_color_in(srcnn_b200_ctx* ctx, const uint8_t* src, unsigned w, unsigned h, unsigned d,
                              float* const* planes)
{
    if (!ctx || !src || !planes || !w || !h || (d != 3 && d != 4)) return SRCNN_B200_EINVAL;
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)w * h, ps = (n + 63) / 64 * 64;
    CK(ctx->src8[0].ensure(n * d)); CK(ctx->planes_src.ensure(ps * d * sizeof(float)));
    CK(cudaMemcpyAsync(ctx->src8[0].p, src, n * d, cudaMemcpyHostToDevice, ctx->stream));
    launch_color_in(static_cast<const uint8_t*>(ctx->src8[0].p), n, d, static_cast<float*>(ctx->planes_src.p), ps, ctx->stream);
    ctx->launches++;
    for (unsigned i = 0; i < d; i++)
        CK(cudaMemcpyAsync(planes[i], static_cast<float*>(ctx->planes_src.p) + i * ps, n * sizeof(float),
                           cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    return 0;
}

int srcnn_b200_stage_resize(srcnn_b200_ctx* ctx, const float* src, unsigned sw, unsigned sh,
                            unsigned dw, unsigned dh, int filter, float* dst)
{
    if (!ctx || !src || !dst || !sw || !sh || !dw || !dh) return SRCNN_B200_EINVAL;
    CK(cudaSetDevice(ctx->device));
    const int f = (filter >= kNearest && filter <= kBspline) ? filter : kNearest;
    const size_t ns = (size_t)sw * sh, nd = (size_t)dw * dh;
    const bool hfirst = dw <= sw, need_h = dw != sw, need_v = dh != sh;
    const size_t nt = hfirst ? (size_t)dw * sh : (size_t)sw * dh;
    CK(ctx->planes_src.ensure(ns * sizeof(float))); CK(ctx->planes_dst.ensure(nd * sizeof(float)));
    CK(ctx->tmp.ensure(nt * sizeof(float)));
    float *A = static_cast<float*>(ctx->planes_src.p), *T = static_cast<float*>(ctx->tmp.p), *B = static_cast<float*>(ctx->planes_dst.p);
    CK(cudaMemcpyAsync(A, src, ns * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    TableDev *tv = nullptr, *th = nullptr; int rc;
    if (need_v && (rc = get_table(ctx, f, dh, sh, &tv))) return rc;
    if (need_h && (rc = get_table(ctx, f, dw, sw, &th))) return rc;
    cudaStream_t s = ctx->stream;
    if (!need_h && !need_v) {
        CK(cudaMemcpyAsync(B, A, ns * sizeof(float), cudaMemcpyDeviceToDevice, s));
    } else if (hfirst) {
        if (need_h) { launch_resize_horizontal(A, 0, sw, need_v ? T : B, 0, dw, sh, 1, th->view, th->view, s); ctx->launches++; }
        if (need_v) { launch_resize_vertical(need_h ? T : A, 0, dw, 0, B, 0, 0, dh, 1, tv->view, tv->view, s); ctx->launches++; }
    } else {
        if (need_v) { launch_resize_vertical(A, 0, sw, 0, need_h ? T : B, 0, 0, dh, 1, tv->view, tv->view, s); ctx->launches++; }
        if (need_h) { launch_resize_horizontal(need_v ? T : A, 0, sw, B, 0, dw, dh, 1, th->view, th->view, s); ctx->launches++; }
    }
    CK(cudaMemcpyAsync(dst, B, nd * sizeof(float), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    return 0;
}

int srcnn_b200_stage_conv(srcnn_b200_ctx* ctx, const float* y, unsigned w, unsigned h, int mode, float* out)
{
    if (!ctx || !y || !out || !w || !h) return SRCNN_B200_EINVAL;
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)w * h;
    CK(ctx->planes_dst.ensure(n * sizeof(float))); CK(ctx->yprime.ensure(n * sizeof(float)));
    float* Y = static_cast<float*>(ctx->planes_dst.p); float* O = static_cast<float*>(ctx->yprime.p);
    cudaStream_t s = ctx->stream;
    CK(cudaMemcpyAsync(Y, y, n * sizeof(float), cudaMemcpyHostToDevice, s));
    if (mode == SRCNN_B200_CONV_EXACT) {
        if (const int ce = launch_conv_exact(Y, w, h, 0, h, O, 0, h, s)) {
            ctx->err = std::string("conv_exact set-up: ") + cudaGetErrorString((cudaError_t)ce);
            return SRCNN_B200_ECUDA;
        }
        ctx->launches++;
    } else if (mode == SRCNN_B200_CONV_TENSOR) {
        const Y16Geom g = y16_geometry(w, h);
        CK(ctx->y16.ensure((size_t)g.pitch * g.rows * sizeof(__half)));
        launch_pack_y16(Y, w, h, 0, static_cast<__half*>(ctx->y16.p), g, 0, s); ctx->launches++;
        if (launch_conv_tc(ctx->tc, static_cast<const __half*>(ctx->y16.p), g, w, h, O, 0, h, ctx->num_sms, s)) {
            ctx->err = std::string("conv_tc: ") + conv_tc_error(ctx->tc);
            return SRCNN_B200_ECUDA;
        }
        ctx->launches++;
    } else return SRCNN_B200_EINVAL;
    CK(cudaMemcpyAsync(out, O, n * sizeof(float), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    return 0;
}

int srcnn_b200_stage_color_out(srcnn_b200_ctx* ctx, const float* const* planes, unsigned w, unsigned h,
                               unsigned d, uint8_t* out, uint8_t* conv)
{
    if (!ctx || !planes || !out || !w || !h || (d != 3 && d != 4)) return SRCNN_B200_EINVAL;
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)w * h, ps = (n + 63) / 64 * 64;
    CK(ctx->planes_dst.ensure(ps * d * sizeof(float))); CK(ctx->out8[0].ensure(n * d)); CK(ctx->conv8[0].ensure(n));
    float* P = static_cast<float*>(ctx->planes_dst.p);
    cudaStream_t s = ctx->stream;
    for (unsigned i = 0; i < d; i++) CK(cudaMemcpyAsync(P + i * ps, planes[i], n * sizeof(float), cudaMemcpyHostToDevice, s));
    launch_color_out(P, P + ps, P + 2 * ps, d == 4 ? P + 3 * ps : nullptr, n, d, static_cast<uint8_t*>(ctx->out8[0].p),
                     conv ? static_cast<uint8_t*>(ctx->conv8[0].p) : nullptr, s);
    ctx->launches++;
    CK(cudaMemcpyAsync(out, ctx->out8[0].p, n * d, cudaMemcpyDeviceToHost, s));
    if (conv) CK(cudaMemcpyAsync(conv, ctx->conv8[0].p, n, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    return 0;
}

int srcnn_b200_contrib_table(int filter, unsigned dst, unsigned src, int* left, int* right, double* weights, int maxwin)
{
    if (!dst || !src || !left || !right || !weights) return SRCNN_B200_EINVAL;
    const int f = (filter >= kNearest && filter <= kBspline) ? filter : kNearest;
    AxisTable t; build_axis_table(f, dst, src, t);
    if (t.win > maxwin) return -1;
    for (unsigned u = 0; u < dst; u++) {
        left[u] = t.left[u]; right[u] = t.right[u];
        for (int i = 0; i < maxwin; i++) weights[(size_t)u * maxwin + i] = i < t.win ? t.w[(size_t)u * t.win + i] : 0.0;
    }
    return t.win;
}

int srcnn_b200_contrib_table_compact(int filter, unsigned dst, unsigned src, int* cleft, double* weights, int maxwin,
                                     int* exact29)
{
    if (!dst || !src || !cleft || !weights) return SRCNN_B200_EINVAL;
    const int f = (filter >= kNearest && filter <= kBspline) ? filter : kNearest;
    AxisTable t; build_axis_table(f, dst, src, t);
    if (!t.compact) return 0;
    if (t.cwin > maxwin) return -1;
    for (unsigned u = 0; u < dst; u++) {
        cleft[u] = t.cleft[u];
        for (int i = 0; i < maxwin; i++) weights[(size_t)u * maxwin + i] = i < t.cwin ? t.cw[(size_t)u * t.cwin + i] : 0.0;
    }
    if (exact29) *exact29 = (t.exact29 ? 1 : 0) | (t.pow2pair ? 2 : 0);
    return t.cwin;
}

} // extern "C"
