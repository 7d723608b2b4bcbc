/* Single-process multi-GPU drivers (include/srcnn_b200.h: srcnn_b200_multi_*).
 *
 * The reference spreads ONE ProcessSRCNN call over every core of the host with OpenMP (src/libsrcnn.cpp:791, 817, 495);
 * this is the same idea one level up: ONE host call spreads over every B200 of the box.  The path shards without any
 * exchange step:
 *   - a pass over one image = row strips, one per GPU, each recomputing its 6-row halo of interpolated Y
 *     (srcnn_b200_pass_host_rows on a per-device context; results land in disjoint row ranges of the caller's buffer and are
 *     bit-identical to the single-GPU result);
 *   - a batch of frames = contiguous blocks of frames, one per GPU (srcnn_b200_batch_host).
 * One persistent worker thread per device owns that device's context (a context is not thread-safe and binds its thread
 * to its device); the calling thread only distributes work and waits.  No collective, no peer copies: every GPU reads the
 * source rows it needs from the host buffer and writes its rows back.
 */
#include "../../include/srcnn_b200.h"

#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Worker {
    int device = 0;
    srcnn_b200_ctx* ctx = nullptr;
    std::thread th;
    std::mutex m;
    std::condition_variable cv, done_cv;
    std::function<int(srcnn_b200_ctx*)> job;
    bool has_job = false, quit = false, done = true;
    int rc = 0;

    void loop()
    {
        for (;;) {
            std::function<int(srcnn_b200_ctx*)> j;
            {
                std::unique_lock<std::mutex> l(m);
                cv.wait(l, [&] { return quit || has_job; });
                if (quit) return;
                j = std::move(job); has_job = false;
            }
            const int r = j(ctx);
            { std::lock_guard<std::mutex> g(m); rc = r; done = true; }
            done_cv.notify_all();
        }
    }
    void submit(std::function<int(srcnn_b200_ctx*)> j)
    {
        { std::lock_guard<std::mutex> g(m); job = std::move(j); has_job = true; done = false; }
        cv.notify_all();
    }
    int wait()
    {
        std::unique_lock<std::mutex> l(m);
        done_cv.wait(l, [&] { return done; });
        return rc;
    }
};

struct HostBuf {
    void* p = nullptr; size_t cap = 0;
    bool ensure(size_t n)
    {
        if (n <= cap) return true;
        if (p) srcnn_b200_host_free(p);
        p = srcnn_b200_host_alloc(n); cap = p ? n : 0;
        return p != nullptr;
    }
    ~HostBuf() { if (p) srcnn_b200_host_free(p); }
};

} // namespace

struct srcnn_b200_multi {
    std::vector<Worker*> w;
    std::mutex call_lock;                 /* one multi-GPU call at a time */
    std::string err;
    HostBuf chain[2];                     /* pinned intermediates of step-scale chains */
    int conv_mode = SRCNN_B200_CONV_TENSOR;
    bool chain_tensor = false;
    unsigned min_rows = 256;              /* a GPU is only given a strip of at least this many output rows (SRCNN_B200_MULTI_MIN_ROWS) */
};

namespace {

/* run job(i, ctx) on the first `n` workers, return the first non-zero result */
int run_on(srcnn_b200_multi* m, unsigned n, const std::function<int(unsigned, srcnn_b200_ctx*)>& job)
{
    for (unsigned i = 0; i < n; i++) m->w[i]->submit([i, &job](srcnn_b200_ctx* c) { return job(i, c); });
    int rc = 0; unsigned bad = 0;
    for (unsigned i = 0; i < n; i++) { const int r = m->w[i]->wait(); if (r && !rc) { rc = r; bad = i; } }
    if (rc) m->err = "device " + std::to_string(m->w[bad]->device) + ": " + srcnn_b200_last_error(m->w[bad]->ctx);
    return rc;
}

int multi_pass(srcnn_b200_multi* m, const uint8_t* src, unsigned w, unsigned h, unsigned d, float multiply, int filter,
               int mode, uint8_t* out, uint8_t* conv)
{
    unsigned ow = 0, oh = 0;
    if (!w || !h || (d != 3 && d != 4)) return SRCNN_B200_EINVAL;
    if ((float)w * multiply <= 0.f || (float)h * multiply <= 0.f) return SRCNN_B200_ESIZE;
    srcnn_b200_output_dims(w, h, multiply, &ow, &oh);
    if (!ow || !oh) return SRCNN_B200_ESIZE;
    /* strips: balanced, contiguous, disjoint, covering -- rows [oh*g/G, oh*(g+1)/G); small images use fewer GPUs */
    const unsigned G = (unsigned)std::max<size_t>(1, std::min<size_t>(m->w.size(), oh / std::max(1u, m->min_rows)));
    return run_on(m, G, [=](unsigned g, srcnn_b200_ctx* c) -> int {
        const unsigned r0 = (unsigned)((size_t)oh * g / G), r1 = (unsigned)((size_t)oh * (g + 1) / G);
        if (r0 >= r1) return 0;
        srcnn_b200_set_conv_mode(c, mode);
        return srcnn_b200_pass_host_rows(c, src, w, h, d, multiply, filter, r0, r1, out + (size_t)r0 * ow * d,
                                         conv ? conv + (size_t)r0 * ow : nullptr);
    });
}

} // namespace

extern "C" {

int srcnn_b200_multi_create(const int* devices, int ndev, srcnn_b200_multi** out)
{
    if (!out) return SRCNN_B200_EINVAL;
    *out = nullptr;
    const int avail = srcnn_b200_device_count();
    if (avail <= 0) return SRCNN_B200_ENODEV;
    std::vector<int> devs;
    if (!devices || ndev <= 0) for (int i = 0; i < avail; i++) devs.push_back(i);
    else for (int i = 0; i < ndev; i++) { if (devices[i] < 0 || devices[i] >= avail) return SRCNN_B200_ENODEV; devs.push_back(devices[i]); }
    srcnn_b200_multi* m = new srcnn_b200_multi;
    if (const char* e = std::getenv("SRCNN_B200_CHAIN")) m->chain_tensor = std::strcmp(e, "tensor") == 0;
    if (const char* e = std::getenv("SRCNN_B200_MULTI_MIN_ROWS")) m->min_rows = (unsigned)std::max(1, std::atoi(e));
    /* the staging-copy threads of all contexts share the host's cores */
    const unsigned hw = std::max(2u, std::thread::hardware_concurrency());
    const unsigned per_ctx = std::max(1u, std::min(6u, hw / 2 / (unsigned)devs.size()));
    int rc = 0;
    for (int dev : devs) {
        Worker* wk = new Worker;
        wk->device = dev;
        m->w.push_back(wk);
        wk->th = std::thread([wk] { wk->loop(); });
        /* the context is created ON its worker thread: creation binds the calling thread to the device */
        wk->submit([dev, per_ctx, wk](srcnn_b200_ctx*) -> int {
            const int r = srcnn_b200_create(dev, &wk->ctx);
            if (r == 0) srcnn_b200_set_copy_threads(wk->ctx, per_ctx);
            return r;
        });
    }
    for (Worker* wk : m->w) { const int r = wk->wait(); if (r && !rc) rc = r; }
    if (rc) { srcnn_b200_multi_destroy(m); return rc; }
    *out = m;
    return 0;
}

void srcnn_b200_multi_destroy(srcnn_b200_multi* m)
{
    if (!m) return;
    for (Worker* wk : m->w) {
        if (wk->ctx) { wk->submit([wk](srcnn_b200_ctx* c) { srcnn_b200_destroy(c); wk->ctx = nullptr; return 0; }); wk->wait(); }
        { std::lock_guard<std::mutex> g(wk->m); wk->quit = true; }
        wk->cv.notify_all();
        if (wk->th.joinable()) wk->th.join();
        delete wk;
    }
    delete m;
}

int srcnn_b200_multi_device_count(const srcnn_b200_multi* m) { return m ? (int)m->w.size() : 0; }
const char* srcnn_b200_multi_last_error(const srcnn_b200_multi* m) { return m ? m->err.c_str() : "no handle"; }

int srcnn_b200_multi_set_conv_mode(srcnn_b200_multi* m, int mode)
{
    if (!m || (mode != SRCNN_B200_CONV_TENSOR && mode != SRCNN_B200_CONV_EXACT)) return SRCNN_B200_EINVAL;
    m->conv_mode = mode;
    return 0;
}

int srcnn_b200_multi_pass_host(srcnn_b200_multi* m, const uint8_t* src, unsigned w, unsigned h, unsigned d,
                               float multiply, int filter, uint8_t* out, uint8_t* conv)
{
    if (!m || !src || !out) return SRCNN_B200_EINVAL;
    std::lock_guard<std::mutex> g(m->call_lock);
    return multi_pass(m, src, w, h, d, multiply, filter, m->conv_mode, out, conv);
}

int srcnn_b200_multi_process_host(srcnn_b200_multi* m, const uint8_t* src, unsigned w, unsigned h, unsigned d,
                                  float multiply, int filter, int stepscale, uint8_t* out, uint8_t* conv)
{
    if (!m || !src || !out || !w || !h) return SRCNN_B200_EINVAL;
    if (d != 3 && d != 4) return SRCNN_B200_EINVAL;
    if ((float)w * multiply <= 0.f || (float)h * multiply <= 0.f) return SRCNN_B200_ESIZE;
    std::lock_guard<std::mutex> g(m->call_lock);
    if (!stepscale) return multi_pass(m, src, w, h, d, multiply, filter, m->conv_mode, out, conv);
    /* step-scale chain (src/libsrcnn.cpp:980-1061): like the reference, every pass reads the previous pass's 8-bit image
     * from a host buffer (pinned here) -- each pass then uses all GPUs */
    float f[64];
    const int n = srcnn_b200_step_factors(w, multiply, f, 64);
    if (n == 0) return SRCNN_B200_ENOTHING;
    if (!srcnn_b200_process_has_conv(w, multiply, 1)) conv = nullptr;
    const uint8_t* cur = src;
    unsigned cw = w, ch = h;
    for (int i = 0; i < n; i++) {
        const bool last = i + 1 == n;
        unsigned nw = 0, nh = 0;
        srcnn_b200_output_dims(cw, ch, f[i], &nw, &nh);
        if (!nw || !nh) return SRCNN_B200_ESIZE;
        uint8_t* dst = out;
        if (!last) {
            if (!m->chain[i & 1].ensure((size_t)nw * nh * d)) { m->err = "pinned chain buffer"; return SRCNN_B200_ECUDA; }
            dst = static_cast<uint8_t*>(m->chain[i & 1].p);
        }
        /* intermediates on the bit-exact kernel unless SRCNN_B200_CHAIN=tensor (srcnn_ctx.cu: process_host_impl) */
        const int mode = (!last && !m->chain_tensor) ? SRCNN_B200_CONV_EXACT : m->conv_mode;
        const int rc = multi_pass(m, cur, cw, ch, d, f[i], filter, mode, dst, last ? conv : nullptr);
        if (rc) return rc;
        cur = dst; cw = nw; ch = nh;
    }
    return 0;
}

int srcnn_b200_multi_batch_host(srcnn_b200_multi* m, const uint8_t* frames, unsigned n, unsigned w, unsigned h,
                                unsigned d, float multiply, int filter, uint8_t* out, uint8_t* conv)
{
    if (!m || !frames || !out) return SRCNN_B200_EINVAL;
    if (n == 0) return 0;
    unsigned ow = 0, oh = 0;
    if (!w || !h || (d != 3 && d != 4)) return SRCNN_B200_EINVAL;
    if ((float)w * multiply <= 0.f || (float)h * multiply <= 0.f) return SRCNN_B200_ESIZE;
    srcnn_b200_output_dims(w, h, multiply, &ow, &oh);
    if (!ow || !oh) return SRCNN_B200_ESIZE;
    std::lock_guard<std::mutex> g(m->call_lock);
    const unsigned G = (unsigned)std::min<size_t>(m->w.size(), n);
    const size_t ib = (size_t)w * h * d, ob = (size_t)ow * oh * d, cb = (size_t)ow * oh;
    const int mode = m->conv_mode;
    return run_on(m, G, [=](unsigned gi, srcnn_b200_ctx* c) -> int {
        const unsigned f0 = (unsigned)((size_t)n * gi / G), f1 = (unsigned)((size_t)n * (gi + 1) / G);
        if (f0 >= f1) return 0;
        srcnn_b200_set_conv_mode(c, mode);
        return srcnn_b200_batch_host(c, frames + f0 * ib, f1 - f0, w, h, d, multiply, filter, out + f0 * ob,
                                     conv ? conv + f0 * cb : nullptr);
    });
}

} // extern "C"
