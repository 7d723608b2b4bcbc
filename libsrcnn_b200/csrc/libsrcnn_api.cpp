/* The C++ drop-in API (include/libsrcnn.h) on top of the extern "C" layer.
 *
 *   ConfigureFilterSRCNN = reference src/libsrcnn.cpp:930-941  (two process-wide settings)
 *   ProcessSRCNN         = reference src/libsrcnn.cpp:943-1064 (validation, step-scale chain, new[] outputs)
 *
 * Same mangled symbols, same return codes, same ownership (both outputs are new[] arrays the caller
 * delete[]s).  All image arithmetic happens on the GPU; if no B200 context can be created the call returns
 * -20 -- there is deliberately no CPU path.
 *
 * Re-entrancy.  The reference is re-entrant apart from its two settings (src/libsrcnn.cpp:91-92): concurrent calls from
 * several threads each spawn their own OpenMP team.  Here concurrent calls each take their own GPU context from a small
 * pool (a context = a device, three streams and grow-only buffers), so they overlap on the GPU instead of queueing behind
 * one lock; a call only waits when every context of the pool is busy.
 *
 * One call, all GPUs.  With SRCNN_B200_DEVICES listing more than one device ("0,1,2,3" or "all") every call is cut into row
 * strips over those GPUs by the in-library multi-GPU driver (srcnn_multi.cpp), the way the reference spreads one call
 * over all host cores.
 *
 * Environment knobs (the header stays untouched):
 *   SRCNN_B200_DEVICE=<n>      CUDA device of the single-GPU contexts (default 0)
 *   SRCNN_B200_DEVICES=<list>  "all" or a comma-separated device list: one call = row strips over all listed GPUs
 *   SRCNN_B200_CONTEXTS=<n>    size of the context pool = concurrent ProcessSRCNN calls on the GPU (default 2, 1..16)
 *   SRCNN_B200_CONV=exact      use the bit-exact FP32 CUDA-core convolution instead of the tcgen05 kernel
 *   SRCNN_B200_CHAIN=tensor    step-scale chains (factor > 2): run the intermediate x2 passes on the tcgen05 kernel too
 *                              (default: intermediates use the bit-exact kernel so the chained result stays within
 *                              1 LSB of the reference -- measured 2-7 LSB otherwise; only the last pass uses
 *                              the tensor cores)
 */
#include "../../include/libsrcnn.h"
#include "../../include/srcnn_b200.h"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

namespace {

std::atomic<int>  g_filter{ (int)SRCNNF_Bicubic };     /* defaults: src/libsrcnn.cpp:91-92 */
std::atomic<bool> g_stepscale{ false };

struct Pool {
    std::mutex m;
    std::condition_variable cv;
    std::vector<srcnn_b200_ctx*> idle, all;
    int creating = 0, limit = 2, device = 0, conv_mode = SRCNN_B200_CONV_TENSOR;
    bool configured = false, dead = false;
    srcnn_b200_multi* multi = nullptr;                 /* SRCNN_B200_DEVICES: one call over several GPUs */
    bool multi_tried = false;

    void configure()                                   /* under m */
    {
        if (configured) return;
        configured = true;
        if (const char* e = std::getenv("SRCNN_B200_DEVICE")) device = std::atoi(e);
        if (const char* e = std::getenv("SRCNN_B200_CONTEXTS")) limit = std::min(16, std::max(1, std::atoi(e)));
        const char* cm = std::getenv("SRCNN_B200_CONV");
        if (cm && std::strcmp(cm, "exact") == 0) conv_mode = SRCNN_B200_CONV_EXACT;
        std::atexit([] { pool().shutdown(); });
    }

    static Pool& pool() { static Pool* p = new Pool; return *p; }     /* (never destroyed: calls may race with exit) */

    srcnn_b200_multi* multi_handle()                   /* NULL: single-GPU contexts */
    {
        std::lock_guard<std::mutex> g(m);
        configure();
        if (multi_tried) return multi;
        multi_tried = true;
        const char* e = std::getenv("SRCNN_B200_DEVICES");
        if (!e || !*e) return nullptr;
        std::vector<int> devs;
        if (std::strcmp(e, "all") != 0) {
            for (const char* q = e; *q;) {
                char* end = nullptr;
                const long v = std::strtol(q, &end, 10);
                if (end == q) break;
                devs.push_back((int)v);
                q = (*end == ',') ? end + 1 : end;
                if (*end && *end != ',') break;
            }
            if (devs.size() < 2) return nullptr;
        } else if (srcnn_b200_device_count() < 2) return nullptr;
        if (srcnn_b200_multi_create(devs.empty() ? nullptr : devs.data(), (int)devs.size(), &multi) != 0) multi = nullptr;
        if (multi) srcnn_b200_multi_set_conv_mode(multi, conv_mode);
        return multi;
    }

    srcnn_b200_ctx* acquire()
    {
        std::unique_lock<std::mutex> l(m);
        configure();
        for (;;) {
            if (dead) return nullptr;
            if (!idle.empty()) { srcnn_b200_ctx* c = idle.back(); idle.pop_back(); return c; }
            if ((int)all.size() + creating < limit) {
                creating++;
                l.unlock();
                srcnn_b200_ctx* c = nullptr;
                if (srcnn_b200_create(device, &c) != 0) c = nullptr;
                if (c) srcnn_b200_set_conv_mode(c, conv_mode);
                l.lock();
                creating--;
                if (c) { all.push_back(c); return c; }
                if (all.empty() && creating == 0) { dead = true; cv.notify_all(); return nullptr; }   /* no usable GPU at all */
                limit = std::max(1, (int)all.size() + creating);     /* (out of memory?) live with the contexts that exist */
                continue;
            }
            cv.wait(l);
        }
    }
    void release(srcnn_b200_ctx* c)
    {
        { std::lock_guard<std::mutex> g(m); idle.push_back(c); }
        cv.notify_one();
    }
    void shutdown()
    {
        std::lock_guard<std::mutex> g(m);
        for (srcnn_b200_ctx* c : idle) srcnn_b200_destroy(c);       /* (contexts still in use by a racing call are left alone) */
        idle.clear(); all.clear();
        if (multi) { srcnn_b200_multi_destroy(multi); multi = nullptr; }
        dead = true;
    }
};

} // namespace

void SRCNN_API ConfigureFilterSRCNN( SRCNNFilterType ftype, bool stepscale )
{
    g_filter.store((int)ftype);
    g_stepscale.store(stepscale);
}

int SRCNN_API ProcessSRCNN( const unsigned char* refbuff,
                            unsigned w, unsigned h, unsigned d,
                            float multiply,
                            unsigned char* &outbuff,
                            unsigned &outbuffsz,
                            unsigned char** convbuff,
                            unsigned* convbuffsz )
{
    if (refbuff == nullptr || w == 0 || h == 0 || d == 0) return -1;
    if (d != 3 && d != 4) return -1;          /* the reference reads garbage / overruns here; we refuse */
    if ((float)w * multiply <= 0.f || (float)h * multiply <= 0.f) return -2;

    /* snapshot the process-wide settings once per call */
    const int  filter = g_filter.load();
    const bool step   = g_stepscale.load();

    unsigned ow = 0, oh = 0;
    const int passes = srcnn_b200_process_dims(w, h, multiply, step ? 1 : 0, &ow, &oh);
    if (passes == 0) {                        /* step mode, effective factor 1: src/libsrcnn.cpp:1007-1008 */
        outbuff = nullptr; outbuffsz = 0;
        return -100;
    }
    if (ow == 0 || oh == 0) return -2;

    /* step mode whose last pass degenerates to factor 1 (e.g. x8): the reference never hands convbuff to a pass */
    const bool want_conv = (convbuff != nullptr) && (convbuffsz != nullptr) &&
                           srcnn_b200_process_has_conv(w, multiply, step ? 1 : 0) != 0;
    const size_t osz = (size_t)ow * oh * d, csz = (size_t)ow * oh;
    unsigned char* obuf = new (std::nothrow) unsigned char[osz];
    if (!obuf) return -11;
    unsigned char* cbuf = nullptr;
    if (want_conv) {
        cbuf = new (std::nothrow) unsigned char[csz];
        if (!cbuf) { delete[] obuf; return -12; }
    }

    int rc;
    Pool& P = Pool::pool();
    if (srcnn_b200_multi* mg = P.multi_handle()) {
        rc = srcnn_b200_multi_process_host(mg, refbuff, w, h, d, multiply, filter, step ? 1 : 0, obuf, cbuf);
    } else {
        srcnn_b200_ctx* ctx = P.acquire();
        rc = ctx ? srcnn_b200_process_host(ctx, refbuff, w, h, d, multiply, filter, step ? 1 : 0, obuf, cbuf)
                 : SRCNN_B200_ENODEV;
        if (ctx) P.release(ctx);
    }
    if (rc != 0) {
        delete[] obuf; delete[] cbuf;
        return rc;
    }
    outbuff = obuf; outbuffsz = (unsigned)osz;
    if (want_conv) { *convbuff = cbuf; *convbuffsz = (unsigned)csz; }
    return 0;
}
