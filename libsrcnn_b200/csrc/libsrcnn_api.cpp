/* The C++ drop-in API (include/libsrcnn.h) on top of the extern "C" layer.
 *
 *   ConfigureFilterSRCNN = reference src/libsrcnn.cpp:930-941  (two process-wide settings)
 *   ProcessSRCNN         = reference src/libsrcnn.cpp:943-1064 (validation, step-scale chain, new[] outputs)
 *
 * Same mangled symbols, same return codes, same ownership (both outputs are new[] arrays the caller
 * delete[]s).  All image arithmetic happens on the GPU; if no B200 context can be created the call returns
 * -20 -- there is deliberately no CPU path.
 *
 * Environment knobs (the header stays untouched):
 *   SRCNN_B200_DEVICE=<n>     CUDA device of the process-wide context (default 0)
 *   SRCNN_B200_CONV=exact     use the bit-exact FP32 CUDA-core convolution instead of the tcgen05 kernel
 *   SRCNN_B200_CHAIN=tensor   step-scale chains (factor > 2): run the intermediate x2 passes on the tcgen05 kernel too
 *                             (default: intermediates use the bit-exact kernel so the chained result stays within
 *                             1 LSB of the reference; only the last pass uses the tensor cores)
 */
#include "../../include/libsrcnn.h"
#include "../../include/srcnn_b200.h"

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>

namespace {

std::atomic<int>  g_filter{ (int)SRCNNF_Bicubic };     /* defaults: src/libsrcnn.cpp:91-92 */
std::atomic<bool> g_stepscale{ false };

std::mutex       g_lock;                               /* one context, calls serialised */
srcnn_b200_ctx*  g_ctx = nullptr;
bool             g_tried = false;

srcnn_b200_ctx* context()
{
    if (!g_tried) {
        g_tried = true;
        int dev = 0;
        if (const char* e = std::getenv("SRCNN_B200_DEVICE")) dev = std::atoi(e);
        if (srcnn_b200_create(dev, &g_ctx) != 0) g_ctx = nullptr;
        if (g_ctx) {
            const char* m = std::getenv("SRCNN_B200_CONV");
            if (m && std::strcmp(m, "exact") == 0) srcnn_b200_set_conv_mode(g_ctx, SRCNN_B200_CONV_EXACT);
            std::atexit([] { std::lock_guard<std::mutex> g(g_lock); srcnn_b200_destroy(g_ctx); g_ctx = nullptr; });
        }
    }
    return g_ctx;
}

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

    const bool want_conv = (convbuff != nullptr) && (convbuffsz != nullptr);
    const size_t osz = (size_t)ow * oh * d, csz = (size_t)ow * oh;
    unsigned char* obuf = new (std::nothrow) unsigned char[osz];
    if (!obuf) return -11;
    unsigned char* cbuf = nullptr;
    if (want_conv) {
        cbuf = new (std::nothrow) unsigned char[csz];
        if (!cbuf) { delete[] obuf; return -12; }
    }

    int rc;
    {
        std::lock_guard<std::mutex> guard(g_lock);
        srcnn_b200_ctx* ctx = context();
        rc = ctx ? srcnn_b200_process_host(ctx, refbuff, w, h, d, multiply, filter, step ? 1 : 0, obuf, cbuf)
                 : SRCNN_B200_ENODEV;
    }
    if (rc != 0) {
        delete[] obuf; delete[] cbuf;
        return rc;
    }
    outbuff = obuf; outbuffsz = (unsigned)osz;
    if (want_conv) { *convbuff = cbuf; *convbuffsz = (unsigned)csz; }
    return 0;
}
