/* TEST INFRASTRUCTURE ONLY -- never linked into, loaded by, or shipped with the product.
 *
 * Compiles the UNMODIFIED reference (rageworx/libsrcnn) in place and exposes its
 * internal stages through a flat extern "C" surface so the tests can pin every
 * stage of the B200 path (and the plain-C restatement in srcnn_oracle.c) against
 * the reference's own arithmetic.  The reference sources are NOT copied: this TU
 * includes them from $(REFSRC) (= /root/reference/src) at build time, with the
 * flags of the reference's Makefiles/Makefile.linux:26-35 (-O2 -fopenmp, no
 * -ffast-math, NEW_FAST_I_II_LAYERS undefined = the shipped three-stage path).
 * Output goes to oracle/_ref/ only (git-ignored, travels to the GPU box).
 */
#include "libsrcnn.cpp"   /* resolved via -I$(REFSRC); brings in frawscale.h, convdata.h */

#include <cstdint>

namespace {
FRAWGenericFilter* make_filter(int kind)
{
    switch (kind) {                      /* same numbering as SRCNNFilterType, src/libsrcnn.h:37-44 */
        case 0:  return new FRAWBoxFilter;
        case 1:  return new FRAWBilinearFilter;
        case 2:  return new FRAWBicubicFilter;
        case 3:  return new FRAWLanczos3Filter;
        case 4:  return new FRAWBSplineFilter;
        default: return new FRAWBoxFilter;
    }
}
libsrcnn::ImgF32 view(const float* p, unsigned w, unsigned h)
{
    libsrcnn::ImgF32 v; v.width = w; v.height = h; v.depth = 1; v.buff = const_cast<float*>(p); return v;
}
}

extern "C" {

int ref_omp_threads() { return omp_get_max_threads(); }

/* public API: ConfigureFilterSRCNN + ProcessSRCNN (src/libsrcnn.cpp:930-1064) */
int ref_process(const unsigned char* src, unsigned w, unsigned h, unsigned d, float multiply,
                int filter, int stepscale,
                unsigned char** out, unsigned* outsz, unsigned char** conv, unsigned* convsz)
{
    ConfigureFilterSRCNN((SRCNNFilterType)filter, stepscale != 0);
    unsigned char* o = NULL; unsigned osz = 0;
    int rc = ProcessSRCNN(src, w, h, d, multiply, o, osz, conv, convsz);
    *out = o; *outsz = osz;
    return rc;
}
void ref_free(unsigned char* p) { delete[] p; }

/* src/libsrcnn.cpp:233-272; planes are caller-allocated w*h floats (a may be NULL for d==3) */
void ref_color_in(const unsigned char* src, unsigned w, unsigned h, unsigned d,
                  float* y, float* cb, float* cr, float* a)
{
    libsrcnn::ImgU8 in = { w, h, d, const_cast<unsigned char*>(src) };
    libsrcnn::ImgYCbCr o;
    libsrcnn::converImgU8toYCbCr(in, o);
    size_t n = (size_t)w * h * sizeof(float);
    memcpy(y, o.Y.buff, n); memcpy(cb, o.Cb.buff, n); memcpy(cr, o.Cr.buff, n);
    if (d == 4 && a) memcpy(a, o.A.buff, n);
    libsrcnn::discardImgYCbCr(o);
}

/* FRAWResizeEngine::scale, src/frawscale.cpp:162-286 */
int ref_resize(const float* src, unsigned sw, unsigned sh, unsigned dw, unsigned dh, int filter, float* dst)
{
    FRAWGenericFilter* f = make_filter(filter);
    FRAWResizeEngine e(f);
    float* o = NULL;
    unsigned rc = e.scale(src, sw, sh, dw, dh, &o);
    if (o) { memcpy(dst, o, (size_t)dw * dh * sizeof(float)); delete[] o; }
    delete f;
    return (int)rc;
}

/* FRawScaleWeightsTable, src/frawscale.cpp:8-112: dump left/right/weights; returns window size */
int ref_contrib_table(int filter, unsigned dst, unsigned src, int* left, int* right, double* weights, int maxwin)
{
    FRAWGenericFilter* f = make_filter(filter);
    FRawScaleWeightsTable t(f, dst, src);
    for (unsigned u = 0; u < dst; u++) {
        left[u] = (int)t.getLeftBoundary(u); right[u] = (int)t.getRightBoundary(u);
        for (int i = 0; i < maxwin; i++)
            weights[(size_t)u * maxwin + i] = (i <= right[u] - left[u]) ? t.getWeight(u, i) : 0.0;
    }
    delete f;
    return 0;
}

/* conv I: 64 x convolution99, src/libsrcnn.cpp:350-422 + driver :785-798; out = 64 planes of w*h */
void ref_conv1(const float* y, unsigned w, unsigned h, float* out)
{
    libsrcnn::ImgF32 s = view(y, w, h);
    #pragma omp parallel for
    for (unsigned k = 0; k < CONV1_FILTERS; k++) {
        libsrcnn::ImgF32 d = view(out + (size_t)k * w * h, w, h);
        libsrcnn::convolution99(s, d, weights_conv1_data[k], biases_conv1[k]);
    }
}
/* conv II: 32 x convolution11, src/libsrcnn.cpp:424-447 + driver :811-824 */
void ref_conv2(const float* l1, unsigned w, unsigned h, float* out)
{
    libsrcnn::ImgConv1Layers s;
    for (unsigned k = 0; k < CONV1_FILTERS; k++) s[k] = view(l1 + (size_t)k * w * h, w, h);
    #pragma omp parallel for
    for (unsigned n = 0; n < CONV2_FILTERS; n++) {
        libsrcnn::ImgF32 d = view(out + (size_t)n * w * h, w, h);
        libsrcnn::convolution11(s, d, weights_conv2_data[n], biases_conv2[n]);
    }
}
/* conv III: convolution55, src/libsrcnn.cpp:449-529 */
void ref_conv3(const float* l2, unsigned w, unsigned h, float* out)
{
    libsrcnn::ImgConv2Layers s;
    for (unsigned n = 0; n < CONV2_FILTERS; n++) s[n] = view(l2 + (size_t)n * w * h, w, h);
    libsrcnn::ImgF32 d = view(out, w, h);
    libsrcnn::convolution55(s, d, weights_conv3_data, biases_conv3);
}
/* colour-out: convertImgF32XtoImgU8, src/libsrcnn.cpp:274-308; planes = d pointers */
void ref_color_out(const float* const* planes, unsigned w, unsigned h, unsigned d, unsigned char* out)
{
    libsrcnn::ImgF32 p[4];
    for (unsigned i = 0; i < d; i++) p[i] = view(planes[i], w, h);
    libsrcnn::ImgU8 o; o.buff = NULL;
    libsrcnn::convertImgF32XtoImgU8(p, d, o);
    memcpy(out, o.buff, (size_t)w * h * d);
    delete[] o.buff;
}
/* raw parameter tables, reference order: b1[64] w1[64][9][9] b2[32] w2[32][64] b3 w3[32][5][5] */
void ref_weights(float* out)
{
    float* p = out;
    memcpy(p, biases_conv1, sizeof(biases_conv1)); p += 64;
    memcpy(p, weights_conv1_data, sizeof(weights_conv1_data)); p += 64 * 81;
    memcpy(p, biases_conv2, sizeof(biases_conv2)); p += 32;
    memcpy(p, weights_conv2_data, sizeof(weights_conv2_data)); p += 32 * 64;
    *p++ = biases_conv3;
    memcpy(p, weights_conv3_data, sizeof(weights_conv3_data));
}

} /* extern "C" */
