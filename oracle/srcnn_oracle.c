/* TEST INFRASTRUCTURE ONLY (see srcnn_oracle.h).
 *
 * Plain-C restatement of rageworx/libsrcnn's ProcessSRCNN path, written from the behaviour of the
 * reference (file:line cited per function, paths relative to the reference root).  Compiled with
 * -ffp-contract=off so that every float/double operation is individually rounded exactly like the
 * reference's x86-64 -O2 build (no FMA contraction).  The three convolutions keep the reference's
 * summation order, so every stage here is bit-identical to the compiled reference
 * (tests/test_oracle.py asserts max|diff| == 0 against oracle/_ref and the golden PNGs).
 *
 * Unlike the reference this port never materialises the 64+32 full-resolution planes when running a
 * whole pass: orc_pass evaluates conv I/II per pixel position (same per-value arithmetic, same order)
 * and keeps only the 32-plane conv II result that conv III needs.
 */
#include "srcnn_oracle.h"
#include "srcnn_weights_data.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int orc_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* src/libsrcnn.cpp:662-663 -- unsigned = unsigned * float evaluates in float and truncates */
void orc_output_dims(unsigned w, unsigned h, float multiply, unsigned* ow, unsigned* oh)
{
    *ow = (unsigned)((float)w * multiply);
    *oh = (unsigned)((float)h * multiply);
}

/* ------------------------------------------------------------------ colour in
 * src/libsrcnn.cpp:233-272 (converImgU8toYCbCr): left-to-right float arithmetic. */
void orc_color_in(const unsigned char* src, unsigned w, unsigned h, unsigned d,
                  float* y, float* cb, float* cr, float* a)
{
    size_t n = (size_t)w * h;
    #pragma omp parallel for
    for (size_t i = 0; i < n; i++) {
        float r = (float)src[i * d + 0], g = (float)src[i * d + 1], b = (float)src[i * d + 2];
        y[i]  = ((0.299f * r) + (0.587f * g)) + (0.114f * b);
        cb[i] = ((128.f - (0.1687f * r)) - (0.3313f * g)) + (0.5f * b);
        cr[i] = ((128.f + (0.5f * r)) - (0.4187f * g)) - (0.0813f * b);
        if (d == 4 && a) a[i] = (float)src[i * d + 3];
    }
}

/* ------------------------------------------------------------------ resampling filters
 * src/frawscale.h:58-172.  support: box .5, bilinear 1, bicubic(Mitchell B=C=1/3) 2, lanczos3 3, bspline 2 */
static double filt_support(int f)
{
    switch (f) { case ORC_BILINEAR: return 1.0; case ORC_BICUBIC: return 2.0;
                 case ORC_LANCZOS3: return 3.0; case ORC_BSPLINE: return 2.0; default: return 0.5; }
}
static double sinc_pi(double v)            /* src/frawscale.h:142-150 */
{
    if (v != 0) { v *= 3.1415926535897932384626433832795; return sin(v) / v; }
    return 1;
}
static double filt_eval(int f, double v)
{
    v = fabs(v);
    switch (f) {
    case ORC_BILINEAR:                     /* src/frawscale.h:78-83 */
        return v < 1.0 ? 1.0 - v : 0.0;
    case ORC_BICUBIC: {                    /* src/frawscale.h:95-120, b = c = 1/3 */
        const double b = 1 / (double)3, c = 1 / (double)3;
        const double p0 = (6 - 2 * b) / 6, p2 = (-18 + 12 * b + 6 * c) / 6, p3 = (12 - 9 * b - 6 * c) / 6;
        const double q0 = (8 * b + 24 * c) / 6, q1 = (-12 * b - 48 * c) / 6,
                     q2 = (6 * b + 30 * c) / 6, q3 = (-b - 6 * c) / 6;
        if (v < 1) return p0 + v * v * (p2 + v * p3);
        if (v < 2) return q0 + v * (q1 + v * (q2 + v * q3));
        return 0; }
    case ORC_LANCZOS3:                     /* src/frawscale.h:131-140 */
        return v < 3.0 ? sinc_pi(v) * sinc_pi(v / 3.0) : 0.0;
    case ORC_BSPLINE:                      /* src/frawscale.h:161-171 */
        if (v < 1) return (4 + v * v * (-6 + 3 * v)) / 6;
        if (v < 2) { double t = 2 - v; return t * t * t / 6; }
        return 0;
    default:                               /* box, src/frawscale.h:68-69 */
        return v <= 0.5 ? 1.0 : 0.0;
    }
}

/* window size of the contribution table, src/frawscale.cpp:23-35 */
int orc_window(int filter, unsigned dst, unsigned src)
{
    double scale = (double)dst / (double)src, width = filt_support(filter);
    if (scale < 1.0) width = width / scale;
    return 2 * (int)ceil(width) + 1;
}

/* FRawScaleWeightsTable ctor, src/frawscale.cpp:8-112.  weights is dst x maxwin, zero padded. */
int orc_contrib_table(int filter, unsigned dst, unsigned src, int* left, int* right, double* weights, int maxwin)
{
    const double scale = (double)dst / (double)src;
    double width = filt_support(filter), fscale = 1.0;
    if (scale < 1.0) { width = width / scale; fscale = scale; }
    const int win = 2 * (int)ceil(width) + 1;
    if (win > maxwin) return -1;
    const double off = (0.5 / scale) - 0.5;
    for (unsigned u = 0; u < dst; u++) {
        double* w = weights + (size_t)u * maxwin;
        for (int i = 0; i < maxwin; i++) w[i] = 0.0;
        const double c = (double)u / scale + off;
        int L = (int)floor(c - width); if (L < 0) L = 0;
        int R = (int)ceil(c + width);  if (R > (int)src - 1) R = (int)src - 1;
        if (R - L + 1 > win) {
            /* reference tests iLeft < int(src) - 1/2 == src, always true (src/frawscale.cpp:57) */
            if (L < (int)src) L++; else R--;
        }
        double tot = 0;
        for (int i = L; i <= R; i++) {
            double wt = fscale * filt_eval(filter, fscale * (c - (double)i));
            w[i - L] = wt; tot += wt;
        }
        if (tot > 0 && tot != 1) {
            for (int i = L; i <= R; i++) w[i - L] /= tot;
            int k = R - L;                 /* drop trailing zero weights, src/frawscale.cpp:95-106 */
            while (w[k] == 0) { R--; k--; if (R == L) break; }
        }
        left[u] = L; right[u] = R;
    }
    return win;
}

/* one separable pass along an axis; stride_in/out select horizontal or vertical use.
 * src/frawscale.cpp:288-332 (horizontal), :335-385 (vertical): double accumulate, ascending taps, float store */
static void resample_axis(const float* src, float* dst, unsigned lines, unsigned n_src, unsigned n_dst,
                          size_t line_stride_src, size_t elem_stride_src,
                          size_t line_stride_dst, size_t elem_stride_dst, int filter)
{
    int win = orc_window(filter, n_dst, n_src);
    int* L = (int*)malloc(sizeof(int) * n_dst); int* R = (int*)malloc(sizeof(int) * n_dst);
    double* W = (double*)malloc(sizeof(double) * (size_t)n_dst * win);
    orc_contrib_table(filter, n_dst, n_src, L, R, W, win);
    #pragma omp parallel for
    for (unsigned l = 0; l < lines; l++) {
        const float* s = src + l * line_stride_src; float* d = dst + l * line_stride_dst;
        for (unsigned u = 0; u < n_dst; u++) {
            double g = 0.0; const double* w = W + (size_t)u * win;
            for (int i = 0; i <= R[u] - L[u]; i++) {
                double px = (double)s[(size_t)(L[u] + i) * elem_stride_src];
                g += w[i] * px;
            }
            d[u * elem_stride_dst] = (float)g;
        }
    }
    free(L); free(R); free(W);
}

/* FRAWResizeEngine::scale, src/frawscale.cpp:162-286.  Upscale (dw > sw): vertical into sw x dh, then
 * horizontal; otherwise horizontal first.  The same-size shortcut copies only sw*sh*2 bytes in the
 * reference (sizeof(unsigned short) bug, :189); here an equal-size call copies the whole plane and
 * returns -1 so callers can tell the case apart (outside every BASELINE config). */
int orc_resize(const float* src, unsigned sw, unsigned sh, unsigned dw, unsigned dh, int filter, float* dst)
{
    if (!src || !sw || !sh || !dw || !dh) return 0;
    if (sw == dw && sh == dh) { memcpy(dst, src, (size_t)sw * sh * sizeof(float)); return -1; }
    if (dw <= sw) {
        const float* t = src; float* tmp = NULL;
        if (sw != dw) {
            float* o = (sh != dh) ? (tmp = (float*)malloc(sizeof(float) * (size_t)dw * sh)) : dst;
            resample_axis(src, o, sh, sw, dw, sw, 1, dw, 1, filter);
            t = o;
        }
        if (sh != dh) resample_axis(t, dst, dw, sh, dh, 1, dw, 1, dw, filter);
        free(tmp);
    } else {
        const float* t = src; float* tmp = NULL;
        if (sh != dh) {
            float* o = (sw != dw) ? (tmp = (float*)malloc(sizeof(float) * (size_t)sw * dh)) : dst;
            resample_axis(src, o, sw, sh, dh, 1, sw, 1, sw, filter);
            t = o;
        }
        if (sw != dw) resample_axis(t, dst, dh, sw, dw, sw, 1, dw, 1, filter);
        free(tmp);
    }
    return (int)(dw * dh);
}

/* ------------------------------------------------------------------ convolutions */
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* conv I at one position for all 64 filters.  src/libsrcnn.cpp:350-422: clamp-padded by 4, taps row-major
 * (x = row, y = col in the reference's loop), float running sum from 0, bias last, ReLU. */
static void conv1_at(const float* y, int w, int h, int row, int col, float* v64)
{
    float win[81];
    for (int i = 0; i < 9; i++) {
        int rr = clampi(row + i - 4, 0, h - 1);
        for (int j = 0; j < 9; j++) win[i * 9 + j] = y[(size_t)rr * w + clampi(col + j - 4, 0, w - 1)];
    }
    for (int k = 0; k < 64; k++) {
        float t = 0;
        for (int tap = 0; tap < 81; tap++) t += SRCNN_W1[tap][k] * win[tap];
        t += SRCNN_B1[k];
        v64[k] = (t >= 0) ? t : 0;
    }
}
/* conv II at one position.  src/libsrcnn.cpp:424-447: sum over fc = 0..63 in order, bias last, ReLU. */
static void conv2_at(const float* v64, float* v32)
{
    for (int n = 0; n < 32; n++) {
        float t = 0;
        for (int k = 0; k < 64; k++) t += v64[k] * SRCNN_W2[k][n];
        t += SRCNN_B2[n];
        v32[n] = (t >= 0) ? t : 0;
    }
}

void orc_conv1(const float* y, unsigned w, unsigned h, float* out)
{
    size_t n = (size_t)w * h;
    #pragma omp parallel for
    for (unsigned r = 0; r < h; r++)
        for (unsigned c = 0; c < w; c++) {
            float v[64]; conv1_at(y, (int)w, (int)h, (int)r, (int)c, v);
            for (int k = 0; k < 64; k++) out[k * n + (size_t)r * w + c] = v[k];
        }
}
void orc_conv2(const float* l1, unsigned w, unsigned h, float* out)
{
    size_t n = (size_t)w * h;
    #pragma omp parallel for
    for (size_t i = 0; i < n; i++) {
        float a[64], b[32];
        for (int k = 0; k < 64; k++) a[k] = l1[k * n + i];
        conv2_at(a, b);
        for (int m = 0; m < 32; m++) out[m * n + i] = b[m];
    }
}
/* conv III.  src/libsrcnn.cpp:449-529: each of the 32 planes clamp-padded by 2 (clamps the conv II plane's
 * own coordinates); per channel a double partial over 25 taps (float product promoted), y (row) outer,
 * x (col) inner, weight kernel[ch][x][y] (our SRCNN_W3[dy*5+dx][ch] already resolves that transposition);
 * float running sum over channels, + bias, clamp to [0,255]. */
void orc_conv3(const float* l2, unsigned w, unsigned h, float* out)
{
    size_t n = (size_t)w * h;
    #pragma omp parallel for
    for (unsigned r = 0; r < h; r++)
        for (unsigned c = 0; c < w; c++) {
            float t = 0;
            for (int ch = 0; ch < 32; ch++) {
                double tp = 0;
                for (int dy = 0; dy < 5; dy++) {
                    int rr = clampi((int)r + dy - 2, 0, (int)h - 1);
                    for (int dx = 0; dx < 5; dx++) {
                        int cc = clampi((int)c + dx - 2, 0, (int)w - 1);
                        float prod = SRCNN_W3[dy * 5 + dx][ch] * l2[ch * n + (size_t)rr * w + cc];
                        tp += prod;
                    }
                }
                t += tp;                   /* float += double: evaluated in double, rounded to float */
            }
            t += SRCNN_B3;
            t = (t > 0.f) ? t : 0.f;       /* MAX(temp,0.f), src/minmax.h:9 */
            t = (t < 255.f) ? t : 255.f;   /* MIN(temp,255.f) */
            out[(size_t)r * w + c] = t;
        }
}

/* ------------------------------------------------------------------ colour out
 * src/libsrcnn.cpp:274-308 (convertImgF32XtoImgU8): MIN then MAX then truncating cast. */
void orc_color_out(const float* const* p, unsigned w, unsigned h, unsigned d, unsigned char* out)
{
    size_t n = (size_t)w * h;
    #pragma omp parallel for
    for (size_t i = 0; i < n; i++) {
        float fy = p[0][i], fcb = p[1][i] - 128.f, fcr = p[2][i] - 128.f;
        float t;
        t = fy + ((45.f * fcr) / 32.f);                       float fr = (255.f < t) ? 255.f : t;
        t = fy - (((11.f * fcb) + (23.f * fcr)) / 32.f);      float fg = (255.f < t) ? 255.f : t;
        t = fy + ((113.f * fcb) / 64.f);                      float fb = (255.f < t) ? 255.f : t;
        out[i * d + 0] = (unsigned char)((0.f > fr) ? 0.f : fr);
        out[i * d + 1] = (unsigned char)((0.f > fg) ? 0.f : fg);
        out[i * d + 2] = (unsigned char)((0.f > fb) ? 0.f : fb);
        if (d == 4) {
            float fa = (255.f < p[3][i]) ? 255.f : p[3][i];
            out[i * d + 3] = (unsigned char)((0.f > fa) ? 0.f : fa);
        }
    }
}

/* ------------------------------------------------------------------ one pass: reference doSRCNN
 * src/libsrcnn.cpp:628-923.  Plane filters: Y = configured (out-of-range -> box, :677-699);
 * Cb/Cr/A = bilinear unless configured nearest -> box (:703-713). */
int orc_pass(const unsigned char* src, unsigned w, unsigned h, unsigned d, float multiply, int filter,
             unsigned char** out, unsigned* outsz, unsigned char** conv, unsigned* convsz,
             float* y_plane, float* yprime)
{
    unsigned ow, oh; orc_output_dims(w, h, multiply, &ow, &oh);
    if (d != 3 && d != 4) return -1;
    if (!ow || !oh) return -2;
    size_t ns = (size_t)w * h, nd = (size_t)ow * oh;
    float* sp[4] = {0, 0, 0, 0}; float* rp[4] = {0, 0, 0, 0};
    for (unsigned i = 0; i < d; i++) { sp[i] = (float*)malloc(ns * sizeof(float)); rp[i] = (float*)malloc(nd * sizeof(float)); }
    orc_color_in(src, w, h, d, sp[0], sp[1], sp[2], sp[3]);
    int fy = (filter >= ORC_NEAREST && filter <= ORC_BSPLINE) ? filter : ORC_NEAREST;
    int fc = (filter == ORC_NEAREST) ? ORC_NEAREST : ORC_BILINEAR;
    for (unsigned i = 0; i < d; i++) { orc_resize(sp[i], w, h, ow, oh, i == 0 ? fy : fc, rp[i]); free(sp[i]); }
    if (y_plane) memcpy(y_plane, rp[0], nd * sizeof(float));

    float* l2 = (float*)malloc(nd * 32 * sizeof(float));
    #pragma omp parallel for schedule(dynamic, 4)
    for (unsigned r = 0; r < oh; r++)
        for (unsigned c = 0; c < ow; c++) {
            float a[64], b[32];
            conv1_at(rp[0], (int)ow, (int)oh, (int)r, (int)c, a);
            conv2_at(a, b);
            for (int m = 0; m < 32; m++) l2[m * nd + (size_t)r * ow + c] = b[m];
        }
    orc_conv3(l2, ow, oh, rp[0]);          /* conv III result replaces Y, src/libsrcnn.cpp:855-857 */
    free(l2);
    if (yprime) memcpy(yprime, rp[0], nd * sizeof(float));

    *outsz = (unsigned)(nd * d);
    *out = (unsigned char*)malloc(nd * d);
    orc_color_out((const float* const*)rp, ow, oh, d, *out);
    if (conv && convsz) {                  /* src/libsrcnn.cpp:889-915: truncating cast of the clamped plane */
        unsigned char* cbuf = (unsigned char*)malloc(nd);
        for (size_t i = 0; i < nd; i++) cbuf[i] = (unsigned char)rp[0][i];
        *conv = cbuf; *convsz = (unsigned)nd;
    }
    for (unsigned i = 0; i < d; i++) free(rp[i]);
    return 0;
}

/* ------------------------------------------------------------------ public entry: reference ProcessSRCNN
 * src/libsrcnn.cpp:943-1064 including the step-scale loop (:980-1061): repeat = (int)(m/2) (+1 if
 * fmodf(m,2) > 0); every pass x2 except the last = (w*m)/current_w; the last pass alone yields the conv
 * plane; a last factor of 0 or 1 stops the loop (then nothing new is produced for that pass). */
int orc_process(const unsigned char* src, unsigned w, unsigned h, unsigned d, float multiply,
                int filter, int stepscale,
                unsigned char** out, unsigned* outsz, unsigned char** conv, unsigned* convsz)
{
    if (!src || !w || !h || !d) return -1;
    if ((float)w * multiply <= 0.f || (float)h * multiply <= 0.f) return -2;
    if (!stepscale) return orc_pass(src, w, h, d, multiply, filter, out, outsz, conv, convsz, NULL, NULL);

    int rc = -100;
    float lf = fmodf(multiply, 2.f);
    int repeat = (int)(multiply / 2.f);
    if (lf > 0.f) repeat++;
    const unsigned char* cur = src; unsigned char* prev_owned = NULL;
    unsigned char* o = NULL; unsigned osz = 0; unsigned sw = w, sh = h;
    for (int cnt = 0; cnt < repeat; cnt++) {
        float m = 2.0f; int last = (cnt + 1 == repeat);
        if (last) { m = ((float)w * multiply) / (float)sw; if (m == 0.f || m == 1.0f) break; }
        if (cnt > 0) { free(prev_owned); prev_owned = o; cur = o; o = NULL; }
        rc = orc_pass(cur, sw, sh, d, m, filter, &o, &osz, last ? conv : NULL, last ? convsz : NULL, NULL, NULL);
        if (rc != 0) break;
        if (repeat > 1) { sw = (unsigned)((float)sw * m); sh = (unsigned)((float)sh * m); }
    }
    free(prev_owned);
    *out = o; *outsz = osz;
    return rc;
}

void orc_free(void* p) { free(p); }
