/* Contribution tables = reference FRawScaleWeightsTable ctor (src/frawscale.cpp:8-112) with the filter
 * kernels of src/frawscale.h:58-172, evaluated in double on the host; a table is a few KB per axis and is
 * cached per (filter, dst, src) by the context, then uploaded once.
 *
 * Compiled with -ffp-contract=off: the expressions below must round exactly like the reference's x86-64 build,
 * because the device kernels reproduce its resampled planes bit for bit.
 */
#include "contrib.h"
#include <cmath>
#include <cstdint>
#include <cstring>

namespace srcnn {

double filter_support(int f)
{
    switch (f) {
        case kBilinear: return 1.0;
        case kBicubic:  return 2.0;
        case kLanczos3: return 3.0;
        case kBspline:  return 2.0;
        default:        return 0.5;     /* box; also any out-of-range filter id (src/libsrcnn.cpp:679-682) */
    }
}

static double sinc_pi(double v)
{
    if (v != 0) { v *= 3.1415926535897932384626433832795; return std::sin(v) / v; }
    return 1;
}

double filter_value(int f, double x)
{
    x = std::fabs(x);
    switch (f) {
    case kBilinear:
        return x < 1.0 ? 1.0 - x : 0.0;
    case kBicubic: {    /* Mitchell-Netravali, B = C = 1/3 */
        const double b = 1 / (double)3, c = 1 / (double)3;
        const double p0 = (6 - 2 * b) / 6, p2 = (-18 + 12 * b + 6 * c) / 6, p3 = (12 - 9 * b - 6 * c) / 6;
        const double q0 = (8 * b + 24 * c) / 6, q1 = (-12 * b - 48 * c) / 6, q2 = (6 * b + 30 * c) / 6,
                     q3 = (-b - 6 * c) / 6;
        if (x < 1) return p0 + x * x * (p2 + x * p3);
        if (x < 2) return q0 + x * (q1 + x * (q2 + x * q3));
        return 0; }
    case kLanczos3:
        return x < 3.0 ? sinc_pi(x) * sinc_pi(x / 3.0) : 0.0;
    case kBspline:
        if (x < 1) return (4 + x * x * (-6 + 3 * x)) / 6;
        if (x < 2) { const double t = 2 - x; return t * t * t / 6; }
        return 0;
    default:
        return x <= 0.5 ? 1.0 : 0.0;
    }
}

int axis_window(int filter, unsigned dst, unsigned src)
{
    const double scale = (double)dst / (double)src;
    double width = filter_support(filter);
    if (scale < 1.0) width = width / scale;         /* minification widens the kernel */
    return 2 * (int)std::ceil(width) + 1;
}

void build_axis_table(int filter, unsigned dst, unsigned src, AxisTable& t)
{
    t.filter = filter; t.dst = dst; t.src = src;
    const double scale = (double)dst / (double)src;
    double width = filter_support(filter), fscale = 1.0;
    if (scale < 1.0) { width = width / scale; fscale = scale; }
    t.win = 2 * (int)std::ceil(width) + 1;
    t.left.assign(dst, 0); t.right.assign(dst, 0); t.w.assign((size_t)dst * t.win, 0.0);
    const double offset = (0.5 / scale) - 0.5;      /* centre of dst sample 0 in source coordinates */
    const int last = (int)src - 1;
    for (unsigned u = 0; u < dst; u++) {
        double* wu = &t.w[(size_t)u * t.win];
        const double centre = (double)u / scale + offset;
        int lo = (int)std::floor(centre - width), hi = (int)std::ceil(centre + width);
        if (lo < 0) lo = 0;
        if (hi > last) hi = last;
        /* one sample too many: the reference always drops the leftmost one (its alternative branch compares
         * against src - 1/2 == src in integer arithmetic and is unreachable, src/frawscale.cpp:55-66) */
        if (hi - lo + 1 > t.win) { if (lo < (int)src) lo++; else hi--; }
        double total = 0;
        for (int i = lo; i <= hi; i++) {
            const double wt = fscale * filter_value(filter, fscale * (centre - (double)i));
            wu[i - lo] = wt; total += wt;
        }
        if (total > 0 && total != 1) {              /* edges truncate the window and renormalise */
            for (int i = lo; i <= hi; i++) wu[i - lo] /= total;
            int k = hi - lo;
            while (wu[k] == 0) { hi--; k--; if (hi == lo) break; }   /* trim trailing zero taps */
        }
        t.left[u] = lo; t.right[u] = hi;
    }
    t.dense = (int)src >= t.win;
    if (t.dense) {
        t.dleft.assign(dst, 0); t.dw.assign((size_t)dst * t.win, 0.0);
        for (unsigned u = 0; u < dst; u++) {
            int dl = t.left[u];
            if (dl + t.win > (int)src) dl = (int)src - t.win;      /* shift the window left, zero weights in front */
            t.dleft[u] = dl;
            for (int i = t.left[u]; i <= t.right[u]; i++) t.dw[(size_t)u * t.win + (i - dl)] = t.w[(size_t)u * t.win + (i - t.left[u])];
        }
    }
    /* compact form */
    int widest = 1;
    std::vector<int> first(dst), lastnz(dst);
    t.exact29 = true;
    for (unsigned u = 0; u < dst; u++) {
        const double* wu = &t.w[(size_t)u * t.win];
        int a = t.left[u], b = t.right[u];
        while (a < b && wu[a - t.left[u]] == 0) a++;
        while (b > a && wu[b - t.left[u]] == 0) b--;
        first[u] = a; lastnz[u] = b;
        widest = widest > b - a + 1 ? widest : b - a + 1;
        for (int i = a; i <= b; i++) {
            uint64_t bits; const double v = wu[i - t.left[u]];
            memcpy(&bits, &v, 8);
            if (v != 0 && (bits & ((1ull << 24) - 1)) != 0) t.exact29 = false;      /* 52 - 28 low mantissa bits clear */
        }
    }
    t.cwin = (widest + 1) / 2 * 2;
    t.compact = (int)src >= t.cwin;
    t.pow2pair = false;
    if (t.compact) {
        t.cleft.assign(dst, 0); t.cw.assign((size_t)dst * t.cwin, 0.0);
        for (unsigned u = 0; u < dst; u++) {
            int cl = first[u];
            if (cl + t.cwin > (int)src) cl = (int)src - t.cwin;
            t.cleft[u] = cl;
            for (int i = first[u]; i <= lastnz[u]; i++) t.cw[(size_t)u * t.cwin + (i - cl)] = t.w[(size_t)u * t.win + (i - t.left[u])];
        }
        if (t.cwin == 2) {
            auto f32_exact = [](double v) { return (double)(float)v == v; };
            auto pow2_or_zero = [](double v) { uint64_t b; memcpy(&b, &v, 8); return v == 0 || (v > 0 && (b & ((1ull << 52) - 1)) == 0); };
            t.pow2pair = true;
            for (unsigned u = 0; u < dst && t.pow2pair; u++) {
                const double a = t.cw[2 * (size_t)u], b = t.cw[2 * (size_t)u + 1];
                if (!f32_exact(a) || !f32_exact(b) || !(pow2_or_zero(a) || pow2_or_zero(b)) || a < 0 || b < 0) t.pow2pair = false;
            }
        }
    }
}

} // namespace srcnn
