/* Host-side contribution tables for the separable resampler (double precision). */
#ifndef SRCNN_B200_CONTRIB_H
#define SRCNN_B200_CONTRIB_H
#include <vector>

namespace srcnn {

enum FilterKind { kNearest = 0, kBilinear = 1, kBicubic = 2, kLanczos3 = 3, kBspline = 4 };

struct AxisTable {
    int filter = 0; unsigned dst = 0, src = 0;
    int win = 0;                     /* weights per destination index (row pitch of w) */
    std::vector<int>    left, right; /* inclusive source range per destination index   */
    std::vector<double> w;           /* [dst][win], zero padded                        */
    /* dense form (only when src >= win): every destination index reads exactly `win` in-bounds source samples
     * starting at dleft; taps outside [left,right] carry weight 0 (leading/trailing zero terms leave the
     * ascending-order double accumulation unchanged), so kernels need no per-sample tap count */
    bool dense = false;
    std::vector<int>    dleft;
    std::vector<double> dw;
    /* compact form (only when src >= cwin): like the dense form, but the window is cut down to the widest run of
     * non-zero weights over all destination indices, rounded up to an even size (2 for box / bilinear, 4 for the
     * cubics, 6 for Lanczos-3 when magnifying).  Taps with weight exactly 0 add +-0 to the accumulator, i.e. nothing,
     * so leaving them out is bit-identical.  exact29: every weight has at most 29 significant bits, so its product
     * with an FP32 sample is exact in double and a fused multiply-add equals the reference's multiply then add. */
    bool compact = false, exact29 = false;
    /* pow2pair: compact window of 2, both weights exactly representable in FP32 and at least one of each pair 0 or a power
     * of two (bilinear at x2: 0.25 / 0.75): fmaf(wA, vA, wB * vB) in FP32 is then the correctly rounded exact value */
    bool pow2pair = false;
    int cwin = 0;
    std::vector<int>    cleft;
    std::vector<double> cw;
};

double filter_support(int filter);
double filter_value(int filter, double x);
int    axis_window(int filter, unsigned dst, unsigned src);
void   build_axis_table(int filter, unsigned dst, unsigned src, AxisTable& t);

} // namespace srcnn
#endif
