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
};

double filter_support(int filter);
double filter_value(int filter, double x);
int    axis_window(int filter, unsigned dst, unsigned src);
void   build_axis_table(int filter, unsigned dst, unsigned src, AxisTable& t);

} // namespace srcnn
#endif
