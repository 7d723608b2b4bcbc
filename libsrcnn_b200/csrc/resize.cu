/* Separable resampling kernels (HBM-bound) + the conv-input pack.
 *
 * = reference FRAWResizeEngine::verticalFilter / horizontalFilter (src/frawscale.cpp:288-385) driven by the
 *   contribution tables of FRawScaleWeightsTable (src/frawscale.cpp:8-112; rebuilt on the host in contrib.cpp).
 *
 * Arithmetic contract (bit-exact against the oracle): per destination sample a double accumulator starting at
 * 0.0, taps in ascending source order, each step  acc = acc + (w * (double)px)  with separately rounded multiply
 * and add (the reference's x86-64 build does not contract to FMA), then one rounding to float.  The tmp plane
 * between the two passes is FP32, exactly like the reference's.
 */
#include "kernels.h"

namespace srcnn {

/* vertical: thread = 4 adjacent columns (float4) or 1 column, one destination row */
template <int VEC>
__global__ void __launch_bounds__(256)
resize_vertical_kernel(const float* __restrict__ in, size_t in_plane_stride, unsigned width, unsigned in_row0,
                       float* __restrict__ out, size_t out_plane_stride, unsigned out_row0, unsigned out_rows,
                       AxisTableDev ty, AxisTableDev tc)
{
    const unsigned xv = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned yy = blockIdx.y;
    const unsigned pl = blockIdx.z;
    const unsigned x  = xv * VEC;
    if (x >= width || yy >= out_rows) return;
    const AxisTableDev t = (pl == 0) ? ty : tc;
    const unsigned gy = out_row0 + yy;
    const int left = __ldg(t.left + gy), cnt = __ldg(t.count + gy);
    const double* w = t.w + (size_t)gy * t.win;
    const float* src = in + pl * in_plane_stride + (size_t)(left - (int)in_row0) * width + x;
    float* dst = out + pl * out_plane_stride + (size_t)yy * width + x;

    double acc[VEC];
    #pragma unroll
    for (int v = 0; v < VEC; v++) acc[v] = 0.0;
    for (int i = 0; i < cnt; i++) {
        const double wi = __ldg(w + i);
        if (VEC == 4) {
            const float4 p = __ldg(reinterpret_cast<const float4*>(src + (size_t)i * width));
            acc[0] = __dadd_rn(acc[0], __dmul_rn(wi, (double)p.x));
            acc[1] = __dadd_rn(acc[1], __dmul_rn(wi, (double)p.y));
            acc[2] = __dadd_rn(acc[2], __dmul_rn(wi, (double)p.z));
            acc[VEC - 1] = __dadd_rn(acc[VEC - 1], __dmul_rn(wi, (double)p.w));
        } else {
            acc[0] = __dadd_rn(acc[0], __dmul_rn(wi, (double)__ldg(src + (size_t)i * width)));
        }
    }
    if (VEC == 4)
        *reinterpret_cast<float4*>(dst) = make_float4(__double2float_rn(acc[0]), __double2float_rn(acc[1]),
                                                       __double2float_rn(acc[2]), __double2float_rn(acc[VEC - 1]));
    else
        *dst = __double2float_rn(acc[0]);
}

/* vertical, dense table with exactly W taps (every magnification): thread = 4 adjacent columns, VROWS destination rows
 * are NOT shared (each destination row has its own taps); no tap count, no predicates */
template <int W>
__global__ void __launch_bounds__(128)
resize_vertical_dense_kernel(const float* __restrict__ in, size_t in_plane_stride, unsigned width, unsigned in_row0,
                             float* __restrict__ out, size_t out_plane_stride, unsigned out_row0, unsigned out_rows,
                             const int* __restrict__ dleft, const double* __restrict__ dw)
{
    const unsigned x  = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const unsigned yy = blockIdx.y;
    const unsigned pl = blockIdx.z;
    if (x >= width) return;
    const unsigned gy = out_row0 + yy;
    const int left = __ldg(dleft + gy);
    const double* w = dw + (size_t)gy * W;
    const float* src = in + pl * in_plane_stride + (size_t)(left - (int)in_row0) * width + x;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    #pragma unroll
    for (int i = 0; i < W; i++) {
        const double wi = __ldg(w + i);
        const float4 p = __ldg(reinterpret_cast<const float4*>(src + (size_t)i * width));
        a0 = __dadd_rn(a0, __dmul_rn(wi, (double)p.x));
        a1 = __dadd_rn(a1, __dmul_rn(wi, (double)p.y));
        a2 = __dadd_rn(a2, __dmul_rn(wi, (double)p.z));
        a3 = __dadd_rn(a3, __dmul_rn(wi, (double)p.w));
    }
    *reinterpret_cast<float4*>(out + pl * out_plane_stride + (size_t)yy * width + x) =
        make_float4(__double2float_rn(a0), __double2float_rn(a1), __double2float_rn(a2), __double2float_rn(a3));
}

template <int W>
static void launch_vdense(const float* in, size_t ips, unsigned width, unsigned in_row0, float* out, size_t ops,
                          unsigned out_row0, unsigned out_rows, unsigned nplanes, const AxisTableDev& t, cudaStream_t s)
{
    dim3 grid((width / 4 + 127) / 128, out_rows, nplanes);
    resize_vertical_dense_kernel<W><<<grid, 128, 0, s>>>(in, ips, width, in_row0, out, ops, out_row0, out_rows, t.dleft, t.dw);
}
static bool vdense_dispatch(const float* in, size_t ips, unsigned width, unsigned in_row0, float* out, size_t ops,
                            unsigned out_row0, unsigned out_rows, unsigned nplanes, const AxisTableDev& t, cudaStream_t s)
{
    switch (t.win) {
        case 3: launch_vdense<3>(in, ips, width, in_row0, out, ops, out_row0, out_rows, nplanes, t, s); return true;
        case 5: launch_vdense<5>(in, ips, width, in_row0, out, ops, out_row0, out_rows, nplanes, t, s); return true;
        case 7: launch_vdense<7>(in, ips, width, in_row0, out, ops, out_row0, out_rows, nplanes, t, s); return true;
        default: return false;
    }
}

void launch_resize_vertical(const float* in, size_t in_plane_stride, unsigned width, unsigned in_row0,
                            float* out, size_t out_plane_stride, unsigned out_row0, unsigned out_rows,
                            unsigned nplanes, AxisTableDev ty, AxisTableDev tc, cudaStream_t s)
{
    const bool vec = (width % 4 == 0) && (in_plane_stride % 4 == 0) && (out_plane_stride % 4 == 0) &&
                     ((uintptr_t)in % 16 == 0) && ((uintptr_t)out % 16 == 0);
    const bool dense = vec && ty.dleft && tc.dleft && (ty.win == 3 || ty.win == 5 || ty.win == 7) &&
                       (tc.win == 3 || tc.win == 5 || tc.win == 7);
    if (dense) {
        /* plane 0 with its own filter, then the remaining planes together */
        vdense_dispatch(in, in_plane_stride, width, in_row0, out, out_plane_stride, out_row0, out_rows, 1, ty, s);
        if (nplanes > 1)
            vdense_dispatch(in + in_plane_stride, in_plane_stride, width, in_row0, out + out_plane_stride, out_plane_stride,
                            out_row0, out_rows, nplanes - 1, tc, s);
    } else if (vec) {
        dim3 grid((width / 4 + 127) / 128, out_rows, nplanes);
        resize_vertical_kernel<4><<<grid, 128, 0, s>>>(in, in_plane_stride, width, in_row0, out, out_plane_stride,
                                                       out_row0, out_rows, ty, tc);
    } else {
        dim3 grid((width + 255) / 256, out_rows, nplanes);
        resize_vertical_kernel<1><<<grid, 256, 0, s>>>(in, in_plane_stride, width, in_row0, out, out_plane_stride,
                                                       out_row0, out_rows, ty, tc);
    }
}

/* horizontal, generic (any window, e.g. minification): thread = one destination sample */
__global__ void __launch_bounds__(256)
resize_horizontal_kernel(const float* __restrict__ in, size_t in_plane_stride, unsigned sw,
                         float* __restrict__ out, size_t out_plane_stride, unsigned dw, unsigned rows,
                         AxisTableDev ty, AxisTableDev tc)
{
    const unsigned x  = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned yy = blockIdx.y;
    const unsigned pl = blockIdx.z;
    if (x >= dw || yy >= rows) return;
    const AxisTableDev t = (pl == 0) ? ty : tc;
    const int left = __ldg(t.left + x), cnt = __ldg(t.count + x);
    const double* w = t.w + (size_t)x * t.win;
    const float* src = in + pl * in_plane_stride + (size_t)yy * sw + left;
    double acc = 0.0;
    for (int i = 0; i < cnt; i++)
        acc = __dadd_rn(acc, __dmul_rn(__ldg(w + i), (double)__ldg(src + i)));
    out[pl * out_plane_stride + (size_t)yy * dw + x] = __double2float_rn(acc);
}

/* horizontal, dense table with exactly W taps: thread = one destination column, HROWS rows.  The column's taps are
 * loaded once into registers and reused down the rows; inner loop = one pixel load + F2F + DMUL + DADD per tap. */
constexpr int HROWS = 16;
template <int W>
__global__ void __launch_bounds__(128)
resize_horizontal_dense_kernel(const float* __restrict__ in, size_t in_plane_stride, unsigned sw,
                               float* __restrict__ out, size_t out_plane_stride, unsigned dw, unsigned rows,
                               const int* __restrict__ dleft, const double* __restrict__ dwt)
{
    const unsigned x  = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned y0 = blockIdx.y * HROWS;
    const unsigned pl = blockIdx.z;
    if (x >= dw) return;
    const int left = __ldg(dleft + x);
    double w[W];
    #pragma unroll
    for (int i = 0; i < W; i++) w[i] = __ldg(dwt + (size_t)x * W + i);
    const float* src = in + pl * in_plane_stride + (size_t)y0 * sw + left;
    float* dst = out + pl * out_plane_stride + (size_t)y0 * dw + x;
    const unsigned ny = min((unsigned)HROWS, rows - y0);
    for (unsigned r = 0; r < ny; r++, src += sw, dst += dw) {
        double acc = 0.0;
        #pragma unroll
        for (int i = 0; i < W; i++) acc = __dadd_rn(acc, __dmul_rn(w[i], (double)__ldg(src + i)));
        *dst = __double2float_rn(acc);
    }
}

template <int W>
static void launch_hdense(const float* in, size_t ips, unsigned sw, float* out, size_t ops, unsigned dw, unsigned rows,
                          unsigned nplanes, const AxisTableDev& t, cudaStream_t s)
{
    dim3 grid((dw + 127) / 128, (rows + HROWS - 1) / HROWS, nplanes);
    resize_horizontal_dense_kernel<W><<<grid, 128, 0, s>>>(in, ips, sw, out, ops, dw, rows, t.dleft, t.dw);
}
static void hdense_dispatch(const float* in, size_t ips, unsigned sw, float* out, size_t ops, unsigned dw, unsigned rows,
                            unsigned nplanes, const AxisTableDev& t, cudaStream_t s)
{
    switch (t.win) {
        case 3: launch_hdense<3>(in, ips, sw, out, ops, dw, rows, nplanes, t, s); break;
        case 5: launch_hdense<5>(in, ips, sw, out, ops, dw, rows, nplanes, t, s); break;
        default: launch_hdense<7>(in, ips, sw, out, ops, dw, rows, nplanes, t, s); break;
    }
}

void launch_resize_horizontal(const float* in, size_t in_plane_stride, unsigned sw,
                              float* out, size_t out_plane_stride, unsigned dw, unsigned rows,
                              unsigned nplanes, AxisTableDev ty, AxisTableDev tc, cudaStream_t s)
{
    const bool dense = ty.dleft && tc.dleft && (ty.win == 3 || ty.win == 5 || ty.win == 7) &&
                       (tc.win == 3 || tc.win == 5 || tc.win == 7);
    if (dense) {
        hdense_dispatch(in, in_plane_stride, sw, out, out_plane_stride, dw, rows, 1, ty, s);
        if (nplanes > 1)
            hdense_dispatch(in + in_plane_stride, in_plane_stride, sw, out + out_plane_stride, out_plane_stride, dw, rows,
                            nplanes - 1, tc, s);
    } else {
        dim3 grid((dw + 255) / 256, rows, nplanes);
        resize_horizontal_kernel<<<grid, 256, 0, s>>>(in, in_plane_stride, sw, out, out_plane_stride, dw, rows, ty, tc);
    }
}

/* ------------------------------------------------------------------------------------------------------------ */

Y16Geom y16_geometry(unsigned dw, unsigned strip_rows)
{
    Y16Geom g;
    g.bands = (int)((dw + kBandCols - 1) / kBandCols);
    /* last band's TMA boxes end at column 124*bands + 2 + 144 + 1 (see conv_tc.cu); round to 128 B */
    const int need = kBandCols * g.bands + 2 + 144 + 2;
    g.pitch = (need + 63) / 64 * 64;
    g.rows  = (int)strip_rows + 2 * kHaloRows;
    return g;
}

/* thread = 8 consecutive halves (one 16 B store) of the replicate-padded FP16 plane */
__global__ void __launch_bounds__(256)
pack_y16_kernel(const float* __restrict__ y, int dw, int dh, int win_row0, __half* __restrict__ y16,
                int pitch, int rows, int strip_row0)
{
    const int xo = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
    const int r  = blockIdx.y;
    if (xo >= pitch || r >= rows) return;
    int g = strip_row0 - kHaloRows + r;                   /* global row, replicate-clamped */
    g = g < 0 ? 0 : (g > dh - 1 ? dh - 1 : g);
    const float* src = y + (size_t)(g - win_row0) * dw;
    __half2 h[4];
    #pragma unroll
    for (int k = 0; k < 4; k++) {
        int c0 = xo + 2 * k - kY16PadL, c1 = c0 + 1;
        c0 = c0 < 0 ? 0 : (c0 > dw - 1 ? dw - 1 : c0);
        c1 = c1 < 0 ? 0 : (c1 > dw - 1 ? dw - 1 : c1);
        h[k] = __floats2half2_rn(__ldg(src + c0), __ldg(src + c1));
    }
    uint4 v;
    v.x = *reinterpret_cast<uint32_t*>(&h[0]); v.y = *reinterpret_cast<uint32_t*>(&h[1]);
    v.z = *reinterpret_cast<uint32_t*>(&h[2]); v.w = *reinterpret_cast<uint32_t*>(&h[3]);
    *reinterpret_cast<uint4*>(y16 + (size_t)r * pitch + xo) = v;
}

void launch_pack_y16(const float* y, unsigned dw, unsigned dh, unsigned win_row0, __half* y16, Y16Geom g,
                     unsigned strip_row0, cudaStream_t s)
{
    dim3 grid((g.pitch / 8 + 255) / 256, g.rows);
    pack_y16_kernel<<<grid, 256, 0, s>>>(y, (int)dw, (int)dh, (int)win_row0, y16, g.pitch, g.rows, (int)strip_row0);
}

} // namespace srcnn
