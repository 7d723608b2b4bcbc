/* jpegdec.cpp -- baseline JPEG decoding for the command-line harness (imgio.cpp).
 *
 * The reference CLI loads JPEG files through FLTK's Fl_JPEG_Image (src/test.cpp:145-148, 563-566), i.e. libjpeg with its
 * defaults; its sample picture Pictures/castle.jpg is a baseline 4:2:0 file.  This is a from-scratch decoder for that class of
 * files: sequential DCT (SOF0 / SOF1), Huffman coding, 8-bit samples, 1 (gray) or 3 (YCbCr) components, sampling factors 1 or 2,
 * restart intervals.  Progressive / arithmetic-coded / 12-bit / CMYK files are refused with a message.  Chroma is upsampled
 * with libjpeg's "fancy" triangle filter and converted with its fixed-point YCbCr -> RGB tables, so the result is within a couple
 * of LSBs of a libjpeg-based loader (the IDCT here is a plain double-precision one; tests/test_cli.py compares with PIL).
 * Pure host code, no CUDA.
 */
#include "imgio.h"

#include <cmath>
#include <cstring>
#include <stdexcept>

namespace imgio {

namespace {

struct Huff {
    unsigned char bits[17] = {0};
    unsigned char vals[256] = {0};
    int mincode[17], maxcode[18], valptr[17];
    bool present = false;
    void build()
    {
        int code = 0, k = 0;
        for (int l = 1; l <= 16; l++) {
            valptr[l] = k; mincode[l] = code;
            code += bits[l]; k += bits[l];
            maxcode[l] = bits[l] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
        present = true;
    }
};

struct Comp { int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0, pred = 0, bw = 0, bh = 0; std::vector<unsigned char> plane; };

struct BitReader {
    const unsigned char* p; size_t n, pos; unsigned buf = 0; int cnt = 0; bool hit_marker = false;
    int bit()
    {
        if (cnt == 0) {
            unsigned b = 0;
            if (pos < n && !hit_marker) {
                b = p[pos++];
                if (b == 0xFF) {
                    const unsigned c = pos < n ? p[pos] : 0xD9;
                    if (c == 0) pos++;                       /* stuffed zero */
                    else { hit_marker = true; pos--; b = 0; } /* a marker: feed zeros, leave it for the caller */
                }
            }
            buf = b; cnt = 8;
        }
        cnt--;
        return (buf >> cnt) & 1;
    }
    int bits(int k) { int v = 0; while (k--) v = (v << 1) | bit(); return v; }
    void reset() { cnt = 0; hit_marker = false; }
};

int decode_sym(BitReader& br, const Huff& h)
{
    int code = 0;
    for (int l = 1; l <= 16; l++) {
        code = (code << 1) | br.bit();
        if (h.maxcode[l] >= 0 && code <= h.maxcode[l] && code >= h.mincode[l]) return h.vals[h.valptr[l] + code - h.mincode[l]];
    }
    throw std::runtime_error("bad Huffman code");
}

inline int extend(int v, int t) { return v < (1 << (t - 1)) ? v - (1 << t) + 1 : v; }

const unsigned char kZigzag[64] = { 0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                                    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63 };

/* 8x8 inverse DCT of dequantised coefficients (row-major), + 128, clamped -> out with the given pitch */
void idct8x8(const int* coef, unsigned char* out, int pitch)
{
    static double c[8][8]; static bool init = false;
    if (!init) {
        for (int x = 0; x < 8; x++) for (int u = 0; u < 8; u++) c[x][u] = (u ? 1.0 : std::sqrt(0.5)) * std::cos((2 * x + 1) * u * M_PI / 16.0) * 0.5;
        init = true;
    }
    double tmp[64];
    for (int v = 0; v < 8; v++)                       /* rows: along u */
        for (int x = 0; x < 8; x++) {
            double s = 0.0;
            for (int u = 0; u < 8; u++) s += c[x][u] * coef[v * 8 + u];
            tmp[v * 8 + x] = s;
        }
    for (int x = 0; x < 8; x++)
        for (int y = 0; y < 8; y++) {
            double s = 0.0;
            for (int v = 0; v < 8; v++) s += c[y][v] * tmp[v * 8 + x];
            const int q = (int)std::floor(s + 128.5);
            out[y * pitch + x] = (unsigned char)(q < 0 ? 0 : q > 255 ? 255 : q);
        }
}

unsigned be16(const unsigned char* p) { return (unsigned)p[0] << 8 | p[1]; }

/* libjpeg's "fancy" (triangle) upsampling by 2 along a row: in[w] -> out[2w] */
void up_h2(const int* in, int w, int* out, int round_lo, int round_hi, int shift)
{
    /* generic form used for both the h2v1 case (in = samples, weights 3:1, >> 2) and the second stage of h2v2 (in = column sums) */
    for (int i = 0; i < w; i++) {
        const int prev = in[i > 0 ? i - 1 : i], next = in[i + 1 < w ? i + 1 : i];
        if (w == 1) { out[0] = out[1] = (in[0] * 4 + round_lo) >> shift; break; }
        out[2 * i]     = i == 0 ? (in[i] * 4 + round_lo) >> shift : (3 * in[i] + prev + round_lo) >> shift;
        out[2 * i + 1] = i + 1 == w ? (in[i] * 4 + round_hi) >> shift : (3 * in[i] + next + round_hi) >> shift;
    }
}

} // namespace

bool decode_jpeg(const unsigned char* d, size_t n, Image& img, std::string& err)
{
    try {
        unsigned short qt[4][64]; bool qt_ok[4] = { false, false, false, false };
        Huff hdc[4], hac[4];
        std::vector<Comp> comps;
        int W = 0, H = 0, restart = 0, hmax = 1, vmax = 1;
        size_t pos = 2;
        bool got_sof = false;
        for (;;) {
            while (pos < n && d[pos] != 0xFF) pos++;
            while (pos < n && d[pos] == 0xFF) pos++;
            if (pos >= n) { err = "JPEG: no scan found"; return false; }
            const unsigned m = d[pos++];
            if (m == 0xD8 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
            if (m == 0xD9) { err = "JPEG: no scan found"; return false; }
            if (pos + 2 > n) { err = "JPEG: truncated"; return false; }
            const unsigned L = be16(d + pos);
            if (L < 2 || pos + L > n) { err = "JPEG: truncated segment"; return false; }
            const unsigned char* s = d + pos + 2; const unsigned sl = L - 2;
            if (m == 0xDB) {                                              /* DQT */
                unsigned o = 0;
                while (o < sl) {
                    const unsigned pq = s[o] >> 4, tq = s[o] & 15; o++;
                    if (tq > 3 || o + (pq ? 128u : 64u) > sl) { err = "JPEG: bad quantisation table"; return false; }
                    for (int i = 0; i < 64; i++) { qt[tq][kZigzag[i]] = pq ? (unsigned short)be16(s + o + 2 * i) : s[o + i]; }
                    o += pq ? 128 : 64; qt_ok[tq] = true;
                }
            } else if (m == 0xC4) {                                       /* DHT */
                unsigned o = 0;
                while (o + 17 <= sl) {
                    const unsigned tc = s[o] >> 4, th = s[o] & 15;
                    if (tc > 1 || th > 3) { err = "JPEG: bad Huffman table"; return false; }
                    Huff& h = tc ? hac[th] : hdc[th];
                    unsigned cnt = 0;
                    for (int l = 1; l <= 16; l++) { h.bits[l] = s[o + l]; cnt += s[o + l]; }
                    if (cnt > 256 || o + 17 + cnt > sl) { err = "JPEG: bad Huffman table"; return false; }
                    memcpy(h.vals, s + o + 17, cnt);
                    h.build();
                    o += 17 + cnt;
                }
            } else if (m == 0xC0 || m == 0xC1) {                          /* SOF0 / SOF1: sequential, Huffman */
                if (sl < 6 || s[0] != 8) { err = "JPEG: only 8-bit samples are supported"; return false; }
                H = (int)be16(s + 1); W = (int)be16(s + 3);
                const int nc = s[5];
                if (!W || !H || (nc != 1 && nc != 3) || sl < 6u + 3u * nc) { err = "JPEG: unsupported layout (components / size)"; return false; }
                if ((size_t)W * H > ((size_t)1 << 28)) { err = "JPEG: image too large"; return false; }
                comps.resize(nc);
                for (int i = 0; i < nc; i++) {
                    comps[i].id = s[6 + 3 * i]; comps[i].h = s[7 + 3 * i] >> 4; comps[i].v = s[7 + 3 * i] & 15; comps[i].tq = s[8 + 3 * i];
                    if (comps[i].h < 1 || comps[i].h > 2 || comps[i].v < 1 || comps[i].v > 2 || comps[i].tq > 3) { err = "JPEG: unsupported sampling factors"; return false; }
                    hmax = std::max(hmax, comps[i].h); vmax = std::max(vmax, comps[i].v);
                }
                got_sof = true;
            } else if (m == 0xC2 || m == 0xC3 || (m >= 0xC5 && m <= 0xCF && m != 0xC8 && m != 0xCC)) {
                err = m == 0xC2 ? "JPEG: progressive files are not supported (baseline only); convert the file first"
                                : "JPEG: this coding process (lossless / arithmetic / hierarchical) is not supported";
                return false;
            } else if (m == 0xDD) {
                if (sl >= 2) restart = (int)be16(s);
            } else if (m == 0xDA) {                                       /* SOS */
                if (!got_sof) { err = "JPEG: scan before frame header"; return false; }
                const int ns = s[0];
                if (ns != (int)comps.size() || sl < 1u + 2u * ns + 3u) { err = "JPEG: multi-scan files are not supported"; return false; }
                for (int i = 0; i < ns; i++) {
                    const int cid = s[1 + 2 * i];
                    Comp* c = nullptr;
                    for (auto& k : comps) if (k.id == cid) c = &k;
                    if (!c) { err = "JPEG: scan refers to an unknown component"; return false; }
                    c->td = s[2 + 2 * i] >> 4; c->ta = s[2 + 2 * i] & 15;
                    if (c->td > 3 || c->ta > 3 || !hdc[c->td].present || !hac[c->ta].present || !qt_ok[c->tq]) { err = "JPEG: missing table"; return false; }
                }
                pos += L;
                break;
            }
            pos += L;
        }
        /* ---- entropy-coded data: MCUs of hmax*8 x vmax*8 pixels */
        const int mcuw = 8 * hmax, mcuh = 8 * vmax, mx = (W + mcuw - 1) / mcuw, my = (H + mcuh - 1) / mcuh;
        for (auto& c : comps) { c.bw = mx * c.h * 8; c.bh = my * c.v * 8; c.plane.assign((size_t)c.bw * c.bh, 0); c.pred = 0; }
        BitReader br{ d, n, pos };
        int rst_left = restart;
        for (int y = 0; y < my; y++)
            for (int x = 0; x < mx; x++) {
                if (restart && rst_left == 0) {
                    /* expect RSTn: skip to it, reset the predictors */
                    br.reset();
                    size_t q = br.pos;
                    while (q + 1 < n && !(d[q] == 0xFF && d[q + 1] >= 0xD0 && d[q + 1] <= 0xD7)) q++;
                    if (q + 1 >= n) { err = "JPEG: missing restart marker"; return false; }
                    br.pos = q + 2;
                    for (auto& c : comps) c.pred = 0;
                    rst_left = restart;
                }
                for (auto& c : comps)
                    for (int by = 0; by < c.v; by++)
                        for (int bx = 0; bx < c.h; bx++) {
                            int coef[64] = { 0 };
                            const int t = decode_sym(br, hdc[c.td]);
                            if (t > 11) throw std::runtime_error("bad DC size");
                            c.pred += t ? extend(br.bits(t), t) : 0;
                            coef[0] = c.pred * qt[c.tq][0];
                            for (int k = 1; k < 64;) {
                                const int rs = decode_sym(br, hac[c.ta]), r = rs >> 4, sz = rs & 15;
                                if (sz == 0) { if (r == 15) { k += 16; continue; } break; }
                                k += r;
                                if (k > 63) throw std::runtime_error("bad AC run");
                                coef[kZigzag[k]] = extend(br.bits(sz), sz) * qt[c.tq][kZigzag[k]];
                                k++;
                            }
                            idct8x8(coef, c.plane.data() + (size_t)((y * c.v + by) * 8) * c.bw + (x * c.h + bx) * 8, c.bw);
                        }
                if (restart) rst_left--;
            }
        /* ---- output */
        img.w = (unsigned)W; img.h = (unsigned)H;
        if (comps.size() == 1) {
            img.d = 1; img.px.resize((size_t)W * H);
            for (int y = 0; y < H; y++) memcpy(&img.px[(size_t)y * W], &comps[0].plane[(size_t)y * comps[0].bw], W);
            return true;
        }
        if (comps[0].h != hmax || comps[0].v != vmax) { err = "JPEG: subsampled luma is not supported"; return false; }
        img.d = 3; img.px.resize((size_t)W * H * 3);
        /* chroma at full resolution, libjpeg's fancy upsampling (jdsample.c: h2v1 / h2v2 triangle filters; v2 alone: replicate) */
        std::vector<unsigned char> full[2];
        for (int ci = 1; ci <= 2; ci++) {
            const Comp& c = comps[ci];
            const int cw = (W * c.h + hmax - 1) / hmax, ch = (H * c.v + vmax - 1) / vmax;     /* real (not block-padded) chroma size */
            std::vector<unsigned char>& o = full[ci - 1];
            o.resize((size_t)W * H);
            const bool uh = c.h < hmax, uv = c.v < vmax;
            std::vector<int> colsum(cw), line(2 * cw + 2);
            for (int y = 0; y < H; y++) {
                const int sy = uv ? y / 2 : y;
                const unsigned char* r0 = &c.plane[(size_t)std::min(sy, ch - 1) * c.bw];
                if (uh && uv) {
                    /* nearer neighbouring row: above for even output rows, below for odd ones */
                    const int ny = (y & 1) ? std::min(sy + 1, ch - 1) : std::max(sy - 1, 0);
                    const unsigned char* r1 = &c.plane[(size_t)ny * c.bw];
                    for (int i = 0; i < cw; i++) colsum[i] = 3 * r0[i] + r1[i];
                    up_h2(colsum.data(), cw, line.data(), 8, 7, 4);
                    for (int xx = 0; xx < W; xx++) o[(size_t)y * W + xx] = (unsigned char)line[xx];
                } else if (uh) {
                    for (int i = 0; i < cw; i++) colsum[i] = r0[i];
                    up_h2(colsum.data(), cw, line.data(), 1, 2, 2);
                    for (int xx = 0; xx < W; xx++) o[(size_t)y * W + xx] = (unsigned char)line[xx];
                } else {
                    for (int xx = 0; xx < W; xx++) o[(size_t)y * W + xx] = r0[xx];
                }
            }
        }
        /* libjpeg's YCbCr -> RGB (jdcolor.c: 16-bit fixed point, ONE_HALF rounding) */
        auto fix = [](double v) { return (int)(v * 65536.0 + 0.5); };
        const int c_r = fix(1.40200), c_b = fix(1.77200), c_g1 = fix(0.34414), c_g2 = fix(0.71414), half = 1 << 15;
        auto clamp = [](int v) { return (unsigned char)(v < 0 ? 0 : v > 255 ? 255 : v); };
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                const int Y = comps[0].plane[(size_t)y * comps[0].bw + x], cb = full[0][(size_t)y * W + x] - 128, cr = full[1][(size_t)y * W + x] - 128;
                unsigned char* o = &img.px[((size_t)y * W + x) * 3];
                o[0] = clamp(Y + ((c_r * cr + half) >> 16));
                o[1] = clamp(Y + ((-c_g1 * cb - c_g2 * cr + half) >> 16));
                o[2] = clamp(Y + ((c_b * cb + half) >> 16));
            }
        return true;
    } catch (const std::runtime_error& e) {
        err = std::string("JPEG: corrupt entropy-coded data (") + e.what() + ")";
        return false;
    }
}

} // namespace imgio
