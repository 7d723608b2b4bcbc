/* imgio.cpp -- see imgio.h.  PNG per the PNG specification (chunks, zlib stream, the five scanline filters). */
#include "imgio.h"

#include <zlib.h>

#include <algorithm>
#include <exception>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace imgio {

namespace {

const unsigned char kPngSig[8] = { 0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A };

inline unsigned be32(const unsigned char* p) { return ((unsigned)p[0] << 24) | ((unsigned)p[1] << 16) | ((unsigned)p[2] << 8) | p[3]; }
inline unsigned le32(const unsigned char* p) { return ((unsigned)p[3] << 24) | ((unsigned)p[2] << 16) | ((unsigned)p[1] << 8) | p[0]; }
inline unsigned le16(const unsigned char* p) { return ((unsigned)p[1] << 8) | p[0]; }
inline void put_be32(std::vector<unsigned char>& v, unsigned x)
{
    v.push_back((unsigned char)(x >> 24)); v.push_back((unsigned char)(x >> 16));
    v.push_back((unsigned char)(x >> 8));  v.push_back((unsigned char)x);
}

inline int paeth(int a, int b, int c)
{
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

bool decode_png(const unsigned char* data, size_t n, Image& img, std::string& err)
{
    size_t pos = 8;
    unsigned w = 0, h = 0, depth = 0, ctype = 0, interlace = 0;
    bool have_ihdr = false, have_end = false;
    std::vector<unsigned char> idat, plte, trns;
    while (pos + 12 <= n && !have_end) {
        const unsigned len = be32(data + pos);
        const unsigned char* tag = data + pos + 4;
        const unsigned char* body = data + pos + 8;
        if ((size_t)len > n - pos - 12) { err = "PNG: truncated chunk"; return false; }
        if ((unsigned)crc32(0L, tag, (uInt)len + 4) != be32(body + len)) { err = "PNG: chunk CRC mismatch (corrupt file)"; return false; }
        if (!memcmp(tag, "IHDR", 4) && len >= 13) {
            w = be32(body); h = be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12];
            have_ihdr = true;
        } else if (!memcmp(tag, "PLTE", 4)) plte.assign(body, body + len);
        else if (!memcmp(tag, "tRNS", 4)) trns.assign(body, body + len);
        else if (!memcmp(tag, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
        else if (!memcmp(tag, "IEND", 4)) have_end = true;
        pos += 12 + (size_t)len;
    }
    if (!have_ihdr || !w || !h || idat.empty()) { err = "PNG: missing IHDR / IDAT"; return false; }
    if (interlace) { err = "PNG: Adam7 interlace is not supported"; return false; }
    unsigned ch;
    switch (ctype) {
        case 0: ch = 1; break; case 2: ch = 3; break; case 3: ch = 1; break; case 4: ch = 2; break; case 6: ch = 4; break;
        default: err = "PNG: bad colour type"; return false;
    }
    if (!(depth == 8 || depth == 16 || ((ctype == 0 || ctype == 3) && (depth == 1 || depth == 2 || depth == 4)))) {
        err = "PNG: unsupported bit depth"; return false;
    }
    if (ctype == 3 && (depth == 16 || plte.size() < 3)) { err = "PNG: bad palette image"; return false; }
    const size_t bpp = (ch * depth + 7) / 8;                     /* filter distance in bytes */
    const size_t stride = ((size_t)w * ch * depth + 7) / 8;
    if (stride == 0 || (size_t)h > ((size_t)1 << 34) / (stride + 1)) { err = "PNG: image too large"; return false; }   /* 16 GiB of raw scanlines */
    std::vector<unsigned char> raw((stride + 1) * (size_t)h);
    {
        z_stream zs; memset(&zs, 0, sizeof zs);
        if (inflateInit(&zs) != Z_OK) { err = "zlib: inflateInit failed"; return false; }
        /* avail_in / avail_out are 32-bit: hand the buffers over in bounded pieces */
        size_t in_left = idat.size(), out_left = raw.size(); int zr;
        zs.next_in = idat.data(); zs.next_out = raw.data();
        do {
            if (zs.avail_in == 0 && in_left)   { const size_t c = std::min<size_t>(in_left, (size_t)1 << 30);  zs.avail_in = (uInt)c;  in_left -= c; }
            if (zs.avail_out == 0 && out_left) { const size_t c = std::min<size_t>(out_left, (size_t)1 << 30); zs.avail_out = (uInt)c; out_left -= c; }
            zr = inflate(&zs, Z_NO_FLUSH);
        } while (zr == Z_OK && (zs.avail_out || out_left));
        const size_t out_off = raw.size() - out_left - zs.avail_out;
        inflateEnd(&zs);
        if (out_off != raw.size()) { err = "PNG: corrupt or short IDAT stream"; return false; }
    }
    /* undo the scanline filters in place */
    for (unsigned y = 0; y < h; y++) {
        unsigned char* cur = raw.data() + (size_t)y * (stride + 1);
        const unsigned char* up = y ? cur - stride : nullptr;       /* previous row's pixel bytes start at cur-(stride+1)+1 */
        const unsigned ft = cur[0];
        unsigned char* s = cur + 1;
        for (size_t i = 0; i < stride; i++) {
            const int a = i >= bpp ? s[i - bpp] : 0;
            const int b = up ? up[i] : 0;
            const int c = (up && i >= bpp) ? up[i - bpp] : 0;
            int v = s[i];
            switch (ft) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: v += paeth(a, b, c); break;
                default: err = "PNG: bad filter type"; return false;
            }
            s[i] = (unsigned char)v;
        }
    }
    /* expand to 8-bit gray / RGB / RGBA */
    const bool pal_alpha = ctype == 3 && !trns.empty();
    unsigned od = (ctype == 0) ? 1 : (ctype == 2) ? 3 : (ctype == 3) ? (pal_alpha ? 4 : 3) : 4;   /* gray+alpha -> RGBA */
    img.w = w; img.h = h; img.d = od;
    img.px.assign((size_t)w * h * od, 0);
    for (unsigned y = 0; y < h; y++) {
        const unsigned char* s = raw.data() + (size_t)y * (stride + 1) + 1;
        unsigned char* o = img.px.data() + (size_t)y * w * od;
        for (unsigned x = 0; x < w; x++) {
            unsigned smp[4] = { 0, 0, 0, 0 };
            for (unsigned k = 0; k < ch; k++) {
                const size_t i = (size_t)x * ch + k;
                if (depth == 8) smp[k] = s[i];
                else if (depth == 16) smp[k] = s[2 * i];            /* keep the high byte */
                else {
                    const unsigned per = 8 / depth, sh = (per - 1 - (unsigned)(i % per)) * depth;
                    smp[k] = (s[i / per] >> sh) & ((1u << depth) - 1);
                    if (ctype == 0) smp[k] = smp[k] * 255u / ((1u << depth) - 1);
                }
            }
            switch (ctype) {
                case 0: o[x] = (unsigned char)smp[0]; break;
                case 2: o[3 * x] = (unsigned char)smp[0]; o[3 * x + 1] = (unsigned char)smp[1]; o[3 * x + 2] = (unsigned char)smp[2]; break;
                case 3: {
                    const unsigned idx = smp[0];
                    unsigned char* q = o + (size_t)x * od;
                    if (3 * idx + 2 < plte.size()) { q[0] = plte[3 * idx]; q[1] = plte[3 * idx + 1]; q[2] = plte[3 * idx + 2]; }
                    if (pal_alpha) q[3] = idx < trns.size() ? trns[idx] : 255;
                    break;
                }
                case 4: { unsigned char* q = o + 4 * (size_t)x; q[0] = q[1] = q[2] = (unsigned char)smp[0]; q[3] = (unsigned char)smp[1]; break; }
                case 6: { unsigned char* q = o + 4 * (size_t)x; for (int k = 0; k < 4; k++) q[k] = (unsigned char)smp[k]; break; }
            }
        }
    }
    return true;
}

bool decode_bmp(const unsigned char* data, size_t n, Image& img, std::string& err)
{
    if (n < 54) { err = "BMP: too short"; return false; }
    const unsigned off = le32(data + 10), hdr = le32(data + 14);
    const int w = (int)le32(data + 18); int h = (int)le32(data + 22);
    const unsigned bpp = le16(data + 28), comp = le32(data + 30);
    if (hdr < 40 || w <= 0 || h == 0) { err = "BMP: unsupported header"; return false; }
    if (!(bpp == 24 || bpp == 32) || !(comp == 0 || (comp == 3 && bpp == 32))) { err = "BMP: only uncompressed 24/32-bit"; return false; }
    const bool flip = h > 0; if (h < 0) h = -h;
    const size_t stride = (((size_t)w * bpp + 31) / 32) * 4;
    if ((size_t)off + stride * (size_t)h > n) { err = "BMP: truncated"; return false; }
    const unsigned d = bpp / 8;
    img.w = (unsigned)w; img.h = (unsigned)h; img.d = d;
    img.px.resize((size_t)w * h * d);
    for (int y = 0; y < h; y++) {
        const unsigned char* s = data + off + stride * (size_t)(flip ? h - 1 - y : y);
        unsigned char* o = img.px.data() + (size_t)y * w * d;
        for (int x = 0; x < w; x++) {
            o[d * x] = s[d * x + 2]; o[d * x + 1] = s[d * x + 1]; o[d * x + 2] = s[d * x];        /* BGR(A) -> RGB(A) */
            if (d == 4) o[4 * x + 3] = s[4 * x + 3];
        }
    }
    return true;
}

bool decode_pnm(const unsigned char* data, size_t n, Image& img, std::string& err)
{
    size_t pos = 2; unsigned v[3] = { 0, 0, 0 };
    for (int k = 0; k < 3; k++) {
        for (;;) {
            while (pos < n && (data[pos] == ' ' || data[pos] == '\n' || data[pos] == '\r' || data[pos] == '\t')) pos++;
            if (pos < n && data[pos] == '#') { while (pos < n && data[pos] != '\n') pos++; continue; }
            break;
        }
        if (pos >= n || data[pos] < '0' || data[pos] > '9') { err = "PNM: bad header"; return false; }
        while (pos < n && data[pos] >= '0' && data[pos] <= '9') v[k] = v[k] * 10 + (data[pos++] - '0');
    }
    pos++;                                                       /* single whitespace after maxval */
    const unsigned d = data[1] == '6' ? 3 : 1;
    if (!v[0] || !v[1] || v[2] != 255) { err = "PNM: only 8-bit binary P5/P6"; return false; }
    const size_t need = (size_t)v[0] * v[1] * d;
    if (pos + need > n) { err = "PNM: truncated"; return false; }
    img.w = v[0]; img.h = v[1]; img.d = d;
    img.px.assign(data + pos, data + pos + need);
    return true;
}

void put_chunk(std::vector<unsigned char>& out, const char* tag, const unsigned char* body, size_t len)
{
    put_be32(out, (unsigned)len);
    const size_t start = out.size();
    out.insert(out.end(), tag, tag + 4);
    if (len) out.insert(out.end(), body, body + len);
    put_be32(out, (unsigned)crc32(0L, out.data() + start, (uInt)(len + 4)));
}

} // namespace

Type sniff(const unsigned char* p, size_t n)
{
    if (n >= 3 && p[0] == 0xFF && p[1] == 0xD8 && p[2] == 0xFF) return T_JPEG;
    if (n >= 8 && !memcmp(p, kPngSig, 8)) return T_PNG;
    if (n >= 2 && p[0] == 'B' && p[1] == 'M') return T_BMP;
    if (n >= 2 && p[0] == 'P' && (p[1] == '5' || p[1] == '6')) return T_PNM;
    return T_UNKNOWN;
}

bool read_file(const std::string& path, std::vector<unsigned char>& out)
{
    FILE* fp = fopen(path.c_str(), "rb");
    if (!fp) return false;
    fseek(fp, 0, SEEK_END); const long sz = ftell(fp); fseek(fp, 0, SEEK_SET);
    if (sz <= 0) { fclose(fp); return false; }
    out.resize((size_t)sz);
    const size_t got = fread(out.data(), 1, out.size(), fp);
    fclose(fp);
    return got == out.size();
}

bool decode(const unsigned char* data, size_t n, Image& img, std::string& err)
{
    try {
        switch (sniff(data, n)) {
            case T_PNG: return decode_png(data, n, img, err);
            case T_BMP: return decode_bmp(data, n, img, err);
            case T_PNM: return decode_pnm(data, n, img, err);
            case T_JPEG: return decode_jpeg(data, n, img, err);
            default: err = "unknown image format"; return false;
        }
    } catch (const std::exception& e) {          /* std::bad_alloc / length_error on absurd dimensions */
        err = std::string("out of memory while decoding (") + e.what() + ")";
        return false;
    }
}

bool load(const std::string& path, Image& img, Type* type, std::string& err)
{
    std::vector<unsigned char> buf;
    if (!read_file(path, buf)) { err = "cannot read file"; if (type) *type = T_UNKNOWN; return false; }
    if (type) *type = sniff(buf.data(), buf.size());
    return decode(buf.data(), buf.size(), img, err);
}

bool encode_png(const unsigned char* px, unsigned w, unsigned h, unsigned d, std::vector<unsigned char>& out)
{
    if (!px || !w || !h || !(d == 1 || d == 3 || d == 4)) return false;
    const size_t stride = (size_t)w * d;
    /* filter type 1 (Sub) on every row: cheap and compresses photographic data far better than None */
    std::vector<unsigned char> raw((stride + 1) * (size_t)h);
    for (unsigned y = 0; y < h; y++) {
        const unsigned char* s = px + stride * y;
        unsigned char* o = raw.data() + (stride + 1) * (size_t)y;
        o[0] = 1;
        for (size_t i = 0; i < stride; i++) o[1 + i] = (unsigned char)(s[i] - (i >= d ? s[i - d] : 0));
    }
    std::vector<unsigned char> comp;
    {
        z_stream zs; memset(&zs, 0, sizeof zs);
        if (deflateInit(&zs, 3) != Z_OK) return false;
        comp.resize(raw.size() / 2 + 65536);
        size_t in_left = raw.size(), out_off = 0; int zr;
        zs.next_in = raw.data();
        do {
            if (zs.avail_in == 0 && in_left) { const size_t c = std::min<size_t>(in_left, (size_t)1 << 30); zs.avail_in = (uInt)c; in_left -= c; }
            if (comp.size() - out_off < 65536) comp.resize(comp.size() * 2);
            const size_t out_now = std::min<size_t>(comp.size() - out_off, (size_t)1 << 30);
            zs.next_out = comp.data() + out_off; zs.avail_out = (uInt)out_now;
            zr = deflate(&zs, in_left == 0 ? Z_FINISH : Z_NO_FLUSH);
            out_off += out_now - zs.avail_out;
            if (zr != Z_OK && zr != Z_STREAM_END && zr != Z_BUF_ERROR) { deflateEnd(&zs); return false; }
        } while (zr != Z_STREAM_END);
        deflateEnd(&zs);
        comp.resize(out_off);
    }
    out.clear();
    out.insert(out.end(), kPngSig, kPngSig + 8);
    unsigned char ihdr[13];
    ihdr[0] = (unsigned char)(w >> 24); ihdr[1] = (unsigned char)(w >> 16); ihdr[2] = (unsigned char)(w >> 8); ihdr[3] = (unsigned char)w;
    ihdr[4] = (unsigned char)(h >> 24); ihdr[5] = (unsigned char)(h >> 16); ihdr[6] = (unsigned char)(h >> 8); ihdr[7] = (unsigned char)h;
    ihdr[8] = 8; ihdr[9] = d == 1 ? 0 : (d == 3 ? 2 : 6); ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;
    put_chunk(out, "IHDR", ihdr, 13);
    for (size_t off = 0; off < comp.size(); off += (size_t)1 << 20)          /* 1 MiB IDAT chunks */
        put_chunk(out, "IDAT", comp.data() + off, std::min<size_t>((size_t)1 << 20, comp.size() - off));
    put_chunk(out, "IEND", nullptr, 0);
    return true;
}

static bool write_all(const std::string& path, const unsigned char* p, size_t n, const char* head = nullptr)
{
    FILE* fp = fopen(path.c_str(), "wb");
    if (!fp) return false;
    bool ok = true;
    if (head) ok = fputs(head, fp) >= 0;
    ok = ok && fwrite(p, 1, n, fp) == n;
    return (fclose(fp) == 0) && ok;
}

bool save_png(const std::string& path, const unsigned char* px, unsigned w, unsigned h, unsigned d)
{
    std::vector<unsigned char> out;
    return encode_png(px, w, h, d, out) && write_all(path, out.data(), out.size());
}

bool save_pnm(const std::string& path, const unsigned char* px, unsigned w, unsigned h, unsigned d)
{
    if (!(d == 1 || d == 3)) return false;
    char head[64];
    snprintf(head, sizeof head, "P%c\n%u %u\n255\n", d == 3 ? '6' : '5', w, h);
    return write_all(path, px, (size_t)w * h * d, head);
}

void to_rgb_or_rgba(Image& img)
{
    if (img.d != 1) return;
    std::vector<unsigned char> rgb((size_t)img.w * img.h * 3);
    for (size_t i = 0, n = (size_t)img.w * img.h; i < n; i++) rgb[3 * i] = rgb[3 * i + 1] = rgb[3 * i + 2] = img.px[i];
    img.px.swap(rgb); img.d = 3;
}

} // namespace imgio
