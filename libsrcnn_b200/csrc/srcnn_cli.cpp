/* srcnn -- command-line harness on the drop-in API (include/libsrcnn.h), the counterpart of the reference's only
 * caller, src/test.cpp.  Same command line (src/test.cpp:290-448, 513-531):
 *
 *     srcnn [source image file] [options] (output image file)
 *       --scale=(ratio)   output size factor, default 2.0
 *       --step            scale by factor-2 steps (ProcessSRCNN's step-scale chain)
 *       --filter=(0..4)   0 nearest, 1 bilinear, 2 bicubic (default), 3 Lanczos-3, 4 B-spline
 *       --waitakey        wait for ENTER before exiting
 *     additions: --noconv (do not ask for / save the convolution plane), --decode-only (load, normalise, save; no GPU)
 *
 * Default output names follow the reference: <source>_resized.png and <source>_convolution.png.  The program calls
 * nothing but ConfigureFilterSRCNN / ProcessSRCNN and frees the results with delete[] (src/test.cpp:647-719), so it
 * runs unchanged against either library.  FLTK / fl_imgtk / libpng are replaced by imgio.cpp (PNG on zlib, BMP, PNM).
 */
#include "libsrcnn.h"
#include "imgio.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

namespace {

struct Options {
    float       multiply = 2.0f;
    bool        step = false, waitakey = false, noconv = false, decode_only = false;
    std::string me, src, dst, conv;
    SRCNNFilterType filter = SRCNNF_Bicubic;
};

std::string stem_with(const std::string& src, const char* suffix)
{
    const size_t dot = src.find_last_of('.');
    const size_t slash = src.find_last_of("/\\");
    const bool has_ext = dot != std::string::npos && (slash == std::string::npos || dot > slash);
    return (has_ext ? src.substr(0, dot) : src) + suffix + ".png";
}

bool parse(int argc, char** argv, Options& o)
{
    o.me = argv[0];
    const size_t sl = o.me.find_last_of("/\\");
    if (sl != std::string::npos) o.me = o.me.substr(sl + 1);
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        if (a.rfind("--scale=", 0) == 0) {
            const float v = (float)atof(a.c_str() + 8);
            if (v > 0.f) o.multiply = v;                         /* non-positive values keep the default, as the reference does */
        } else if (a.rfind("--step", 0) == 0) o.step = true;
        else if (a.rfind("--filter=", 0) == 0) {
            if (a.size() > 9) {
                const int f = atoi(a.c_str() + 9);
                o.filter = (f >= 0 && f <= 4) ? (SRCNNFilterType)f : SRCNNF_Bicubic;
            }
        } else if (a.rfind("--waitakey", 0) == 0) o.waitakey = true;
        else if (a == "--noconv") o.noconv = true;
        else if (a == "--decode-only") o.decode_only = true;
        else if (o.src.empty()) o.src = a;
        else if (o.dst.empty()) o.dst = a;
    }
    if (o.src.empty()) return false;
    if (o.dst.empty()) o.dst = stem_with(o.src, "_resized");
    o.conv = stem_with(o.src, "_convolution");
    return true;
}

void usage(const Options& o)
{
    printf("%s : libsrcnn testing program (B200 library, no FLTK), ver %s\n\n", o.me.c_str(), "0.1.10.40-b200");
    printf("  usage:\n      %s [source image file] [options] (output image file)\n\n", o.me.c_str());
    printf("  options:\n");
    printf("      --scale=(ratio : 0.0<999...) : adjust size of output image.\n");
    printf("      --step                       : scaling by fator 2 steps.\n");
    printf("      --waitakey                   : wait for ENTER for end of job.\n");
    printf("      --filter=(0...4)             : Changes interpolation filter as ...\n");
    printf("                   0 = Nearest filter\n                   1 = Bilinear filter\n");
    printf("                   2 = Bicubic filter (default)\n                   3 = Lanzcos-3 filter\n");
    printf("                   4 = B-Spline filter\n");
    printf("      --noconv                     : skip the convolution-output image.\n");
    printf("      --decode-only                : load, convert to RGB(A), save; no scaling.\n\n");
}

const char* filter_name(SRCNNFilterType f)
{
    switch (f) {
        case SRCNNF_Nearest: return "Nearest"; case SRCNNF_Bilinear: return "Bilinear"; case SRCNNF_Bicubic: return "Bicubic";
        case SRCNNF_Lanczos3: return "Lanczos3"; case SRCNNF_Bspline: return "B-Spline";
    }
    return "?";
}

bool save_any(const std::string& path, const unsigned char* px, unsigned w, unsigned h, unsigned d)
{
    const size_t dot = path.find_last_of('.');
    const std::string ext = dot == std::string::npos ? "" : path.substr(dot);
    if ((ext == ".ppm" && d == 3) || (ext == ".pgm" && d == 1)) return imgio::save_pnm(path, px, w, h, d);
    return imgio::save_png(path, px, w, h, d);
}

} // namespace

int main(int argc, char** argv)
{
    Options o;
    if (!parse(argc, argv, o)) { usage(o); return -1; }
    if (o.multiply <= 0.f) { printf("- Error: scaling value under zero.\n"); return -2; }

    printf("- Loading image : %s", o.src.c_str());
    imgio::Image img; imgio::Type type = imgio::T_UNKNOWN; std::string err;
    const bool ok = imgio::load(o.src, img, &type, err);
    printf("\n");
    if (!ok) { printf("- Failed to load image. (%s)\n", err.c_str()); return 1; }
    printf("- Image loaded type : %s | %u x %u x %u bytes\n",
           type == imgio::T_JPEG ? "JPEG" : type == imgio::T_PNG ? "PNG" : type == imgio::T_BMP ? "BMP" : type == imgio::T_PNM ? "PNM" : "?", img.w, img.h, img.d);
    imgio::to_rgb_or_rgba(img);                                    /* gray -> RGB, as src/test.cpp:44-119 */
    if (img.d < 3) { printf("- Error: Unsupported image.\n"); return 1; }
    if (o.decode_only) {
        const bool s = save_any(o.dst, img.px.data(), img.w, img.h, img.d);
        printf("- Saving decoded image to %s ... %s\n", o.dst.c_str(), s ? "Ok." : "Failure.");
        return s ? 0 : 1;
    }

    printf("- Scaling ratio : %.2f\n", o.multiply);
    if (o.step) printf("- Step scaling enabled, warning: takes a lot of times.\n");
    printf("- Filter : %s\n", filter_name(o.filter));
    ConfigureFilterSRCNN(o.filter, o.step);
    printf("- Processing SRCNN ... ");
    fflush(stdout);

    unsigned char *out = nullptr, *conv = nullptr;
    unsigned outsz = 0, convsz = 0;
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = ProcessSRCNN(img.px.data(), img.w, img.h, img.d, o.multiply, out, outsz,
                                o.noconv ? nullptr : &conv, o.noconv ? nullptr : &convsz);
    const auto t1 = std::chrono::steady_clock::now();
    int status = 0;
    if (rc == 0 && outsz > 0) {
        /* the reference recomputes the size as unsigned(w * multiply) (src/test.cpp:667-668) */
        const unsigned nw = (unsigned)((float)img.w * o.multiply);
        unsigned nh = (unsigned)((float)img.h * o.multiply);
        if ((size_t)nw * nh * img.d != outsz && nw && outsz % (nw * img.d) == 0) nh = outsz / (nw * img.d);   /* step chains may round the height differently */
        printf("Test Ok, took %u ms.\n", (unsigned)std::chrono::duration_cast<std::chrono::milliseconds>(t1 - t0).count());
        printf("- Saving resized result to %s ... ", o.dst.c_str());
        const bool s = (size_t)nw * nh * img.d == outsz && save_any(o.dst, out, nw, nh, img.d);
        printf(s ? "Ok.\n" : "Failure.\n");
        if (!s) status = 1;
        if (convsz > 0 && conv) {
            printf("- Saving convolution result to %s ... ", o.conv.c_str());
            const bool c = (size_t)nw * nh == convsz && save_any(o.conv, conv, nw, nh, 1);
            printf(c ? "Ok.\n" : "Failure.\n");
            if (!c) status = 1;
        }
    } else {
        printf("Failed, error code = %d\n", rc);
        status = 1;
    }
    delete[] out;
    delete[] conv;
    fflush(stdout);
    if (o.waitakey) {
        printf("- Input any number and press ENTER to terminate, check memory state.\n");
        fflush(stdout);
        unsigned meaningless = 0;
        if (scanf("%u", &meaningless) != 1) meaningless = 0;
    }
    return status;
}
