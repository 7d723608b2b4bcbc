/* imgio.h -- minimal image file I/O for the command-line harness (srcnn_cli.cpp).
 *
 * The reference CLI (src/test.cpp:134-178, 180-274, 545-590) loads JPEG / PNG / BMP through FLTK + fl_imgtk and writes
 * PNG through libpng; none of those libraries is part of the library proper.  Here: PNG (8/16-bit, all colour types,
 * non-interlaced) decoded and encoded directly on zlib, uncompressed 24/32-bit BMP, and binary PNM (P5/P6).
 * Baseline JPEG (sequential Huffman, 8-bit, gray / YCbCr, the reference's Pictures/castle.jpg) is decoded by jpegdec.cpp;
 * progressive files are refused with a message.  Pure host code, no CUDA.
 */
#ifndef SRCNN_IMGIO_H
#define SRCNN_IMGIO_H

#include <string>
#include <vector>

namespace imgio {

struct Image {
    unsigned w = 0, h = 0, d = 0;            /* d = bytes per pixel: 1 gray, 3 RGB, 4 RGBA (tightly packed, row-major) */
    std::vector<unsigned char> px;
};

enum Type { T_UNKNOWN = -1, T_JPEG = 1, T_PNG = 2, T_BMP = 3, T_PNM = 4 };

/* sniff the container from the first bytes (same numbering as the reference's testImageFile, PNM added) */
Type sniff(const unsigned char* data, size_t n);

bool read_file(const std::string& path, std::vector<unsigned char>& out);
/* decode; on failure returns false and fills `err` */
bool decode(const unsigned char* data, size_t n, Image& img, std::string& err);
bool load(const std::string& path, Image& img, Type* type, std::string& err);
bool decode_jpeg(const unsigned char* data, size_t n, Image& img, std::string& err);     /* jpegdec.cpp */

/* 8-bit PNG, colour type from d (1 gray, 3 RGB, 4 RGBA), no interlace -- what the reference's savetopng writes */
bool encode_png(const unsigned char* px, unsigned w, unsigned h, unsigned d, std::vector<unsigned char>& out);
bool save_png(const std::string& path, const unsigned char* px, unsigned w, unsigned h, unsigned d);
bool save_pnm(const std::string& path, const unsigned char* px, unsigned w, unsigned h, unsigned d);

/* what the reference's convImage does before calling ProcessSRCNN (src/test.cpp:44-119): gray -> RGB; 3/4 kept */
void to_rgb_or_rgba(Image& img);

} // namespace imgio
#endif
