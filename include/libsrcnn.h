/* libsrcnn.h -- public C++ API of the B200-native SRCNN library.
 *
 * Source- and link-compatible with rageworx/libsrcnn's public header (reference src/libsrcnn.h:34-54):
 * same enum, same two C++-linkage functions, hence the same mangled symbols
 *   _Z20ConfigureFilterSRCNN15SRCNNFilterTypeb
 *   _Z12ProcessSRCNNPKhjjjfRPhRjPS1_Pj
 * so an application built against the reference re-links (or LD_PRELOADs) against this library unchanged.
 * Everything behind these two calls runs as hand-written sm_100a CUDA; there is no CPU fallback: without a
 * usable B200 the call fails with a negative code (see below).
 */
#ifndef LIBSRCNN_B200_PUBLIC_H
#define LIBSRCNN_B200_PUBLIC_H

#if defined(LIBSRCNNSTATIC) || !defined(__GNUC__)
  #define SRCNN_API
#else
  #define SRCNN_API __attribute__((visibility("default")))
#endif
/* kept for sources that spell the reference's macro names */
#ifndef DLL_PUBLIC
  #define DLL_PUBLIC SRCNN_API
#endif

/* reference version word 0.1.10.40 (src/libsrcnn.h:34-35) -- the behaviour this build is a drop-in for */
#define LIBSRCNN_VERSION    0x00010A28

/* interpolation used for the Y plane before the network (src/libsrcnn.h:37-44); chroma/alpha planes use
 * bilinear, or nearest when SRCNNF_Nearest is selected (src/libsrcnn.cpp:703-713) */
typedef enum
{
    SRCNNF_Nearest = 0,
    SRCNNF_Bilinear,
    SRCNNF_Bicubic,     /* Mitchell-Netravali B = C = 1/3; the default */
    SRCNNF_Lanczos3,
    SRCNNF_Bspline
} SRCNNFilterType;

/* Process-wide settings read by every later ProcessSRCNN call (src/libsrcnn.cpp:930-941).
 * stepscale = "precision step": factors above 2 are applied as a chain of x2 passes plus a remainder pass. */
SRCNN_API void ConfigureFilterSRCNN( SRCNNFilterType ftype, bool stepscale = false );

/* Upscale a tightly packed interleaved 8-bit RGB (d == 3) or RGBA (d == 4) image by `multiply`
 * (src/libsrcnn.cpp:943-1064).
 *   outbuff    receives a new[]-allocated image of (unsigned)(w*multiply) x (unsigned)(h*multiply) x d bytes;
 *              outbuffsz its size.  The caller releases it with delete[].
 *   convbuff / convbuffsz  when BOTH are non-NULL additionally receive a new[]-allocated 1 byte/pixel plane:
 *              the network's luminance output truncated to 8 bit ("convolution output").
 * Returns 0 on success; -1 bad input (NULL, zero size, d not 3/4); -2 scaled size <= 0; -11/-12 output
 * allocation failure; -100 nothing produced (step mode with an effective factor of 1);
 * -20 no usable CUDA device / context creation failed; -21 a CUDA call or kernel failed (never a CPU fallback).
 */
SRCNN_API int ProcessSRCNN( const unsigned char* refbuff,
                            unsigned w, unsigned h, unsigned d,
                            float multiply,
                            unsigned char* &outbuff,
                            unsigned &outbuffsz,
                            unsigned char** convbuff,
                            unsigned* convbuffsz );

#endif /* LIBSRCNN_B200_PUBLIC_H */
