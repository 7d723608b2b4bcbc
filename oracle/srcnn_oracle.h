/* TEST INFRASTRUCTURE ONLY -- plain-C CPU restatement of the reference's ProcessSRCNN path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this.  The product (libsrcnn_b200/lib/libsrcnn.so) never links or calls it.
 * Parity status: PINNED -- checked against the reference's golden vectors
 * (Pictures/butterfly_srcnn*.png, committed under tests/golden/) and, stage by stage, against the
 * unmodified reference compiled into oracle/_ref/libsrcnn_ref.so (tests/test_oracle.py). */
#ifndef SRCNN_ORACLE_H
#define SRCNN_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_NEAREST = 0, ORC_BILINEAR = 1, ORC_BICUBIC = 2, ORC_LANCZOS3 = 3, ORC_BSPLINE = 4 };

int  orc_threads(void);
void orc_output_dims(unsigned w, unsigned h, float multiply, unsigned* ow, unsigned* oh);

/* stages (all planes tightly packed row-major float) */
void orc_color_in(const unsigned char* src, unsigned w, unsigned h, unsigned d,
                  float* y, float* cb, float* cr, float* a);
int  orc_contrib_table(int filter, unsigned dst, unsigned src, int* left, int* right, double* weights, int maxwin);
int  orc_window(int filter, unsigned dst, unsigned src);
int  orc_resize(const float* src, unsigned sw, unsigned sh, unsigned dw, unsigned dh, int filter, float* dst);
void orc_conv1(const float* y, unsigned w, unsigned h, float* out64);
void orc_conv2(const float* l1, unsigned w, unsigned h, float* out32);
void orc_conv3(const float* l2, unsigned w, unsigned h, float* out);
void orc_color_out(const float* const* planes, unsigned w, unsigned h, unsigned d, unsigned char* out);

/* one pass (reference doSRCNN) and the public entry (reference ProcessSRCNN incl. step-scale loop).
 * out/conv are malloc'd here; free with orc_free.  y_plane (optional, ow*oh floats, caller-allocated)
 * receives the interpolated Y plane of the LAST pass; yprime likewise the conv III plane. */
int  orc_pass(const unsigned char* src, unsigned w, unsigned h, unsigned d, float multiply, int filter,
              unsigned char** out, unsigned* outsz, unsigned char** conv, unsigned* convsz,
              float* y_plane, float* yprime);
int  orc_process(const unsigned char* src, unsigned w, unsigned h, unsigned d, float multiply,
                 int filter, int stepscale,
                 unsigned char** out, unsigned* outsz, unsigned char** conv, unsigned* convsz);
void orc_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
