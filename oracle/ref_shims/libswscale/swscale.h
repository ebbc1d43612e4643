/* Shim of libswscale/swscale.h (colour conversion of the decoded frame; never reached by the tests). */
#ifndef LSFA_SHIM_SWSCALE_H
#define LSFA_SHIM_SWSCALE_H
#include "../libavcodec/avcodec.h"
struct SwsContext;
#define SWS_BICUBIC 4
static inline struct SwsContext* sws_getCachedContext(struct SwsContext* c, int sw, int sh, int sf, int dw, int dh, int df, int flags,
                                                      void* a, void* b, const double* p) {
  (void)c; (void)sw; (void)sh; (void)sf; (void)dw; (void)dh; (void)df; (void)flags; (void)a; (void)b; (void)p; LSFA_NO_FFMPEG(); return NULL;
}
static inline int sws_scale(struct SwsContext* c, uint8_t* const s[], const int ss[], int y, int h, uint8_t* const d[], const int ds[]) {
  (void)c; (void)s; (void)ss; (void)y; (void)h; (void)d; (void)ds; LSFA_NO_FFMPEG(); return -1;
}
static inline void sws_freeContext(struct SwsContext* c) { (void)c; LSFA_NO_FFMPEG(); }
#endif
