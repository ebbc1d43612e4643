/* Shim of libavutil/pixfmt.h: the two pixel formats the translation unit names. */
#ifndef LSFA_SHIM_AVUTIL_PIXFMT_H
#define LSFA_SHIM_AVUTIL_PIXFMT_H
enum AVPixelFormat { AV_PIX_FMT_YUV420P = 0, AV_PIX_FMT_BGR24 = 3 };
#endif
