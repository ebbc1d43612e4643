/* Shim of libavcodec/avcodec.h + libavutil/frame.h: only what coviar_data_loader.c names.
 * AVFrameSideData is the public struct the function under test reads (type, data, size);
 * everything else is an opaque stand-in with an aborting stub - no decoder exists here. */
#ifndef LSFA_SHIM_AVCODEC_H
#define LSFA_SHIM_AVCODEC_H
#include <assert.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../libavutil/pixfmt.h"

#define LSFA_NO_FFMPEG() do { fprintf(stderr, "oracle/_ref: FFmpeg entry point called; no decoder in this build\n"); abort(); } while (0)

enum AVFrameSideDataType { AV_FRAME_DATA_MOTION_VECTORS = 8 };
typedef struct AVFrameSideData {
  enum AVFrameSideDataType type;
  uint8_t* data;
  int size;
  void* metadata;
  void* buf;
} AVFrameSideData;

typedef struct AVFrame { uint8_t* data[8]; int linesize[8]; int width, height; } AVFrame;
typedef struct AVPicture { uint8_t* data[8]; int linesize[8]; } AVPicture;
typedef struct AVPacket { uint8_t* data; int size; } AVPacket;
typedef struct AVCodec { int id; } AVCodec;
typedef struct AVCodecContext { int width, height; } AVCodecContext;
typedef struct AVCodecParserContext { int pict_type; } AVCodecParserContext;
typedef struct AVDictionary AVDictionary;
enum AVCodecID { AV_CODEC_ID_MPEG4 = 12, AV_CODEC_ID_H264 = 27 };
enum AVPictureType { AV_PICTURE_TYPE_I = 1 };
#define AV_NOPTS_VALUE ((int64_t)0x8000000000000000ULL)
#define AV_LOG_QUIET (-8)

static inline void avcodec_register_all(void) { LSFA_NO_FFMPEG(); }
static inline AVCodec* avcodec_find_decoder(int id) { (void)id; LSFA_NO_FFMPEG(); return NULL; }
static inline AVCodecContext* avcodec_alloc_context3(const AVCodec* c) { (void)c; LSFA_NO_FFMPEG(); return NULL; }
static inline AVCodecParserContext* av_parser_init(int id) { (void)id; LSFA_NO_FFMPEG(); return NULL; }
static inline int av_dict_set(AVDictionary** d, const char* k, const char* v, int f) { (void)d; (void)k; (void)v; (void)f; LSFA_NO_FFMPEG(); return -1; }
static inline int avcodec_open2(AVCodecContext* c, const AVCodec* d, AVDictionary** o) { (void)c; (void)d; (void)o; LSFA_NO_FFMPEG(); return -1; }
static inline AVFrame* av_frame_alloc(void) { LSFA_NO_FFMPEG(); return NULL; }
static inline void av_frame_free(AVFrame** f) { (void)f; LSFA_NO_FFMPEG(); }
static inline void av_init_packet(AVPacket* p) { (void)p; LSFA_NO_FFMPEG(); }
static inline int av_parser_parse2(AVCodecParserContext* s, AVCodecContext* c, uint8_t** ob, int* os, const uint8_t* b, int bs,
                                   int64_t pts, int64_t dts, int64_t pos) {
  (void)s; (void)c; (void)ob; (void)os; (void)b; (void)bs; (void)pts; (void)dts; (void)pos; LSFA_NO_FFMPEG(); return -1;
}
static inline int avcodec_decode_video2(AVCodecContext* c, AVFrame* f, int* got, const AVPacket* p) { (void)c; (void)f; (void)got; (void)p; LSFA_NO_FFMPEG(); return -1; }
static inline AVFrameSideData* av_frame_get_side_data(const AVFrame* f, int t) { (void)f; (void)t; LSFA_NO_FFMPEG(); return NULL; }
static inline void av_parser_close(AVCodecParserContext* s) { (void)s; LSFA_NO_FFMPEG(); }
static inline int avcodec_close(AVCodecContext* c) { (void)c; LSFA_NO_FFMPEG(); return -1; }
static inline void av_free(void* p) { (void)p; LSFA_NO_FFMPEG(); }
static inline void* av_malloc(size_t n) { (void)n; LSFA_NO_FFMPEG(); return NULL; }
static inline int avpicture_get_size(int fmt, int w, int h) { (void)fmt; (void)w; (void)h; LSFA_NO_FFMPEG(); return -1; }
static inline int avpicture_fill(AVPicture* p, const uint8_t* b, int fmt, int w, int h) { (void)p; (void)b; (void)fmt; (void)w; (void)h; LSFA_NO_FFMPEG(); return -1; }
static inline void av_log_set_level(int l) { (void)l; }
#endif
