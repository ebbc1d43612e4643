/* Shim of FFmpeg's public libavutil/motion_vector.h (test infrastructure, see README.md).
 * Field order and types follow the public FFmpeg ABI at the reference's pinned commit:
 * source, w, h, src_x, src_y, dst_x, dst_y, flags. */
#ifndef LSFA_SHIM_AVUTIL_MOTION_VECTOR_H
#define LSFA_SHIM_AVUTIL_MOTION_VECTOR_H
#include <stdint.h>
typedef struct AVMotionVector {
  int32_t source;      /* -1: past reference, +1: future reference */
  uint8_t w, h;        /* block size */
  int16_t src_x, src_y;
  int16_t dst_x, dst_y;
  uint64_t flags;
} AVMotionVector;
#endif
