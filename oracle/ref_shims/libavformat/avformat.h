/* Shim of libavformat/avformat.h: coviar_data_loader.c includes it but uses nothing from it. */
#ifndef LSFA_SHIM_AVFORMAT_H
#define LSFA_SHIM_AVFORMAT_H
#include "../libavcodec/avcodec.h"
#endif
