/* tools/ffstub/ffstub.h -- TEST INFRASTRUCTURE.  The few libav* declarations the reference's include/ffapi.h and its
 * scan.c / motion.c / zoom.c touch, backed by raw files instead of FFmpeg (which is not in this image), so that the
 * UNMODIFIED tools compile (oracle/Makefile `reftools`) and run against shim/fftw3.h + libdspdct.  Together with
 * ffstub.c this replaces include/ffapi.c; include/ffapi.h itself is used as it lies in the reference checkout.
 *
 * Video file ("DSPV"): one text line  `DSPV1 <pix_fmt> <width> <height> <frames> <rate_num> <rate_den>\n`  (frames is a
 * fixed-width field, patched when an output is closed), then the frames, each the format's planes in plane order,
 * rows tightly packed.  Formats: gray, yuv420p, yuv444p, gbrp (8-bit), grayf32le, gbrpf32le (float32).
 * No codecs, no scaling / colour conversion (a requested intermediate format must be the file's), expressions limited
 * to + - * / ^ ( ) numbers and the named constants. */
#ifndef FFSTUB_H
#define FFSTUB_H
#include <errno.h>
#include <inttypes.h>
#include <math.h>
#include <limits.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define AV_VERSION_INT(a, b, c) ((a) << 16 | (b) << 8 | (c))
#define LIBAVUTIL_VERSION_INT AV_VERSION_INT(59, 48, 100)
#define AVERROR(e) (-(e))
#define AVERROR_EOF (-0x5fb9b0bb)
#define FFMIN(a, b) ((a) > (b) ? (b) : (a))
#define FFMAX(a, b) ((a) > (b) ? (a) : (b))
#define AV_LOG_QUIET (-8)
#define AV_LOG_ERROR 16
#define AV_LOG_WARNING 24
#define AV_LOG_INFO 32
void av_log_set_level(int level);
void av_log(void *avcl, int level, const char *fmt, ...);
const char *ffstub_err2str(int err);
#define av_err2str(e) ffstub_err2str(e)

typedef struct AVRational { int num, den; } AVRational;
int av_parse_video_rate(AVRational *rate, const char *str);
int av_reduce(int *dst_num, int *dst_den, int64_t num, int64_t den, int64_t max);
AVRational av_mul_q(AVRational b, AVRational c);

enum AVPixelFormat { AV_PIX_FMT_NONE = -1, AV_PIX_FMT_GRAY8, AV_PIX_FMT_YUV420P, AV_PIX_FMT_YUV444P, AV_PIX_FMT_GBRP, AV_PIX_FMT_GRAYF32LE, AV_PIX_FMT_GBRPF32LE, AV_PIX_FMT_NB };
enum AVColorRange { AVCOL_RANGE_UNSPECIFIED = 0, AVCOL_RANGE_MPEG = 1, AVCOL_RANGE_JPEG = 2, AVCOL_RANGE_NB };
enum AVColorPrimaries { AVCOL_PRI_RESERVED0 = 0, AVCOL_PRI_BT709 = 1, AVCOL_PRI_UNSPECIFIED = 2, AVCOL_PRI_NB };
enum AVColorTransferCharacteristic { AVCOL_TRC_RESERVED0 = 0, AVCOL_TRC_BT709 = 1, AVCOL_TRC_UNSPECIFIED = 2, AVCOL_TRC_RESERVED = 3, AVCOL_TRC_GAMMA22 = 4,
                                      AVCOL_TRC_GAMMA28 = 5, AVCOL_TRC_SMPTE170M = 6, AVCOL_TRC_SMPTE240M = 7, AVCOL_TRC_LINEAR = 8, AVCOL_TRC_LOG = 9,
                                      AVCOL_TRC_LOG_SQRT = 10, AVCOL_TRC_IEC61966_2_4 = 11, AVCOL_TRC_BT1361_ECG = 12, AVCOL_TRC_IEC61966_2_1 = 13, AVCOL_TRC_NB };
enum AVColorSpace { AVCOL_SPC_RGB = 0, AVCOL_SPC_BT709 = 1, AVCOL_SPC_UNSPECIFIED = 2, AVCOL_SPC_NB };
enum AVChromaLocation { AVCHROMA_LOC_UNSPECIFIED = 0, AVCHROMA_LOC_LEFT = 1, AVCHROMA_LOC_CENTER = 2, AVCHROMA_LOC_NB };
enum AVCodecID { AV_CODEC_ID_NONE = 0, AV_CODEC_ID_FFV1 };
const char *av_color_range_name(enum AVColorRange v);
const char *av_color_primaries_name(enum AVColorPrimaries v);
const char *av_color_transfer_name(enum AVColorTransferCharacteristic v);
const char *av_color_space_name(enum AVColorSpace v);
const char *av_chroma_location_name(enum AVChromaLocation v);

#define AV_PIX_FMT_FLAG_BE (1 << 0)
#define AV_PIX_FMT_FLAG_PLANAR (1 << 4)
#define AV_PIX_FMT_FLAG_RGB (1 << 5)
#define AV_PIX_FMT_FLAG_FLOAT (1 << 9)
typedef struct AVComponentDescriptor { int plane, step, offset, shift, depth; } AVComponentDescriptor;
typedef struct AVPixFmtDescriptor {
	const char *name;
	uint8_t nb_components, log2_chroma_w, log2_chroma_h;
	uint64_t flags;
	AVComponentDescriptor comp[4];
} AVPixFmtDescriptor;
const AVPixFmtDescriptor *av_pix_fmt_desc_get(enum AVPixelFormat fmt);
const char *av_get_pix_fmt_name(enum AVPixelFormat fmt);
enum AVPixelFormat av_get_pix_fmt(const char *name);

#define AV_NUM_DATA_POINTERS 8
typedef struct AVFrame {
	uint8_t *data[AV_NUM_DATA_POINTERS];
	int linesize[AV_NUM_DATA_POINTERS];
	int width, height, format;
	int64_t pts;
	enum AVColorRange color_range;
	enum AVColorPrimaries color_primaries;
	enum AVColorTransferCharacteristic color_trc;
	enum AVColorSpace colorspace;
	enum AVChromaLocation chroma_location;
	size_t ffstub_plane_bytes[AV_NUM_DATA_POINTERS];
} AVFrame;
typedef struct AVCodecContext {
	enum AVPixelFormat pix_fmt;
	int width, height;
	enum AVColorRange color_range;
	enum AVColorPrimaries color_primaries;
	enum AVColorTransferCharacteristic color_trc;
	enum AVColorSpace colorspace;
	enum AVChromaLocation chroma_sample_location;
} AVCodecContext;
typedef struct AVFormatContext {
	FILE *fp;
	int is_output;
	uint64_t frames, pos;      /* frames in the file (input) / written so far (output); next frame to read */
	long count_field;          /* file offset of the frame-count field of an output */
} AVFormatContext;
typedef struct AVStream { AVRational r_frame_rate; uint64_t nb_frames; } AVStream;
struct SwsContext;

#define AV_RL32(p) ((uint32_t)((const uint8_t *)(p))[0] | (uint32_t)((const uint8_t *)(p))[1] << 8 | (uint32_t)((const uint8_t *)(p))[2] << 16 | (uint32_t)((const uint8_t *)(p))[3] << 24)
#define AV_RB32(p) ((uint32_t)((const uint8_t *)(p))[3] | (uint32_t)((const uint8_t *)(p))[2] << 8 | (uint32_t)((const uint8_t *)(p))[1] << 16 | (uint32_t)((const uint8_t *)(p))[0] << 24)
#define AV_WL32(p, v) do { uint32_t ffstub_v = (v); uint8_t *ffstub_p = (uint8_t *)(p); ffstub_p[0] = ffstub_v; ffstub_p[1] = ffstub_v >> 8; ffstub_p[2] = ffstub_v >> 16; ffstub_p[3] = ffstub_v >> 24; } while (0)
#define AV_WB32(p, v) do { uint32_t ffstub_v = (v); uint8_t *ffstub_p = (uint8_t *)(p); ffstub_p[3] = ffstub_v; ffstub_p[2] = ffstub_v >> 8; ffstub_p[1] = ffstub_v >> 16; ffstub_p[0] = ffstub_v >> 24; } while (0)

/* libavutil/csp.h */
typedef double (*av_csp_trc_function)(double);
av_csp_trc_function av_csp_trc_func_from_id(enum AVColorTransferCharacteristic trc);
av_csp_trc_function av_csp_trc_func_inv_from_id(enum AVColorTransferCharacteristic trc);

/* libavutil/eval.h */
typedef struct AVExpr AVExpr;
int av_expr_parse(AVExpr **expr, const char *s, const char *const *const_names, const char *const *func1_names, double (*const *funcs1)(void *, double),
                  const char *const *func2_names, double (*const *funcs2)(void *, double, double), int log_offset, void *log_ctx);
double av_expr_eval(AVExpr *e, const double *const_values, void *opaque);
void av_expr_free(AVExpr *e);
#endif
