/* tools/ffstub/ffstub.c -- TEST INFRASTRUCTURE: raw-file implementation of the reference's ffapi interface
 * (/root/reference/include/ffapi.h:39-56, used as it lies) and of the handful of libavutil helpers its tools call.
 * See ffstub.h for the file format.  Written from the interface, not from include/ffapi.c: no codecs, no swscale. */
#include "ffapi.h"
#include <ctype.h>
#include <math.h>
#include <stdarg.h>

/* ------------------------------------------------------------------------------------------------ small libavutil */
static int g_loglevel = AV_LOG_INFO;
void av_log_set_level(int level) { g_loglevel = level; }
void av_log(void *avcl, int level, const char *fmt, ...) {
	(void)avcl;
	if (level > g_loglevel) return;
	va_list ap;
	va_start(ap, fmt);
	vfprintf(stderr, fmt, ap);
	va_end(ap);
}
const char *ffstub_err2str(int err) {
	static char buf[96];
	if (err == AVERROR_EOF) return "End of file";
	snprintf(buf, sizeof buf, "%s", strerror(err < 0 ? -err : err));
	return buf;
}
int av_parse_video_rate(AVRational *rate, const char *str) {
	int n = 0, d = 1;
	if (sscanf(str, "%d/%d", &n, &d) >= 1 && n > 0 && d > 0) { rate->num = n; rate->den = d; return 0; }
	double f = atof(str);
	if (f <= 0) return AVERROR(EINVAL);
	rate->num = (int)lrint(f * 1000); rate->den = 1000;
	return 0;
}
static int64_t gcd64(int64_t a, int64_t b) { while (b) { int64_t t = a % b; a = b; b = t; } return a < 0 ? -a : a; }
int av_reduce(int *dst_num, int *dst_den, int64_t num, int64_t den, int64_t max) {
	int64_t g = gcd64(num, den);
	if (g) { num /= g; den /= g; }
	while ((num > max || den > max) && den > 1) { num /= 2; den /= 2; }      /* coarse; exact whenever the ratio fits */
	*dst_num = (int)num; *dst_den = (int)(den ? den : 1);
	return 1;
}
AVRational av_mul_q(AVRational b, AVRational c) {
	AVRational r;
	av_reduce(&r.num, &r.den, (int64_t)b.num * c.num, (int64_t)b.den * c.den, INT_MAX);
	return r;
}
const char *av_color_range_name(enum AVColorRange v) {
	static const char *n[] = {"unknown", "tv", "pc"};
	return (unsigned)v < AVCOL_RANGE_NB ? n[v] : NULL;
}
const char *av_color_primaries_name(enum AVColorPrimaries v) {
	static const char *n[] = {"reserved", "bt709", "unknown"};
	return (unsigned)v < AVCOL_PRI_NB ? n[v] : NULL;
}
const char *av_color_transfer_name(enum AVColorTransferCharacteristic v) {
	static const char *n[] = {"reserved", "bt709", "unknown", "reserved", "bt470m", "bt470bg", "smpte170m", "smpte240m", "linear", "log100", "log316",
	                          "iec61966-2-4", "bt1361e", "iec61966-2-1"};
	return (unsigned)v < AVCOL_TRC_NB ? n[v] : NULL;
}
const char *av_color_space_name(enum AVColorSpace v) {
	static const char *n[] = {"gbr", "bt709", "unknown"};
	return (unsigned)v < AVCOL_SPC_NB ? n[v] : NULL;
}
const char *av_chroma_location_name(enum AVChromaLocation v) {
	static const char *n[] = {"unspecified", "left", "center"};
	return (unsigned)v < AVCHROMA_LOC_NB ? n[v] : NULL;
}

static const AVPixFmtDescriptor g_desc[AV_PIX_FMT_NB] = {
	[AV_PIX_FMT_GRAY8] = {"gray", 1, 0, 0, 0, {{0, 1, 0, 0, 8}}},
	[AV_PIX_FMT_YUV420P] = {"yuv420p", 3, 1, 1, AV_PIX_FMT_FLAG_PLANAR, {{0, 1, 0, 0, 8}, {1, 1, 0, 0, 8}, {2, 1, 0, 0, 8}}},
	[AV_PIX_FMT_YUV444P] = {"yuv444p", 3, 0, 0, AV_PIX_FMT_FLAG_PLANAR, {{0, 1, 0, 0, 8}, {1, 1, 0, 0, 8}, {2, 1, 0, 0, 8}}},
	[AV_PIX_FMT_GBRP] = {"gbrp", 3, 0, 0, AV_PIX_FMT_FLAG_PLANAR | AV_PIX_FMT_FLAG_RGB, {{2, 1, 0, 0, 8}, {0, 1, 0, 0, 8}, {1, 1, 0, 0, 8}}},
	[AV_PIX_FMT_GRAYF32LE] = {"grayf32le", 1, 0, 0, AV_PIX_FMT_FLAG_FLOAT, {{0, 4, 0, 0, 32}}},
	[AV_PIX_FMT_GBRPF32LE] = {"gbrpf32le", 3, 0, 0, AV_PIX_FMT_FLAG_PLANAR | AV_PIX_FMT_FLAG_RGB | AV_PIX_FMT_FLAG_FLOAT, {{2, 4, 0, 0, 32}, {0, 4, 0, 0, 32}, {1, 4, 0, 0, 32}}},
};
const AVPixFmtDescriptor *av_pix_fmt_desc_get(enum AVPixelFormat fmt) { return (unsigned)fmt < AV_PIX_FMT_NB ? &g_desc[fmt] : NULL; }
const char *av_get_pix_fmt_name(enum AVPixelFormat fmt) { return (unsigned)fmt < AV_PIX_FMT_NB ? g_desc[fmt].name : NULL; }
enum AVPixelFormat av_get_pix_fmt(const char *name) {
	for (int i = 0; i < AV_PIX_FMT_NB; i++)
		if (!strcmp(name, g_desc[i].name)) return (enum AVPixelFormat)i;
	return AV_PIX_FMT_NONE;
}

/* transfer curves: the sRGB pair and the identity; anything else is not offered */
static double trc_srgb(double l) { return l <= 0.0031308 ? 12.92 * l : 1.055 * pow(l, 1.0 / 2.4) - 0.055; }
static double trc_srgb_inv(double e) { return e <= 0.04045 ? e / 12.92 : pow((e + 0.055) / 1.055, 2.4); }
static double trc_linear(double v) { return v; }
av_csp_trc_function av_csp_trc_func_from_id(enum AVColorTransferCharacteristic trc) {
	return trc == AVCOL_TRC_IEC61966_2_1 ? trc_srgb : trc == AVCOL_TRC_LINEAR ? trc_linear : NULL;
}
av_csp_trc_function av_csp_trc_func_inv_from_id(enum AVColorTransferCharacteristic trc) {
	return trc == AVCOL_TRC_IEC61966_2_1 ? trc_srgb_inv : trc == AVCOL_TRC_LINEAR ? trc_linear : NULL;
}

/* expressions: numbers, the caller's named constants, + - * / ^, unary minus, parentheses */
struct AVExpr { char *src; char **names; int nnames; };
typedef struct { const char *p; const struct AVExpr *e; const double *vals; int err; } ExprParse;
static double expr_sum(ExprParse *s);
static void expr_ws(ExprParse *s) { while (isspace((unsigned char)*s->p)) s->p++; }
static double expr_atom(ExprParse *s) {
	expr_ws(s);
	if (*s->p == '(') { s->p++; double v = expr_sum(s); expr_ws(s); if (*s->p == ')') s->p++; else s->err = 1; return v; }
	if (*s->p == '-') { s->p++; return -expr_atom(s); }
	if (isdigit((unsigned char)*s->p) || *s->p == '.') { char *end; double v = strtod(s->p, &end); s->p = end; return v; }
	if (isalpha((unsigned char)*s->p)) {
		const char *b = s->p;
		while (isalnum((unsigned char)*s->p) || *s->p == '_') s->p++;
		for (int i = 0; i < s->e->nnames; i++)
			if (strlen(s->e->names[i]) == (size_t)(s->p - b) && !strncmp(b, s->e->names[i], s->p - b)) return s->vals ? s->vals[i] : 0.0;
	}
	s->err = 1;
	return 0;
}
static double expr_pow(ExprParse *s) {
	double v = expr_atom(s);
	expr_ws(s);
	if (*s->p == '^') { s->p++; v = pow(v, expr_pow(s)); }
	return v;
}
static double expr_prod(ExprParse *s) {
	double v = expr_pow(s);
	for (;;) {
		expr_ws(s);
		if (*s->p == '*') { s->p++; v *= expr_pow(s); }
		else if (*s->p == '/') { s->p++; v /= expr_pow(s); }
		else return v;
	}
}
static double expr_sum(ExprParse *s) {
	double v = expr_prod(s);
	for (;;) {
		expr_ws(s);
		if (*s->p == '+') { s->p++; v += expr_prod(s); }
		else if (*s->p == '-') { s->p++; v -= expr_prod(s); }
		else return v;
	}
}
static double expr_run(const struct AVExpr *e, const double *vals, int *err) {
	ExprParse s = {e->src, e, vals, 0};
	double v = expr_sum(&s);
	expr_ws(&s);
	if (*s.p) s.err = 1;
	if (err) *err = s.err;
	return v;
}
int av_expr_parse(AVExpr **expr, const char *str, const char *const *const_names, const char *const *func1_names, double (*const *funcs1)(void *, double),
                  const char *const *func2_names, double (*const *funcs2)(void *, double, double), int log_offset, void *log_ctx) {
	(void)func1_names; (void)funcs1; (void)func2_names; (void)funcs2; (void)log_offset; (void)log_ctx;
	struct AVExpr *e = calloc(1, sizeof *e);
	e->src = strdup(str);
	while (const_names && const_names[e->nnames]) e->nnames++;
	e->names = calloc(e->nnames + 1, sizeof *e->names);
	for (int i = 0; i < e->nnames; i++) e->names[i] = strdup(const_names[i]);
	int err = 0;
	expr_run(e, NULL, &err);
	if (err) { av_expr_free(e); fprintf(stderr, "ffstub: cannot parse expression '%s'\n", str); return AVERROR(EINVAL); }
	*expr = e;
	return 0;
}
double av_expr_eval(AVExpr *e, const double *const_values, void *opaque) { (void)opaque; return expr_run(e, const_values, NULL); }
void av_expr_free(AVExpr *e) {
	if (!e) return;
	for (int i = 0; i < e->nnames; i++) free(e->names[i]);
	free(e->names); free(e->src); free(e);
}

/* ------------------------------------------------------------------------------------------------ ffapi over DSPV files */
bool ffapi_pixfmts_8bit_pel(const AVPixFmtDescriptor *d) {
	if (d->flags & AV_PIX_FMT_FLAG_FLOAT) return false;
	for (int i = 0; i < d->nb_components; i++)
		if (d->comp[i].depth != 8) return false;
	return true;
}
bool ffapi_pixfmts_32_bit_float_pel(const AVPixFmtDescriptor *d) { return (d->flags & AV_PIX_FMT_FLAG_FLOAT) && d->comp[0].depth == 32; }

void ffapi_parse_color_props(FFColorProperties *c, const char *props) {
	c->color_range = AVCOL_RANGE_UNSPECIFIED; c->color_primaries = AVCOL_PRI_UNSPECIFIED; c->color_trc = AVCOL_TRC_UNSPECIFIED;
	c->color_space = AVCOL_SPC_UNSPECIFIED; c->chroma_location = AVCHROMA_LOC_UNSPECIFIED; c->pix_fmt = AV_PIX_FMT_NONE;
	if (!props) return;
	const char *k = strstr(props, "pixel_format=");
	if (k) {
		char name[32] = {0};
		sscanf(k + 13, "%31[^:]", name);
		c->pix_fmt = av_get_pix_fmt(name);
	}
	if ((k = strstr(props, "color_trc="))) {
		char name[32] = {0};
		sscanf(k + 10, "%31[^:]", name);
		for (int t = 0; t < AVCOL_TRC_NB; t++)
			if (!strcmp(name, av_color_transfer_name(t))) c->color_trc = t;
	}
}

static int plane_count(const AVPixFmtDescriptor *d) {
	int n = 0;
	for (int i = 0; i < d->nb_components; i++) n = FFMAX(n, d->comp[i].plane + 1);
	return n;
}
/* geometry of plane p: the component stored there decides subsampling and sample size */
static void plane_geom(const AVPixFmtDescriptor *d, int p, int w, int h, int *pw, int *ph, int *step) {
	for (int i = 0; i < d->nb_components; i++)
		if (d->comp[i].plane == p) {
			const int sub = (i == 1 || i == 2) && !(d->flags & AV_PIX_FMT_FLAG_RGB);
			*pw = sub ? -(-w >> d->log2_chroma_w) : w;
			*ph = sub ? -(-h >> d->log2_chroma_h) : h;
			*step = d->comp[i].step;
			return;
		}
	*pw = *ph = *step = 0;
}
static int frame_buffers(FFContext *ctx, AVFrame *f) {
	const AVPixFmtDescriptor *d = ctx->pixdesc;
	f->width = ctx->codec->width; f->height = ctx->codec->height; f->format = ctx->codec->pix_fmt;
	for (int p = 0; p < plane_count(d); p++) {
		int pw, ph, step;
		plane_geom(d, p, f->width, f->height, &pw, &ph, &step);
		f->linesize[p] = pw * step;
		f->ffstub_plane_bytes[p] = (size_t)pw * step * ph;
		if (!(f->data[p] = calloc(1, f->ffstub_plane_bytes[p] ? f->ffstub_plane_bytes[p] : 1))) return AVERROR(ENOMEM);
	}
	return 0;
}
static FFContext *ctx_new(void) {
	FFContext *c = calloc(1, sizeof *c);
	c->fmt = calloc(1, sizeof *c->fmt);
	c->codec = calloc(1, sizeof *c->codec);
	c->st = calloc(1, sizeof *c->st);
	return c;
}

FFContext *ffapi_open_input(const char *file, const char *options, const char *format, FFColorProperties *color_props, ffapi_pix_fmt_filter *filter,
                            uint8_t *components, int (*widths)[4], int (*heights)[4], uint64_t *frames, AVRational *rate, bool calc_frames, int *averror) {
	(void)options; (void)format; (void)calc_frames;
	int err = 0;
	FFContext *in = ctx_new();
	char name[32] = {0};
	unsigned long long nframes = 0;
	int w = 0, h = 0, rn = 25, rd = 1;
	if (!(in->fmt->fp = fopen(file, "rb"))) { err = AVERROR(errno); goto fail; }
	char line[256];
	if (!fgets(line, sizeof line, in->fmt->fp) || sscanf(line, "DSPV1 %31s %d %d %llu %d %d", name, &w, &h, &nframes, &rn, &rd) != 6) { err = AVERROR(EINVAL); goto fail; }
	in->codec->pix_fmt = av_get_pix_fmt(name);
	if (in->codec->pix_fmt == AV_PIX_FMT_NONE) { err = AVERROR(EINVAL); goto fail; }
	in->codec->width = w; in->codec->height = h;
	in->codec->color_range = (g_desc[in->codec->pix_fmt].flags & AV_PIX_FMT_FLAG_RGB) ? AVCOL_RANGE_JPEG : AVCOL_RANGE_MPEG;
	in->codec->color_primaries = AVCOL_PRI_BT709; in->codec->color_trc = AVCOL_TRC_BT709;
	in->codec->colorspace = (g_desc[in->codec->pix_fmt].flags & AV_PIX_FMT_FLAG_RGB) ? AVCOL_SPC_RGB : AVCOL_SPC_BT709;
	in->codec->chroma_sample_location = AVCHROMA_LOC_LEFT;
	in->pixdesc = (AVPixFmtDescriptor *)&g_desc[in->codec->pix_fmt];
	in->fmt->frames = nframes;
	in->st->nb_frames = nframes; in->st->r_frame_rate = (AVRational){rn, rd};
	if (filter && !filter(in->pixdesc)) { fprintf(stderr, "ffstub: pixel format %s is not accepted by this tool (no conversion in the stub)\n", name); err = AVERROR(EINVAL); goto fail; }
	FFColorProperties own;
	if (!color_props) { ffapi_parse_color_props(&own, NULL); color_props = &own; }
	if (color_props->pix_fmt != AV_PIX_FMT_NONE && color_props->pix_fmt != in->codec->pix_fmt) { fprintf(stderr, "ffstub: no pixel format conversion\n"); err = AVERROR(EINVAL); goto fail; }
	color_props->pix_fmt = in->codec->pix_fmt;
	if (color_props->color_range == AVCOL_RANGE_UNSPECIFIED) color_props->color_range = in->codec->color_range;
	if (color_props->color_primaries == AVCOL_PRI_UNSPECIFIED) color_props->color_primaries = in->codec->color_primaries;
	if (color_props->color_trc == AVCOL_TRC_UNSPECIFIED) color_props->color_trc = in->codec->color_trc;
	if (color_props->color_space == AVCOL_SPC_UNSPECIFIED) color_props->color_space = in->codec->colorspace;
	if (color_props->chroma_location == AVCHROMA_LOC_UNSPECIFIED) color_props->chroma_location = in->codec->chroma_sample_location;
	in->color_props = *color_props;
	if (rate) *rate = in->st->r_frame_rate;
	if (frames) *frames = nframes;
	if (components) *components = in->pixdesc->nb_components;
	for (int i = 0; i < in->pixdesc->nb_components; i++) {
		int pw, ph, step;
		plane_geom(in->pixdesc, in->pixdesc->comp[i].plane, w, h, &pw, &ph, &step);
		if (widths) (*widths)[i] = pw;
		if (heights) (*heights)[i] = ph;
	}
	if (averror) *averror = 0;
	return in;
fail:
	ffapi_close(in);
	if (averror) *averror = err;
	return NULL;
}

FFContext *ffapi_open_output(const char *file, const char *options, const char *format, const char *encoder, enum AVCodecID preferred_encoder,
                             const FFColorProperties *props, size_t width, size_t height, AVRational rate, int *averror) {
	(void)options; (void)format; (void)encoder; (void)preferred_encoder;
	FFContext *out = ctx_new();
	int err = 0;
	enum AVPixelFormat fmt = props ? props->pix_fmt : AV_PIX_FMT_NONE;
	if (fmt == AV_PIX_FMT_NONE) { err = AVERROR(EINVAL); goto fail; }
	if (!(out->fmt->fp = fopen(file, "wb"))) { err = AVERROR(errno); goto fail; }
	out->fmt->is_output = 1;
	out->codec->pix_fmt = fmt; out->codec->width = (int)width; out->codec->height = (int)height;
	out->codec->color_range = props->color_range; out->codec->color_primaries = props->color_primaries; out->codec->color_trc = props->color_trc;
	out->codec->colorspace = props->color_space; out->codec->chroma_sample_location = props->chroma_location;
	out->pixdesc = (AVPixFmtDescriptor *)&g_desc[fmt];
	out->color_props = *props;
	out->st->r_frame_rate = rate;
	fprintf(out->fmt->fp, "DSPV1 %s %d %d ", g_desc[fmt].name, (int)width, (int)height);
	out->fmt->count_field = ftell(out->fmt->fp);
	fprintf(out->fmt->fp, "%020llu %d %d\n", 0ull, rate.num > 0 ? rate.num : 25, rate.den > 0 ? rate.den : 1);
	if (averror) *averror = 0;
	return out;
fail:
	ffapi_close(out);
	if (averror) *averror = err;
	return NULL;
}

AVFrame *ffapi_alloc_frame(FFContext *ctx) {
	AVFrame *f = calloc(1, sizeof *f);
	if (f && frame_buffers(ctx, f)) { ffapi_free_frame(f); return NULL; }
	return f;
}
void ffapi_free_frame(AVFrame *f) {
	if (!f) return;
	for (int p = 0; p < AV_NUM_DATA_POINTERS; p++) free(f->data[p]);
	free(f);
}
void ffapi_clear_frame(AVFrame *f) {
	for (int p = 0; p < AV_NUM_DATA_POINTERS; p++)
		if (f->data[p]) memset(f->data[p], 0, f->ffstub_plane_bytes[p]);
}
int ffapi_read_frame(FFContext *in, AVFrame *f) {
	if (!f->data[0] && frame_buffers(in, f)) return AVERROR(ENOMEM);
	if (in->fmt->pos >= in->fmt->frames) return AVERROR_EOF;
	for (int p = 0; p < AV_NUM_DATA_POINTERS && f->data[p]; p++)
		if (fread(f->data[p], 1, f->ffstub_plane_bytes[p], in->fmt->fp) != f->ffstub_plane_bytes[p]) return AVERROR_EOF;
	f->pts = (int64_t)in->fmt->pos++;
	return 0;
}
int ffapi_seek_frame(FFContext *ctx, uint64_t *offset, void (*progress)(uint64_t)) {
	if (!*offset) return 0;
	AVFrame *f = calloc(1, sizeof *f);
	uint64_t seek = 0;
	int err = 0;
	for (; seek < *offset && !(err = ffapi_read_frame(ctx, f)); seek++)
		if (progress) progress(seek);
	ffapi_free_frame(f);
	*offset = seek;
	return err;
}
int ffapi_write_frame(FFContext *out, AVFrame *f) {
	for (int p = 0; p < AV_NUM_DATA_POINTERS && f->data[p]; p++)
		if (fwrite(f->data[p], 1, f->ffstub_plane_bytes[p], out->fmt->fp) != f->ffstub_plane_bytes[p]) return AVERROR(EIO);
	out->fmt->frames++;
	return 0;
}
int ffapi_close(FFContext *ctx) {
	if (!ctx) return 0;
	int err = 0;
	if (ctx->fmt && ctx->fmt->fp) {
		if (ctx->fmt->is_output) {
			fseek(ctx->fmt->fp, ctx->fmt->count_field, SEEK_SET);
			fprintf(ctx->fmt->fp, "%020llu", (unsigned long long)ctx->fmt->frames);
		}
		if (fclose(ctx->fmt->fp)) err = AVERROR(EIO);
	}
	free(ctx->fmt); free(ctx->codec); free(ctx->st); free(ctx);
	return err;
}
