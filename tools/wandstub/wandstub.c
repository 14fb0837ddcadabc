/* tools/wandstub/wandstub.c -- see wand/MagickWand.h.  Raw-file "MagickWand" for running the reference's
 * unmodified spec / ispec against libdspdct. */
#include "wand/MagickWand.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct _MagickWand {
	size_t w, h, d;
	double *px;
	char *keys[8], *vals[8];
	int nprops;
	char err[256];
};

void MagickWandGenesis(void) {}
void MagickWandTerminus(void) {}
MagickWand *NewMagickWand(void) { return calloc(1, sizeof(MagickWand)); }
MagickWand *DestroyMagickWand(MagickWand *w) {
	if (!w) return NULL;
	free(w->px);
	for (int i = 0; i < w->nprops; i++) { free(w->keys[i]); free(w->vals[i]); }
	free(w);
	return NULL;
}
char *MagickGetException(const MagickWand *w, ExceptionType *t) { if (t) *t = UndefinedException; return strdup(w->err); }
void *RelinquishMagickMemory(void *p) { free(p); return NULL; }
void *MagickRelinquishMemory(void *p) { free(p); return NULL; }
size_t MagickGetImageWidth(MagickWand *w) { return w->w; }
size_t MagickGetImageHeight(MagickWand *w) { return w->h; }
MagickBooleanType MagickTransformImageColorspace(MagickWand *w, ColorspaceType c) { (void)w; (void)c; return MagickTrue; }
MagickBooleanType MagickSetImageColorspace(MagickWand *w, ColorspaceType c) { (void)w; (void)c; return MagickTrue; }

static const char *wand_prop(MagickWand *w, const char *key) {
	for (int i = 0; i < w->nprops; i++)
		if (!strcmp(w->keys[i], key)) return w->vals[i];
	return NULL;
}
ColorspaceType MagickGetImageColorspace(MagickWand *w) {
	const char *v = wand_prop(w, "colorspace");
	return v && !strcmp(v, "RGB") ? RGBColorspace : sRGBColorspace;
}
size_t MagickGetImageDepth(MagickWand *w) {
	const char *v = wand_prop(w, "depth");
	return v ? (size_t)strtoul(v, NULL, 10) : 8;
}

MagickBooleanType MagickReadImage(MagickWand *w, const char *path) {
	FILE *f = fopen(path, "rb");
	if (!f) { snprintf(w->err, sizeof w->err, "wandstub: cannot open %s", path); return MagickFalse; }
	char line[8192];
	if (!fgets(line, sizeof line, f) || sscanf(line, "DSPRAW %zu %zu %zu", &w->w, &w->h, &w->d) != 3) {
		snprintf(w->err, sizeof w->err, "wandstub: %s is not a DSPRAW file", path);
		fclose(f);
		return MagickFalse;
	}
	while (fgets(line, sizeof line, f) && strncmp(line, "DATA", 4)) {
		char key[64];
		int off = 0;
		if (sscanf(line, "PROP %63s %n", key, &off) == 1 && w->nprops < 8) {
			line[strcspn(line, "\n")] = 0;
			w->keys[w->nprops] = strdup(key);
			w->vals[w->nprops++] = strdup(line + off);
		}
	}
	size_t l = w->w * w->h * w->d;
	w->px = malloc(sizeof(double) * l);
	if (fread(w->px, sizeof(double), l, f) != l) { snprintf(w->err, sizeof w->err, "wandstub: short read"); fclose(f); return MagickFalse; }
	fclose(f);
	return MagickTrue;
}

MagickBooleanType MagickWriteImage(MagickWand *w, const char *path) {
	FILE *f = fopen(path, "wb");
	if (!f) { snprintf(w->err, sizeof w->err, "wandstub: cannot create %s", path); return MagickFalse; }
	fprintf(f, "DSPRAW %zu %zu %zu\n", w->w, w->h, w->d);
	for (int i = 0; i < w->nprops; i++) fprintf(f, "PROP %s %s\n", w->keys[i], w->vals[i]);
	fprintf(f, "DATA\n");
	fwrite(w->px, sizeof(double), w->w * w->h * w->d, f);
	fclose(f);
	return MagickTrue;
}

MagickBooleanType MagickExportImagePixels(MagickWand *w, long x, long y, size_t cw, size_t ch, const char *map, StorageType t, void *out) {
	(void)x; (void)y;
	size_t l = cw * ch * strlen(map);
	if (strlen(map) != w->d || cw != w->w || ch != w->h) return MagickFalse;
	for (size_t i = 0; i < l; i++) {
		double v = w->px[i];
		if (t == FloatPixel) ((float *)out)[i] = (float)v;
		else if (t == DoublePixel) ((double *)out)[i] = v;
		else { double q = v < 0 ? 0 : v > 1 ? 1 : v; ((unsigned char *)out)[i] = (unsigned char)(q * 255.0 + 0.5); }
	}
	return MagickTrue;
}

MagickBooleanType MagickConstituteImage(MagickWand *w, size_t cw, size_t ch, const char *map, StorageType t, const void *in) {
	w->w = cw; w->h = ch; w->d = strlen(map);
	size_t l = cw * ch * w->d;
	free(w->px);
	w->px = malloc(sizeof(double) * l);
	for (size_t i = 0; i < l; i++)
		w->px[i] = t == FloatPixel ? (double)((const float *)in)[i] : t == DoublePixel ? ((const double *)in)[i] : ((const unsigned char *)in)[i] / 255.0;
	return MagickTrue;
}

MagickBooleanType MagickSetImageProperty(MagickWand *w, const char *k, const char *v) {
	if (w->nprops >= 8) return MagickFalse;
	w->keys[w->nprops] = strdup(k);
	w->vals[w->nprops++] = strdup(v);
	return MagickTrue;
}

char *MagickGetImageProperty(MagickWand *w, const char *k) {
	for (int i = 0; i < w->nprops; i++)
		if (!strcmp(w->keys[i], k)) return strdup(w->vals[i]);
	return NULL;
}
