/*
 * tools/wandstub -- a few-dozen-line stand-in for the slice of the MagickWand API that dspfun's spec / ispec use
 * (spec/spec.c:46-61,141-158; spec/ispec.c:54-98,170-186), so that the reference's UNMODIFIED sources can be
 * compiled and run against libdspdct on a box without ImageMagick.  Images are ".dspraw" files:
 *     "DSPRAW <w> <h> <d>\n" { "PROP <key> <value>\n" } "DATA\n" <h*w*d float64 samples, [y][x][channel]>
 * This is test tooling for the drop-in claim, not an image library.
 */
#ifndef DSP_WANDSTUB_H
#define DSP_WANDSTUB_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum { MagickFalse = 0, MagickTrue = 1 } MagickBooleanType;
typedef enum { UndefinedException = 0 } ExceptionType;
typedef enum { UndefinedPixel, CharPixel, DoublePixel, FloatPixel } StorageType;
typedef enum { UndefinedColorspace, RGBColorspace, sRGBColorspace } ColorspaceType;
typedef struct _MagickWand MagickWand;

void MagickWandGenesis(void);
void MagickWandTerminus(void);
MagickWand *NewMagickWand(void);
MagickWand *DestroyMagickWand(MagickWand *);
MagickBooleanType MagickReadImage(MagickWand *, const char *);
MagickBooleanType MagickWriteImage(MagickWand *, const char *);
char *MagickGetException(const MagickWand *, ExceptionType *);
void *RelinquishMagickMemory(void *);
void *MagickRelinquishMemory(void *);
size_t MagickGetImageWidth(MagickWand *);
size_t MagickGetImageHeight(MagickWand *);
MagickBooleanType MagickTransformImageColorspace(MagickWand *, ColorspaceType);
MagickBooleanType MagickSetImageColorspace(MagickWand *, ColorspaceType);
ColorspaceType MagickGetImageColorspace(MagickWand *);      /* sRGB unless the file carries PROP colorspace RGB */
size_t MagickGetImageDepth(MagickWand *);                   /* PROP depth N, default 8 */
MagickBooleanType MagickExportImagePixels(MagickWand *, long x, long y, size_t w, size_t h, const char *map, StorageType, void *);
MagickBooleanType MagickConstituteImage(MagickWand *, size_t w, size_t h, const char *map, StorageType, const void *);
MagickBooleanType MagickSetImageProperty(MagickWand *, const char *, const char *);
char *MagickGetImageProperty(MagickWand *, const char *);

#ifdef __cplusplus
}
#endif
#endif
