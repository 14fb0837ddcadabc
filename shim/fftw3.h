/*
 * shim/fftw3.h -- source-compatible stand-in for <fftw3.h>, covering exactly the FFTW surface dspfun uses, on top
 * of libdspdct (include/dsp_dct.h).  Building a dspfun tool with `-Ishim` ahead of the system include path and
 * linking `-ldspdct` instead of `-lfftw3f/-lfftw3` moves its transforms to the GPU with no source change:
 *
 *   fftw(plan_many_r2r)  spec/spec.c:63  spec/ispec.c:165  zoom/zoom.c:263  scan/scan.c:292,359  motion/motion.c:535,549
 *   fftw(plan_r2r_2d)    applybasis/draw.c:74
 *   fftw(execute)        spec/spec.c:64  ...  fftw(destroy_plan) spec/spec.c:65 ...
 *   fftw(alloc_real) / fftw(free)            spec/spec.c:59,143  motion/motion.c:500,826  scan/scan.c:352-353
 *   fftw(init_threads) / fftw(plan_with_nthreads) / fftw(cleanup_threads)   motion/motion.c:485-486,837  scan/scan.c:289-290
 *   fftw(import_wisdom_from_filename) / fftw(export_wisdom_to_filename)     motion/motion.c:519,557
 *   fftw(cleanup)        zoom/zoom.c:266  motion/motion.c:836
 *
 * The prefix comes from include/precision.h:115: fftwf_ (COEFF_PRECISION=F), fftw_ (D).  fftwl_ (L, long double)
 * has no GPU equivalent and fails at compile time.  No reference caller checks a plan for NULL
 * (SURVEY.md 2.3), so a failing plan call prints dsp_dct_last_error() and aborts instead of returning NULL.
 */
#ifndef DSP_SHIM_FFTW3_H
#define DSP_SHIM_FFTW3_H

#include <stdio.h>
#include <stdlib.h>
#include "../include/dsp_dct.h"

#ifdef __cplusplus
extern "C" {
#endif

/* same numeric values as FFTW's enum fftw_r2r_kind_do_not_use_me */
typedef enum {
	FFTW_R2HC = 0, FFTW_HC2R = 1, FFTW_DHT = 2, FFTW_REDFT00 = 3, FFTW_REDFT01 = 4, FFTW_REDFT10 = 5, FFTW_REDFT11 = 6,
	FFTW_RODFT00 = 7, FFTW_RODFT01 = 8, FFTW_RODFT10 = 9, FFTW_RODFT11 = 10
} fftw_r2r_kind;
typedef fftw_r2r_kind fftwf_r2r_kind;
typedef fftw_r2r_kind fftwl_r2r_kind;

#define FFTW_MEASURE    (0U)
#define FFTW_EXHAUSTIVE (1U << 3)
#define FFTW_PATIENT    (1U << 5)
#define FFTW_ESTIMATE   (1U << 6)

typedef dsp_dct_plan fftwf_plan;
typedef dsp_dct_plan fftw_plan;

static inline dsp_dct_plan dsp_shim_checked(dsp_dct_plan p) {
	if (!p) {
		fprintf(stderr, "libdspdct: %s\n", dsp_dct_last_error());
		abort();
	}
	return p;
}

#define DSP_SHIM_DEFINE(PFX, R, PREC)                                                                                \
	static inline PFX##_plan PFX##_plan_many_r2r(int rank, const int *n, int howmany, R *in, const int *inembed,      \
	                                             int istride, int idist, R *out, const int *onembed, int ostride,  \
	                                             int odist, const fftw_r2r_kind *kind, unsigned flags) {              \
		return dsp_shim_checked(dsp_dct_plan_many(PREC, rank, n, howmany, in, inembed, istride, idist, out, onembed,   \
		                                          ostride, odist, (const int *)kind, flags));                         \
	}                                                                                                                 \
	static inline PFX##_plan PFX##_plan_r2r_2d(int n0, int n1, R *in, R *out, fftw_r2r_kind k0, fftw_r2r_kind k1,       \
	                                           unsigned flags) {                                                      \
		return dsp_shim_checked(dsp_dct_plan_2d(PREC, n0, n1, in, out, (int)k0, (int)k1, flags));                      \
	}                                                                                                                 \
	static inline void PFX##_execute(const PFX##_plan p) { dsp_dct_execute(p); }                                       \
	static inline void PFX##_destroy_plan(PFX##_plan p) { dsp_dct_destroy(p); }                                        \
	static inline R *PFX##_alloc_real(size_t n) { return (R *)dsp_dct_alloc(n * sizeof(R)); }                          \
	static inline void *PFX##_malloc(size_t n) { return dsp_dct_alloc(n); }                                            \
	static inline void PFX##_free(void *p) { dsp_dct_free(p); }                                                        \
	static inline void PFX##_cleanup(void) { dsp_dct_cleanup(); }                                                      \
	static inline int PFX##_init_threads(void) { return 1; }             /* non-zero = success */                      \
	static inline void PFX##_plan_with_nthreads(int nthreads) { dsp_dct_plan_with_ngpus(nthreads); }                                      \
	static inline void PFX##_cleanup_threads(void) {}                                                                  \
	static inline int PFX##_import_wisdom_from_filename(const char *f) { (void)f; return 0; }  /* 0 = nothing read */  \
	static inline int PFX##_export_wisdom_to_filename(const char *f) { (void)f; return 0; }

DSP_SHIM_DEFINE(fftwf, float, 'f')
DSP_SHIM_DEFINE(fftw, double, 'd')
#undef DSP_SHIM_DEFINE

/* COEFF_PRECISION=L: there is no long double on the GPU */
#if defined(__GNUC__)
typedef struct dsp_shim_no_long_double *fftwl_plan;
fftwl_plan fftwl_plan_many_r2r(int, const int *, int, long double *, const int *, int, int, long double *, const int *, int,
                               int, const fftw_r2r_kind *, unsigned)
    __attribute__((error("libdspdct: COEFF_PRECISION=L (fftwl_*) is not supported on the GPU; build with F or D")));
long double *fftwl_alloc_real(size_t)
    __attribute__((error("libdspdct: COEFF_PRECISION=L (fftwl_*) is not supported on the GPU; build with F or D")));
#endif

#ifdef __cplusplus
}
#endif
#endif
