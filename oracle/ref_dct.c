/*
 * oracle/ref_dct.c -- TEST INFRASTRUCTURE ONLY.  Never linked into, imported by, or called from the
 * product library (dspfun_b200/csrc).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it, and only as the checker.
 *
 * CPU restatement of the arithmetic behind dspfun's transform hot path.  In the reference that arithmetic is
 * not in the repository: it lives in FFTW 3 (libfftw3f / libfftw3 / libfftw3l, un-vendored, no version pinned;
 * selected by pkg-config name only -- /root/reference/spec/Makefile:6,10), reached through the planner call
 * sites spec/spec.c:63, spec/ispec.c:165, zoom/zoom.c:263, scan/scan.c:292,359, motion/motion.c:535-538,549-552
 * and applybasis/draw.c:74.  What is restated here is FFTW's published definition of the two r2r kinds the
 * reference plans (FFTW manual, "1d Real-even DFTs (DCTs)"):
 *
 *     REDFT10:  Y_k = 2 * sum_{j=0}^{n-1} X_j cos(pi (j + 1/2) k / n)
 *     REDFT01:  Y_k = X_0 + 2 * sum_{j=1}^{n-1} X_j cos(pi j (k + 1/2) / n)
 *
 * applied separably over rank dimensions, with the advanced interface's howmany/stride/dist/embed addressing
 * (FFTW manual, "Advanced Real-to-real Transforms"): element (j_0..j_{r-1}) of transform b lives at
 *     base + b*dist + stride * (j_{r-1} + embed[r-1] * (j_{r-2} + embed[r-2] * (...)))
 *
 * Every sum is evaluated in long double straight from the definition (O(n^2) per line) and rounded once to the
 * coefficient type at the end, so this is the arbiter, not a fast path.
 *
 * PARITY PIN: the reference has no tests or golden vectors of its own for this path, so
 * parity against the reference's tests is unpinned.  The restatement is instead pinned to the FFTW-generated
 * known-answer vectors shipped with scipy (scipy/fftpack/tests/fftw_{single,double,longdouble}_ref.npz,
 * copied to tests/golden/fftw_dct_ref.npz by tests/golden/make_golden.py) -- see tests/test_oracle.py.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stddef.h>
#include <pthread.h>
#include <unistd.h>

typedef long double ld;

enum { REF_REDFT10 = 10, REF_REDFT01 = 1 };

/* cos(pi*m/(2n)) for m in [0,4n): every cosine either definition needs is ctab[(odd*idx) mod 4n] */
static ld *make_ctab(int n) {
	ld *t = malloc(sizeof(ld) * 4 * (size_t)n);
	const ld pi = 3.141592653589793238462643383279502884L;
	for (long m = 0; m < 4L * n; m++)
		t[m] = cosl(pi * (ld)m / (2 * (ld)n));
	return t;
}

/* one line, REDFT10, by definition */
static void line_redft10(const ld *x, ld *y, int n, const ld *ctab) {
	const long p = 4L * n;
	for (long k = 0; k < n; k++) {
		ld s = 0;
		for (long j = 0; j < n; j++)
			s += x[j] * ctab[((2 * j + 1) * k) % p];
		y[k] = 2 * s;
	}
}

/* one line, REDFT01, by definition */
static void line_redft01(const ld *x, ld *y, int n, const ld *ctab) {
	const long p = 4L * n;
	for (long k = 0; k < n; k++) {
		ld s = 0;
		for (long j = 1; j < n; j++)
			s += x[j] * ctab[(j * (2 * k + 1)) % p];
		y[k] = x[0] + 2 * s;
	}
}

/* transform axis `ax` of a dense row-major long-double box dims[rank] in place (lines split over pthreads) */
struct axis_job { ld *box; size_t inner, lines; int n, kind; const ld *ctab; size_t lo, hi; };

static void *axis_worker(void *arg) {
	struct axis_job *j = arg;
	const int n = j->n;
	ld *x = malloc(sizeof(ld) * (size_t)n), *y = malloc(sizeof(ld) * (size_t)n);
	for (size_t l = j->lo; l < j->hi; l++) {
		size_t o = l / j->inner, in = l % j->inner;
		ld *base = j->box + o * (size_t)n * j->inner + in;
		for (int i = 0; i < n; i++) x[i] = base[(size_t)i * j->inner];
		if (j->kind == REF_REDFT10) line_redft10(x, y, n, j->ctab);
		else                        line_redft01(x, y, n, j->ctab);
		for (int i = 0; i < n; i++) base[(size_t)i * j->inner] = y[i];
	}
	free(x); free(y);
	return NULL;
}

static int ref_threads = 0;
void ref_set_threads(int t) { ref_threads = t; }

static void axis_pass(ld *box, int rank, const int *dims, int ax, int kind) {
	size_t inner = 1, outer = 1;
	for (int i = ax + 1; i < rank; i++) inner *= (size_t)dims[i];
	for (int i = 0; i < ax; i++) outer *= (size_t)dims[i];
	const int n = dims[ax];
	ld *ctab = make_ctab(n);
	size_t lines = outer * inner;
	int nt = ref_threads > 0 ? ref_threads : (int)sysconf(_SC_NPROCESSORS_ONLN);
	if (nt > 64) nt = 64;
	if ((size_t)nt > lines) nt = (int)lines;
	if (nt < 1) nt = 1;
	pthread_t th[64];
	struct axis_job jobs[64];
	for (int t = 0; t < nt; t++) {
		jobs[t] = (struct axis_job){box, inner, lines, n, kind, ctab, lines * (size_t)t / (size_t)nt, lines * (size_t)(t + 1) / (size_t)nt};
		if (nt == 1) axis_worker(&jobs[t]);
		else pthread_create(&th[t], NULL, axis_worker, &jobs[t]);
	}
	if (nt > 1) for (int t = 0; t < nt; t++) pthread_join(th[t], NULL);
	free(ctab);
}

static size_t embed_offset(int rank, const int *embed, const size_t *idx) {
	size_t off = 0;
	for (int i = 0; i < rank; i++) off = off * (size_t)embed[i] + idx[i];
	return off;
}

#define DEFINE_MANY(NAME, T)                                                                              \
int NAME(int rank, const int *n, int howmany,                                                             \
         const T *in, const int *inembed, int istride, int idist,                                         \
         T *out, const int *onembed, int ostride, int odist, const int *kind) {                           \
	if (rank < 1 || rank > 3) return -1;                                                                  \
	for (int i = 0; i < rank; i++)                                                                        \
		if (n[i] < 1 || (kind[i] != REF_REDFT10 && kind[i] != REF_REDFT01)) return -1;                    \
	if (!inembed) inembed = n;                                                                            \
	if (!onembed) onembed = n;                                                                            \
	size_t total = 1;                                                                                     \
	for (int i = 0; i < rank; i++) total *= (size_t)n[i];                                                 \
	ld *box = malloc(sizeof(ld) * total);                                                                 \
	if (!box) return -2;                                                                                  \
	for (int b = 0; b < howmany; b++) {                                                                   \
		size_t idx[3] = {0, 0, 0};                                                                        \
		for (size_t e = 0; e < total; e++) {                                                              \
			size_t r = e;                                                                                 \
			for (int i = rank - 1; i >= 0; i--) { idx[i] = r % (size_t)n[i]; r /= (size_t)n[i]; }         \
			box[e] = (ld)in[(size_t)b * (size_t)idist + (size_t)istride * embed_offset(rank, inembed, idx)]; \
		}                                                                                                 \
		for (int ax = rank - 1; ax >= 0; ax--) axis_pass(box, rank, n, ax, kind[ax]);                     \
		for (size_t e = 0; e < total; e++) {                                                              \
			size_t r = e;                                                                                 \
			for (int i = rank - 1; i >= 0; i--) { idx[i] = r % (size_t)n[i]; r /= (size_t)n[i]; }         \
			out[(size_t)b * (size_t)odist + (size_t)ostride * embed_offset(rank, onembed, idx)] = (T)box[e]; \
		}                                                                                                 \
	}                                                                                                     \
	free(box);                                                                                            \
	return 0;                                                                                             \
}

/* Same argument meaning as fftw{f,,l}_plan_many_r2r + execute in one call (kind: 10 = REDFT10, 1 = REDFT01). */
DEFINE_MANY(ref_r2r_many_f, float)
DEFINE_MANY(ref_r2r_many_d, double)
DEFINE_MANY(ref_r2r_many_l, long double)

/* ---- pointwise helpers restated from the reference, used by the python pipelines for scalar spot checks ---- */

/* include/speclib.c:79-85 : sqrt(2)^n */
ld ref_spec_normalization(size_t n) {
	ld r = 1;
	for (size_t i = 0; i < n / 2; i++) r *= 2;
	return (n & 1) ? r * 1.41421356237309504880168872420969808L : r;
}

/* motion/motion.c:776 : pel > 255 ? 255 : pel < 0 ? 0 : lround(pel) */
unsigned char ref_quant_u8(ld pel) {
	return pel > 255 ? 255 : pel < 0 ? 0 : (unsigned char)lroundl(pel);
}

/* spec/spec.h:157-163 base16enc : low nibble first, alphabet 'A'+nibble */
void ref_base16enc(const unsigned char *in, char *out, size_t size) {
	for (size_t i = 0; i < size; i++) {
		*out++ = (char)((in[i] & 15) + 65);
		*out++ = (char)((in[i] >> 4) + 65);
	}
}
/* spec/spec.h:164-168 base16dec */
void ref_base16dec(const char *in, unsigned char *out, size_t size) {
	for (size_t i = 0; i < size; i++, in += 2)
		out[i] = (unsigned char)((in[0] - 65) | ((in[1] - 65) << 4));
}
