/*
 * tools/dct_cli.c -- headless C driver: plans and executes through the FFTW names (shim/fftw3.h -> libdspdct) with
 * exactly the argument lists of dspfun's call sites, on raw float buffers instead of MagickWand / FFmpeg frames.
 *
 *   dct_cli <site> <prec f|d> <kind 10|01> <n0> <n1> <n2> <d> <in.raw> <out.raw> [e0 e1 e2]
 *
 *   site = image   plan_many_r2r(2,{n0,n1},d, f,NULL,d,1, f,NULL,d,1, {k,k}, ESTIMATE)      spec.c:63 ispec.c:165 zoom.c:263 scan.c:292
 *          scan    same plan, out of place, FFTW_MEASURE                                    scan.c:359
 *          motion  plan_many_r2r(3,{n0,n1,n2},1, c,{e0,e1,e2},1,0, c,{e0,e1,e2},1,0, {k,k,k}) motion.c:535-552
 *          draw    plan_r2r_2d(n0,n1, c,c, k,k, ESTIMATE)                                   draw.c:74
 *
 * Built by tests/test_gpu_tools.py with: gcc -Ishim tools/dct_cli.c -Ldspfun_b200 -ldspdct
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <fftw3.h>

#define RUN(PFX, R)                                                                                                  \
	do {                                                                                                             \
		R *buf = PFX##_alloc_real(total), *out = buf;                                                                \
		if (!buf || fread(buf, sizeof(R), total, fi) != total) { fprintf(stderr, "short read\n"); return 1; }        \
		PFX##_plan p;                                                                                                \
		if (!strcmp(site, "image"))                                                                                  \
			p = PFX##_plan_many_r2r(2, (int[]){n0, n1}, d, buf, NULL, d, 1, buf, NULL, d, 1,                         \
			                        (fftw_r2r_kind[]){kind, kind}, FFTW_ESTIMATE);                                   \
		else if (!strcmp(site, "scan")) {                                                                            \
			out = PFX##_alloc_real(total);                                                                           \
			memset(out, 0, sizeof(R) * total);                                                                       \
			p = PFX##_plan_many_r2r(2, (int[2]){n0, n1}, d, buf, NULL, d, 1, out, NULL, d, 1,                        \
			                        (fftw_r2r_kind[2]){kind, kind}, FFTW_MEASURE);                                   \
		} else if (!strcmp(site, "motion"))                                                                          \
			p = PFX##_plan_many_r2r(3, (const int[3]){n0, n1, n2}, 1, buf, (const int[3]){e0, e1, e2}, 1, 0, buf,     \
			                        (const int[3]){e0, e1, e2}, 1, 0, (const fftw_r2r_kind[3]){kind, kind, kind},     \
			                        FFTW_ESTIMATE);                                                                  \
		else                                                                                                         \
			p = PFX##_plan_r2r_2d(n0, n1, buf, buf, kind, kind, FFTW_ESTIMATE);                                      \
		PFX##_execute(p);                                                                                            \
		PFX##_destroy_plan(p);                                                                                       \
		if (fwrite(out, sizeof(R), total, fo) != total) { fprintf(stderr, "short write\n"); return 1; }              \
		if (out != buf) PFX##_free(out);                                                                             \
		PFX##_free(buf);                                                                                             \
		PFX##_cleanup();                                                                                             \
	} while (0)

int main(int argc, char **argv) {
	if (argc < 10) {
		fprintf(stderr, "usage: %s image|scan|motion|draw f|d 10|01 n0 n1 n2 d in.raw out.raw [e0 e1 e2]\n", argv[0]);
		return 2;
	}
	const char *site = argv[1];
	const char prec = argv[2][0];
	const fftw_r2r_kind kind = !strcmp(argv[3], "10") ? FFTW_REDFT10 : FFTW_REDFT01;
	const int n0 = atoi(argv[4]), n1 = atoi(argv[5]), n2 = atoi(argv[6]), d = atoi(argv[7]);
	const int e0 = argc > 12 ? atoi(argv[10]) : n0, e1 = argc > 12 ? atoi(argv[11]) : n1, e2 = argc > 12 ? atoi(argv[12]) : n2;
	size_t total;
	if (!strcmp(site, "motion")) total = (size_t)e0 * e1 * e2;
	else total = (size_t)n0 * n1 * d;
	FILE *fi = fopen(argv[8], "rb"), *fo = fopen(argv[9], "wb");
	if (!fi || !fo) { perror("open"); return 1; }
	fftwf_init_threads();
	fftwf_plan_with_nthreads(4);
	if (prec == 'f') RUN(fftwf, float);
	else RUN(fftw, double);
	fftwf_cleanup_threads();
	fclose(fi);
	fclose(fo);
	return 0;
}
