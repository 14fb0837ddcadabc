// kern_zoom.cu -- zoom's scaled-basis generation (zoom/zoom.c:36-68) and separable cosine synthesis
// (zoom/zoom.c:361-375) as two dense contractions per channel:  out = Yb * C * Xb^T / (W H).
//
// This is the reference's O(N^3) loop restated as GEMMs.  Products of coeff-precision values are accumulated in
// double (the reference accumulates in `intermediate`), on the FP32/FP64 pipes: a first, correctness-oriented
// version -- a split-precision tensor-core path is the planned replacement.
#include "dsp_kernels.h"
#include <math.h>
#include <vector>

namespace dsp {

enum { ZOOM_INTERPOLATED = 0, ZOOM_CENTERED = 1, ZOOM_NATIVE = 2 };

// basis[b][0] = 1/2 (the halved DC terms of zoom.c:364,370) ; basis[b][n] = cos(pi (k_b + 1/2) n / N'), n >= 1
template <class T>
DSP_DEV void zoom_basis_entry(T *basis, int b, int n, int ncomp, int type, double num, double den, double offset, int len) {
	const double PI = 3.14159265358979323846264338327950288;
	double k, N;
	if (type == ZOOM_NATIVE) { k = b + offset; N = len * num / den; }                                  // zoom.c:50-53
	else if (type == ZOOM_INTERPOLATED) { k = (b + offset) * den / num; N = len; }                     // zoom.c:54-57
	else { k = (b + offset) * (len - 1) * den / (len * num - den); N = len; }                          // zoom.c:58-61
	basis[(size_t)b * ncomp + n] = n == 0 ? (T)0.5 : (T)cos(PI * (k + 0.5) * n / N);                   // zoom.c:63
}

// C[m][n] = alpha * sum_k A[m][k] * B[k][n], element strides for every operand (so transposes and the channel
// interleave are free), double accumulation.  64x64 output tile per CTA, 16-deep k slabs, 4x4 outputs per thread.
template <class T>
DSP_DEV void zoom_gemm_cta(int M, int N, int K, const T *A, long long ar, long long ac, const T *B, long long br, long long bc,
                           T *Cm, long long cr, long long cc, double alpha, int bm, int bn, int t0, int t1, T *sA, T *sB,
                           double *acc_all) {
	const int TM = 64, TN = 64, TK = 16;
	for (int tid = t0; tid < t1; tid++) {
		double *acc = acc_all + (size_t)(tid - t0) * 16 * (t1 - t0 > 1 ? 1 : 0);
		for (int i = 0; i < 16; i++) acc[i] = 0;
	}
	for (int k0 = 0; k0 < K; k0 += TK) {
		for (int tid = t0; tid < t1; tid++) {
			for (int e = tid; e < TM * TK; e += 256) {
				const int m = e / TK, k = e % TK;
				const int gm = bm * TM + m, gk = k0 + k;
				sA[k * (TM + 1) + m] = (gm < M && gk < K) ? A[gm * ar + gk * ac] : (T)0;
			}
			for (int e = tid; e < TK * TN; e += 256) {
				const int k = e / TN, n = e % TN;
				const int gk = k0 + k, gn = bn * TN + n;
				sB[k * (TN + 1) + n] = (gk < K && gn < N) ? B[gk * br + gn * bc] : (T)0;
			}
		}
		DSP_SYNC();
		for (int tid = t0; tid < t1; tid++) {
			double *acc = acc_all + (size_t)(tid - t0) * 16 * (t1 - t0 > 1 ? 1 : 0);
			const int tm = (tid / 16) * 4, tn = (tid % 16) * 4;
			for (int k = 0; k < TK; k++) {
				double a[4], b[4];
				for (int i = 0; i < 4; i++) { a[i] = (double)sA[k * (TM + 1) + tm + i]; b[i] = (double)sB[k * (TN + 1) + tn + i]; }
				for (int i = 0; i < 4; i++)
					for (int j = 0; j < 4; j++) acc[i * 4 + j] += a[i] * b[j];
			}
		}
		DSP_SYNC();
	}
	for (int tid = t0; tid < t1; tid++) {
		const double *acc = acc_all + (size_t)(tid - t0) * 16 * (t1 - t0 > 1 ? 1 : 0);
		const int tm = (tid / 16) * 4, tn = (tid % 16) * 4;
		for (int i = 0; i < 4; i++)
			for (int j = 0; j < 4; j++) {
				const int gm = bm * TM + tm + i, gn = bn * TN + tn + j;
				if (gm < M && gn < N) Cm[gm * cr + gn * cc] = (T)(alpha * acc[i * 4 + j]);
			}
	}
}

#if DSP_GPU
template <class T>
__global__ void k_zoom_basis(T *basis, int nvec, int ncomp, int type, double num, double den, double offset, int len) {
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < (long long)nvec * ncomp) zoom_basis_entry<T>(basis, (int)(i / ncomp), (int)(i % ncomp), ncomp, type, num, den, offset, len);
}
template <class T>
__global__ void __launch_bounds__(256) k_zoom_gemm(int M, int N, int K, const T *A, long long ar, long long ac, const T *B,
                                                   long long br, long long bc, T *Cm, long long cr, long long cc, double alpha) {
	__shared__ T sA[16 * 65], sB[16 * 65];
	double acc[16];
	zoom_gemm_cta<T>(M, N, K, A, ar, ac, B, br, bc, Cm, cr, cc, alpha, (int)blockIdx.y, (int)blockIdx.x, (int)threadIdx.x,
	                 (int)threadIdx.x + 1, sA, sB, acc);
}
#endif

template <class T>
static bool zoom_basis_t(void *basis, int nvec, int ncomp, int type, double num, double den, double offset, int len, rt_stream st,
                         std::string &err) {
#if DSP_GPU
	const long long total = (long long)nvec * ncomp;
	k_zoom_basis<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>((T *)basis, nvec, ncomp, type, num, den, offset, len);
	return rt_ok(cudaGetLastError(), err, "zoom basis launch");
#else
	(void)st; (void)err;
	for (int b = 0; b < nvec; b++)
		for (int n = 0; n < ncomp; n++) zoom_basis_entry<T>((T *)basis, b, n, ncomp, type, num, den, offset, len);
	return true;
#endif
}

template <class T>
static bool zoom_gemm_t(int M, int N, int K, const void *A, long long ar, long long ac, const void *B, long long br, long long bc,
                        void *Cm, long long cr, long long cc, double alpha, rt_stream st, std::string &err) {
#if DSP_GPU
	dim3 grid((unsigned)((N + 63) / 64), (unsigned)((M + 63) / 64));
	k_zoom_gemm<T><<<grid, 256, 0, st>>>(M, N, K, (const T *)A, ar, ac, (const T *)B, br, bc, (T *)Cm, cr, cc, alpha);
	return rt_ok(cudaGetLastError(), err, "zoom synthesis launch");
#else
	(void)st; (void)err;
	std::vector<T> sA(16 * 65), sB(16 * 65);
	std::vector<double> acc(256 * 16);
	for (int bm = 0; bm < (M + 63) / 64; bm++)
		for (int bn = 0; bn < (N + 63) / 64; bn++)
			zoom_gemm_cta<T>(M, N, K, (const T *)A, ar, ac, (const T *)B, br, bc, (T *)Cm, cr, cc, alpha, bm, bn, 0, 256, sA.data(),
			                 sB.data(), acc.data());
	return true;
#endif
}

bool launch_zoom_basis(char prec, void *basis, int nvec, int ncomp, int type, double num, double den, double offset, int len,
                       rt_stream st, std::string &err) {
	return prec == 'f' ? zoom_basis_t<float>(basis, nvec, ncomp, type, num, den, offset, len, st, err)
	                   : zoom_basis_t<double>(basis, nvec, ncomp, type, num, den, offset, len, st, err);
}

bool launch_zoom_gemm(char prec, int M, int N, int K, const void *A, long long ar, long long ac, const void *B, long long br,
                      long long bc, void *Cm, long long cr, long long cc, double alpha, rt_stream st, std::string &err) {
	return prec == 'f' ? zoom_gemm_t<float>(M, N, K, A, ar, ac, B, br, bc, Cm, cr, cc, alpha, st, err)
	                   : zoom_gemm_t<double>(M, N, K, A, ar, ac, B, br, bc, Cm, cr, cc, alpha, st, err);
}

}  // namespace dsp
