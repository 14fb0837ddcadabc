// kern_zoom.cu -- zoom's scaled-basis generation (zoom/zoom.c:36-68) and separable cosine synthesis
// (zoom/zoom.c:361-375) as two dense contractions per channel:  out = Yb * C * Xb^T / (W H).
//
// This is the reference's O(N^3) loop restated as GEMMs.  Products of coeff-precision values are accumulated in
// double (the reference accumulates in `intermediate`), on the FP32/FP64 pipes.  Float sessions take the split-precision
// tensor-core GEMM of kern_gemm_tc.cu instead (dsp_zoom_frame); this one serves double and is the float fallback.
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

// ------------------------------------------------------------------------------------------------ shifted-DCT path
// Interpolated basis with an integer scaled size N' = N s: the sample angle is pi u (i + 1/2 + delta) / N' with
// delta = offset + (s - 1)/2, and cos(alpha + phi) = cos alpha cos phi - sin alpha sin phi turns the synthesis into
// REDFT01s of length N':   sum_u C_u cos(.) = 1/2 [ T(A)[i] - (-1)^i T(G)[i] ],
//   A_u = C_u cos(phi_u),  G_m = C_{N'-m} sin(phi_{N'-m})  (sin(pi (N'-m)(i+1/2)/N') = (-1)^i cos(pi m (i+1/2)/N')),
// phi_u = pi u delta / N'.  In two dimensions: four planes (AA, AG, GA, GG), four REDFT01 x REDFT01, one combine.
DSP_DEV void zoom_sincospi(double x, double &s, double &c) {
#if DSP_GPU
	sincospi(x, &s, &c);
#else
	const double r = x - 2.0 * floor(x / 2.0);
	s = sin(3.14159265358979323846264338327950288 * r);
	c = cos(3.14159265358979323846264338327950288 * r);
#endif
}

// element (v, u, ch) of the coefficient plane -> its four destinations
template <class T>
DSP_DEV void zoom_shift_build_elem(const T *coef, T *planes, int W, int Nh, int Nw, int cw, double dx, double dy, long long idx) {
	const int ch = (int)(idx % 3);
	const long long vu = idx / 3;
	const int u = (int)(vu % cw), v = (int)(vu / cw);
	const double C = (double)coef[((size_t)v * W + u) * 3 + ch];
	double sx = 0, cx = 1, sy = 0, cy = 1;
	if (u) zoom_sincospi((double)u * dx / (double)Nw, sx, cx);
	if (v) zoom_sincospi((double)v * dy / (double)Nh, sy, cy);
	const size_t plane = (size_t)Nh * Nw * 3;
	planes[((size_t)v * Nw + u) * 3 + ch] = (T)(C * cy * cx);                                        // AA
	if (u) planes[plane + ((size_t)v * Nw + (Nw - u)) * 3 + ch] = (T)(C * cy * sx);                  // AG
	if (v) planes[2 * plane + ((size_t)(Nh - v) * Nw + u) * 3 + ch] = (T)(C * sy * cx);              // GA
	if (u && v) planes[3 * plane + ((size_t)(Nh - v) * Nw + (Nw - u)) * 3 + ch] = (T)(C * sy * sx);  // GG
}

template <class T>
DSP_DEV void zoom_shift_combine_elem(const T *planes, T *out, int Nh, int Nw, int vw, double alpha, long long idx) {
	const int ch = (int)(idx % 3);
	const long long ji = idx / 3;
	const int i = (int)(ji % vw), j = (int)(ji / vw);
	const size_t plane = (size_t)Nh * Nw * 3, at = ((size_t)j * Nw + i) * 3 + ch;
	const double si = (i & 1) ? -1.0 : 1.0, sj = (j & 1) ? -1.0 : 1.0;
	const double r = (double)planes[at] - si * (double)planes[plane + at] - sj * (double)planes[2 * plane + at] +
	                 si * sj * (double)planes[3 * plane + at];
	out[idx] = (T)(alpha * r);
}

#if DSP_GPU
template <class T>
__global__ void k_zoom_shift_build(const T *coef, T *planes, int W, int Nh, int Nw, int cw, double dx, double dy, long long total) {
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
		zoom_shift_build_elem<T>(coef, planes, W, Nh, Nw, cw, dx, dy, i);
}
template <class T>
__global__ void k_zoom_shift_combine(const T *planes, T *out, int Nh, int Nw, int vw, double alpha, long long total) {
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
		zoom_shift_combine_elem<T>(planes, out, Nh, Nw, vw, alpha, i);
}
#endif

template <class T>
static bool zoom_shift_build_t(const void *coef, void *planes, int W, int Nh, int Nw, int ch, int cw, double dx, double dy, rt_stream st, std::string &err) {
	const long long total = (long long)ch * cw * 3;
#if DSP_GPU
	k_zoom_shift_build<T><<<148 * 8, 256, 0, st>>>((const T *)coef, (T *)planes, W, Nh, Nw, cw, dx, dy, total);
	return rt_ok(cudaGetLastError(), err, "zoom shift build launch");
#else
	(void)st; (void)err;
	for (long long i = 0; i < total; i++) zoom_shift_build_elem<T>((const T *)coef, (T *)planes, W, Nh, Nw, cw, dx, dy, i);
	return true;
#endif
}
template <class T>
static bool zoom_shift_combine_t(const void *planes, void *out, int Nh, int Nw, int vh, int vw, double alpha, rt_stream st, std::string &err) {
	const long long total = (long long)vh * vw * 3;
#if DSP_GPU
	k_zoom_shift_combine<T><<<148 * 8, 256, 0, st>>>((const T *)planes, (T *)out, Nh, Nw, vw, alpha, total);
	return rt_ok(cudaGetLastError(), err, "zoom shift combine launch");
#else
	(void)st; (void)err;
	for (long long i = 0; i < total; i++) zoom_shift_combine_elem<T>((const T *)planes, (T *)out, Nh, Nw, vw, alpha, i);
	return true;
#endif
}
bool launch_zoom_shift_build(char prec, const void *coef, void *planes, int W, int Nh, int Nw, int ch, int cw, double dx, double dy,
                             rt_stream st, std::string &err) {
	return prec == 'f' ? zoom_shift_build_t<float>(coef, planes, W, Nh, Nw, ch, cw, dx, dy, st, err)
	                   : zoom_shift_build_t<double>(coef, planes, W, Nh, Nw, ch, cw, dx, dy, st, err);
}
bool launch_zoom_shift_combine(char prec, const void *planes, void *out, int Nh, int Nw, int vh, int vw, double alpha, rt_stream st,
                               std::string &err) {
	return prec == 'f' ? zoom_shift_combine_t<float>(planes, out, Nh, Nw, vh, vw, alpha, st, err)
	                   : zoom_shift_combine_t<double>(planes, out, Nh, Nw, vh, vw, alpha, st, err);
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
