// kern_misc.cu -- small helper kernels
#include "dsp_kernels.h"
#include <math.h>

namespace dsp {

#if DSP_GPU
template <class T> __global__ void k_spec_resolve(OpAny op, const double *acc, double *scale_z) {
	if (threadIdx.x == 0 && blockIdx.x == 0) spec_resolve_range<T>(op, acc, scale_z);
}
#endif

#if DSP_GPU
// pulls a [nrows][row_bytes] region (row pitch `pitch` bytes) into L2, one prefetch per 128-byte line; the kernel
// only issues the prefetches, the transfers complete while the following kernels of the stream run
__global__ void k_l2_prefetch(const char *base, long long pitch, int nrows, int lines_per_row) {
	const long long total = (long long)nrows * lines_per_row;
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
		const long long r = i / lines_per_row;
		const int l = (int)(i - r * lines_per_row);
		const char *p = base + r * pitch + (long long)l * 128;
		asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
	}
}
#endif

bool launch_l2_prefetch(const void *base, long long pitch_bytes, int nrows, int row_bytes, rt_stream st, std::string &err) {
#if DSP_GPU
	if (nrows <= 0 || row_bytes <= 0) return true;
	const int lpr = (row_bytes + 127) / 128;
	const long long total = (long long)nrows * lpr;
	int grid = (int)((total + 255) / 256);
	if (grid > 148 * 8) grid = 148 * 8;
	k_l2_prefetch<<<grid, 256, 0, st>>>((const char *)base, pitch_bytes, nrows, lpr);
	return rt_ok(cudaGetLastError(), err, "L2 prefetch launch");
#else
	(void)base; (void)pitch_bytes; (void)nrows; (void)row_bytes; (void)st; (void)err;
	return true;
#endif
}

bool launch_spec_resolve(char prec, const OpAny &op, const double *acc, double *scale_z, rt_stream st, std::string &err) {
#if DSP_GPU
	if (prec == 'f') k_spec_resolve<float><<<1, 32, 0, st>>>(op, acc, scale_z);
	else k_spec_resolve<double><<<1, 32, 0, st>>>(op, acc, scale_z);
	return rt_ok(cudaGetLastError(), err, "spec resolve launch");
#else
	(void)st; (void)err;
	if (prec == 'f') spec_resolve_range<float>(op, acc, scale_z);
	else spec_resolve_range<double>(op, acc, scale_z);
	return true;
#endif
}

// ------------------------------------------------------------------------------------------------ block-tiled motion
// The coefficient stage of `motion -b BWxBHxBD --quant` over a whole block-tiled volume (motion/motion.c:644-647
// normalise, :740-744 quantise + count, :748-751 de-normalise), in place, and the 8-bit store (:757-776).
// One pass each; the in-block indices come from the flat element index.
struct BlockQuantArgs {
	long long n;                 // D * H * W
	int H, W, bd, bh, bw;
	double nf[4];                // 2 sqrt2 / sqrt2^k for k = number of zero in-block indices
	double rnf[4], rquant;       // reciprocals (dsp_block_dquant)
	double quantizer;            // 0 = no quantisation (the normalise / de-normalise rounding still applies)
	unsigned long long *count;   // device counter of non-zero quantised coefficients (may be null)
};

template <class T>
DSP_DEV int block_quant_elem(const BlockQuantArgs &a, T *c, long long i) {
	const long long row = i / a.W;
	const int x = (int)(i - row * a.W);
	const long long zz = row / a.H;
	const int y = (int)(row - zz * a.H);
	const int k = ((x % a.bw) == 0) + ((y % a.bh) == 0) + (((int)(zz % a.bd)) == 0);
	const double nf = a.nf[k];
	T f = (T)((double)c[i] * nf);
	int nz = 0;
	if (a.quantizer != 0.0) {
		f = (T)(round((double)f / a.quantizer) * a.quantizer);
		nz = f != 0;
	}
	c[i] = (T)((double)f / nf);
	return nz;
}

template <class T>
DSP_DEV unsigned char block_store_elem(T c, double scale) {
	const double pel = (double)c * scale;
	return (unsigned char)(pel > 255.0 ? 255.0 : pel < 0.0 ? 0.0 : (double)lround(pel));
}

#if DSP_GPU
template <class T> __global__ void k_block_quant(BlockQuantArgs a, T *c) {
	unsigned long long local = 0;
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x)
		local += (unsigned long long)block_quant_elem<T>(a, c, i);
	if (a.count) {
		for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
		if ((threadIdx.x & 31) == 0 && local) atomicAdd(a.count, local);
	}
}
template <class T> __global__ void k_block_store_u8(const T *c, unsigned char *pels, long long n, double scale) {
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
		pels[i] = block_store_elem<T>(c[i], scale);
}
#endif

// ------------------------------------------------------------------------------------------------ d axis + quantiser in one pass
// The blocks of a tiled volume are independent along d as well: once the two spatial axes are done (dsp_block_dct2d), one
// thread per pixel and group of BD frames holds the BD samples in registers, applies REDFT10 along d (a BD x BD matrix
// product with FFTW's unnormalised cosines), the coefficient stage above on the now complete 3-D coefficients, and REDFT01
// along d -- three sweeps over the volume (d forward, quantise, d inverse = 24 B per sample) become one (8 B per sample).
// Accesses are coalesced along (h, w).
template <int BD>
DSP_DEV int block_dquant_pixel(const BlockQuantArgs &a, float *c, long long hw, long long plane, long long zg, const float *m10, const float *m01) {
	const long long row = hw / a.W;
	const int x = (int)(hw - row * a.W), y = (int)row;
	const int kxy = ((x % a.bw) == 0) + ((y % a.bh) == 0);
	float v[BD], t[BD];
	float *p = c + zg * BD * plane + hw;
#pragma unroll
	for (int z = 0; z < BD; z++) v[z] = p[z * plane];
	int nz = 0;
#pragma unroll
	for (int k = 0; k < BD; k++) {                                   // REDFT10 along d, then motion.c:644-647, 740-751
		float s = 0.0f;
#pragma unroll
		for (int z = 0; z < BD; z++) s = fmaf(m10[k * BD + z], v[z], s);
		// (the two divisions of the reference are multiplications by reciprocals prepared once, as in the spec store stage:
		// the double result differs by at most one ulp, which the cast to float absorbs)
		const double nf = a.nf[kxy + (k == 0)], rnf = a.rnf[kxy + (k == 0)];
		float f = (float)((double)s * nf);
		if (a.quantizer != 0.0) {
			f = (float)(round((double)f * a.rquant) * a.quantizer);
			nz += f != 0.0f;
		}
		t[k] = (float)((double)f * rnf);
	}
#pragma unroll
	for (int z = 0; z < BD; z++) {                                   // REDFT01 along d
		float s = 0.0f;
#pragma unroll
		for (int k = 0; k < BD; k++) s = fmaf(m01[z * BD + k], t[k], s);
		p[z * plane] = s;
	}
	return nz;
}
#if DSP_GPU
template <int BD> __global__ void __launch_bounds__(256) k_block_dquant(BlockQuantArgs a, float *c, long long plane, long long groups) {
	__shared__ float m10[BD * BD], m01[BD * BD];
	const double PI = 3.14159265358979323846264338327950288;
	for (int i = threadIdx.x; i < BD * BD; i += blockDim.x) {
		const int r = i / BD, q = i - r * BD;
		m10[i] = (float)(2.0 * cos(PI * (q + 0.5) * r / BD));                       // Y[r] = 2 sum_q x[q] cos(pi (q + 1/2) r / BD)
		m01[i] = q == 0 ? 1.0f : (float)(2.0 * cos(PI * (r + 0.5) * q / BD));       // y[r] = x[0] + 2 sum_{q>=1} x[q] cos(pi (r + 1/2) q / BD)
	}
	__syncthreads();
	unsigned long long local = 0;
	const long long total = plane * groups;
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
		const long long zg = i / plane, hw = i - zg * plane;
		local += (unsigned long long)block_dquant_pixel<BD>(a, c, hw, plane, zg, m10, m01);
	}
	if (a.count) {
		for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
		if ((threadIdx.x & 31) == 0 && local) atomicAdd(a.count, local);
	}
}
#endif
template <int BD>
static bool block_dquant_t(const BlockQuantArgs &a, float *c, long long plane, long long groups, rt_stream st, std::string &err) {
#if DSP_GPU
	k_block_dquant<BD><<<148 * 8, 256, 0, st>>>(a, c, plane, groups);
	return rt_ok(cudaGetLastError(), err, "block d-axis + quantiser launch");
#else
	(void)st; (void)err;
	const double PI = 3.14159265358979323846264338327950288;
	float m10[BD * BD], m01[BD * BD];
	for (int i = 0; i < BD * BD; i++) {
		const int r = i / BD, q = i - r * BD;
		m10[i] = (float)(2.0 * cos(PI * (q + 0.5) * r / BD));
		m01[i] = q == 0 ? 1.0f : (float)(2.0 * cos(PI * (r + 0.5) * q / BD));
	}
	unsigned long long total = 0;
	for (long long zg = 0; zg < groups; zg++)
		for (long long hw = 0; hw < plane; hw++) total += (unsigned long long)block_dquant_pixel<BD>(a, c, hw, plane, zg, m10, m01);
	if (a.count) *a.count += total;
	return true;
#endif
}
bool block_dquant_supports(int bd) { return bd == 2 || bd == 4 || bd == 8 || bd == 16; }
bool launch_block_dquant(float *coeffs, int D, int H, int W, int bd, int bh, int bw, double quantizer, unsigned long long *count, rt_stream st,
                         std::string &err) {
	BlockQuantArgs a;
	a.n = (long long)D * H * W; a.H = H; a.W = W; a.bd = bd; a.bh = bh; a.bw = bw; a.quantizer = quantizer; a.count = count;
	const long double s2 = 1.41421356237309504880168872420969808L;
	long double den = 1.0L;
	for (int k = 0; k < 4; k++) { a.nf[k] = (double)((2.0L * s2) / den); a.rnf[k] = (double)(den / (2.0L * s2)); den *= s2; }
	a.rquant = quantizer != 0.0 ? 1.0 / quantizer : 0.0;
	const long long plane = (long long)H * W, groups = D / bd;
	switch (bd) {
	case 2: return block_dquant_t<2>(a, coeffs, plane, groups, st, err);
	case 4: return block_dquant_t<4>(a, coeffs, plane, groups, st, err);
	case 8: return block_dquant_t<8>(a, coeffs, plane, groups, st, err);
	case 16: return block_dquant_t<16>(a, coeffs, plane, groups, st, err);
	default: break;
	}
	err = "block d-axis + quantiser: depth must be 2, 4, 8 or 16";
	return false;
}

bool launch_block_quant(char prec, void *coeffs, long long n, int H, int W, int bd, int bh, int bw, double quantizer,
                        unsigned long long *count, rt_stream st, std::string &err) {
	BlockQuantArgs a;
	a.n = n; a.H = H; a.W = W; a.bd = bd; a.bh = bh; a.bw = bw; a.quantizer = quantizer; a.count = count;
	const long double s2 = 1.41421356237309504880168872420969808L;
	long double den = 1.0L;
	for (int k = 0; k < 4; k++) { a.nf[k] = (double)((2.0L * s2) / den); den *= s2; }
#if DSP_GPU
	const int grid = 148 * 16;
	if (prec == 'f') k_block_quant<float><<<grid, 256, 0, st>>>(a, (float *)coeffs);
	else k_block_quant<double><<<grid, 256, 0, st>>>(a, (double *)coeffs);
	return rt_ok(cudaGetLastError(), err, "block quant launch");
#else
	(void)st; (void)err;
	unsigned long long total = 0;
	for (long long i = 0; i < n; i++)
		total += (unsigned long long)(prec == 'f' ? block_quant_elem<float>(a, (float *)coeffs, i) : block_quant_elem<double>(a, (double *)coeffs, i));
	if (count) *count += total;
	return true;
#endif
}

// 8-bit pels -> coefficients-to-be (motion.c:618-624 for 8-bit input: the value itself)
#if DSP_GPU
template <class T> __global__ void k_block_load_u8(const unsigned char *pels, T *c, long long n) {
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) c[i] = (T)pels[i];
}
#endif
bool launch_block_load_u8(char prec, const unsigned char *pels, void *coeffs, long long n, rt_stream st, std::string &err) {
#if DSP_GPU
	const int grid = 148 * 16;
	if (prec == 'f') k_block_load_u8<float><<<grid, 256, 0, st>>>(pels, (float *)coeffs, n);
	else k_block_load_u8<double><<<grid, 256, 0, st>>>(pels, (double *)coeffs, n);
	return rt_ok(cudaGetLastError(), err, "block load launch");
#else
	(void)st; (void)err;
	for (long long i = 0; i < n; i++) {
		if (prec == 'f') ((float *)coeffs)[i] = (float)pels[i];
		else ((double *)coeffs)[i] = (double)pels[i];
	}
	return true;
#endif
}

bool launch_block_store_u8(char prec, const void *coeffs, unsigned char *pels, long long n, double scale, rt_stream st, std::string &err) {
#if DSP_GPU
	const int grid = 148 * 16;
	if (prec == 'f') k_block_store_u8<float><<<grid, 256, 0, st>>>((const float *)coeffs, pels, n, scale);
	else k_block_store_u8<double><<<grid, 256, 0, st>>>((const double *)coeffs, pels, n, scale);
	return rt_ok(cudaGetLastError(), err, "block store launch");
#else
	(void)st; (void)err;
	for (long long i = 0; i < n; i++)
		pels[i] = prec == 'f' ? block_store_elem<float>(((const float *)coeffs)[i], scale) : block_store_elem<double>(((const double *)coeffs)[i], scale);
	return true;
#endif
}

// ------------------------------------------------------------------------------------------------ motion coefficient stage, standalone
// The same pointwise map the plans can carry in a pass (OP_MOTION_COEFF), as one lean sweep over a [D][H][W] volume of
// coefficients.  Fused into the temporal pass of the 256 x 1080 x 1920 volume it costs 2.4 ms on top of the plain pass (the
// coordinate-carrying tile moves of the transform kernels are slow); this sweep moves 8 B per sample at stream speed.
// H > 0: a [D][H][W] volume, coordinates (z, y, x).  H == 0: a [D][W] array whose W columns are a slice of the flattened
// (y, x) positions (the temporal layout of a slab-sharded volume): the stage decodes (y, x) from op.lo + column itself.
template <class T, class Op>
DSP_DEV void motion_coeff_elem(const Op &op, T *c, uint32_t i, uint32_t W, uint32_t H, const FastDiv &dW, const FastDiv &dH) {
	const uint32_t row = fd_div(i, dW), x = i - row * W;
	Coord cc = {0, 0, 0, 0, 0};
	if (H == 0) { cc.ch = (int)x; cc.i2 = (int)row; }
	else {
		const uint32_t z = fd_div(row, dH), y = row - z * H;
		cc.i0 = (int)z; cc.i1 = (int)y; cc.i2 = (int)x;
	}
	c[i] = op(c[i], cc);
}
#if DSP_GPU
template <class T, class Op>
__global__ void __launch_bounds__(256) k_motion_coeff(const __grid_constant__ Op op, T *c, uint32_t n, uint32_t W, uint32_t H, FastDiv dW, FastDiv dH) {
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) motion_coeff_elem<T, Op>(op, c, i, W, H, dW, dH);
}
#endif
static FastDiv misc_mk_fd(uint32_t d) {           // same scheme as the planner's (fd_div: n < 2^31)
	FastDiv f;
	f.d = d ? d : 1;
	if (f.d == 1) { f.mul = 0; f.shr = 0; return f; }
	uint32_t k = 0;
	while ((1ull << k) < f.d) k++;
	const uint32_t p = 31 + k;
	f.mul = (uint32_t)(((1ull << p) + f.d - 1) / f.d);
	f.shr = p - 32;
	return f;
}
template <class T>
static bool motion_coeff_t(const OpAny &op, T *c, int D, int H, int W, const FastDiv &dW, const FastDiv &dH, rt_stream st, std::string &err) {
	const long long n = (long long)D * (H > 0 ? H : 1) * W;
	if (n >= (1ll << 31)) { err = "coefficient stage: volume too large for one sweep"; return false; }
#if DSP_GPU
	const int grid = 148 * 16;
	if (op.fast) k_motion_coeff<T, OpMotionCoeff><<<grid, 256, 0, st>>>(OpMotionCoeff::from(op), c, (uint32_t)n, (uint32_t)W, (uint32_t)H, dW, dH);
	else k_motion_coeff<T, OpAny><<<grid, 256, 0, st>>>(op, c, (uint32_t)n, (uint32_t)W, (uint32_t)H, dW, dH);
	return rt_ok(cudaGetLastError(), err, "motion coefficient stage launch");
#else
	(void)st;
	for (long long i = 0; i < n; i++) motion_coeff_elem<T, OpAny>(op, c, (uint32_t)i, (uint32_t)W, (uint32_t)H, dW, dH);
	return true;
#endif
}
bool launch_motion_coeff(char prec, const OpAny &op, void *coeffs, int D, int H, int W, rt_stream st, std::string &err) {
	const FastDiv dW = misc_mk_fd((uint32_t)W), dH = misc_mk_fd((uint32_t)(H > 0 ? H : 1));
	return prec == 'f' ? motion_coeff_t<float>(op, (float *)coeffs, D, H, W, dW, dH, st, err) : motion_coeff_t<double>(op, (double *)coeffs, D, H, W, dW, dH, st, err);
}

// ------------------------------------------------------------------------------------------------ motion spectrograms
// motion --ispec: the pels ARE a spectrogram; they become coefficients without a forward transform (motion.c:627-637).
// motion --spec : the coefficients are written out as a spectrogram instead of being inverted (motion.c:755-776).
// Both are pointwise over the block's padded box [md][mh][mw]; layouts and rounding points as in the reference.
template <class T>
DSP_DEV T motion_ispec_elem(const MotionSpecArgs &a, double pel) {
	typedef double I;
	switch (a.type) {
	case 2: pel = copysign(expm1(fabs((pel - 127.5) / a.c)), pel - 127.5) / a.norm; break;            // shift  :628
	case 3: pel = (pel - 127.5) * 2 / a.norm / a.norm; break;                                         // flat   :629
	case 4: pel = pel / a.norm / a.norm; break;                                                       // copy   :630
	default: break;
	}
	return (T)(I)pel;
}

template <class T>
DSP_DEV double motion_spec_elem(const MotionSpecArgs &a, T coeff, double cabs) {
	double pel = (double)coeff * a.sf * a.norm;                                                       // :757
	switch (a.type) {
	case 1: pel = cabs * log1p(fabs(pel)); break;                                                     // abs    :760
	case 2: pel = a.c * copysign(log1p(fabs(pel)), pel) + 127.5; break;                               // shift  :761
	case 3: pel = pel * a.norm / 2 + 127.5; break;                                                    // flat   :762
	default: pel *= a.norm; break;                                                                    // copy   :764-765
	}
	return pel;
}

template <class T>
DSP_DEV void motion_ispec_at(const MotionSpecArgs &a, const void *pels, T *coeffs, long long i) {
	const long long row = i / a.mw;
	const int x = (int)(i - row * a.mw), y = (int)(row % a.mh), z = (int)(row / a.mh);
	if (z >= a.bd || y >= a.bh || x >= a.bw) { coeffs[i] = (T)0; return; }                            // :617 memset
	const double pel = a.float_pixels ? (double)((const float *)pels)[i] * 255 : (double)((const unsigned char *)pels)[i];   // :621-624
	coeffs[i] = motion_ispec_elem<T>(a, pel);
}

template <class T>
DSP_DEV void motion_spec_at(const MotionSpecArgs &a, const T *coeffs, void *pels, long long i, double cabs) {
	const long long row = i / a.mw;
	const int x = (int)(i - row * a.mw), y = (int)(row % a.mh), z = (int)(row / a.mh);
	if (z >= a.bd || y >= a.bh || x >= a.bw) return;                                                  // outside the scaled box: pels stay
	const Coord c = {z, y, x, 0, 0};
	const T f = a.coeff.template operator()<T>(coeffs[i], c);                                         // zero outside the active box, filters, quantiser
	const double pel = motion_spec_elem<T>(a, f, cabs);
	if (a.float_pixels) ((float *)pels)[i] = (float)(pel / 255);                                      // :773
	else ((unsigned char *)pels)[i] = (unsigned char)(pel > 255.0 ? 255.0 : pel < 0.0 ? 0.0 : (double)lround(pel));   // :776
}

// c of --spec abs: 255 / log1p(|dc sf norm|), dc = the block's (normalised, unfiltered) DC coefficient (:649, :754)
template <class T>
DSP_DEV double motion_spec_cabs(const MotionSpecArgs &a, const T *coeffs) {
	if (a.type != 1) return 0.0;
	T dc = coeffs[0];
	if (!a.coeff.skipn) dc = (T)((double)dc * a.coeff.nf[3]);                                         // :644-647 at x = y = z = 0
	return 255.0 / log1p(fabs((double)dc * a.sf * a.norm));
}

#if DSP_GPU
template <class T> __global__ void k_motion_ispec(MotionSpecArgs a, const void *pels, T *coeffs) {
	const long long n = (long long)a.md * a.mh * a.mw;
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) motion_ispec_at<T>(a, pels, coeffs, i);
}
template <class T> __global__ void k_motion_spec(MotionSpecArgs a, const T *coeffs, void *pels) {
	const long long n = (long long)a.md * a.mh * a.mw;
	const double cabs = motion_spec_cabs<T>(a, coeffs);
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) motion_spec_at<T>(a, coeffs, pels, i, cabs);
}
#endif

bool launch_motion_ispec(char prec, const MotionSpecArgs &a, const void *pels, void *coeffs, rt_stream st, std::string &err) {
	const long long n = (long long)a.md * a.mh * a.mw;
#if DSP_GPU
	long long g = (n + 255) / 256;
	const int grid = (int)(g < 148 * 16 ? g : 148 * 16);
	if (prec == 'f') k_motion_ispec<float><<<grid, 256, 0, st>>>(a, pels, (float *)coeffs);
	else k_motion_ispec<double><<<grid, 256, 0, st>>>(a, pels, (double *)coeffs);
	return rt_ok(cudaGetLastError(), err, "motion ispec launch");
#else
	(void)st; (void)err;
	for (long long i = 0; i < n; i++) {
		if (prec == 'f') motion_ispec_at<float>(a, pels, (float *)coeffs, i); else motion_ispec_at<double>(a, pels, (double *)coeffs, i);
	}
	return true;
#endif
}

bool launch_motion_spec(char prec, const MotionSpecArgs &a, const void *coeffs, void *pels, rt_stream st, std::string &err) {
	const long long n = (long long)a.md * a.mh * a.mw;
#if DSP_GPU
	long long g = (n + 255) / 256;
	const int grid = (int)(g < 148 * 16 ? g : 148 * 16);
	if (prec == 'f') k_motion_spec<float><<<grid, 256, 0, st>>>(a, (const float *)coeffs, pels);
	else k_motion_spec<double><<<grid, 256, 0, st>>>(a, (const double *)coeffs, pels);
	return rt_ok(cudaGetLastError(), err, "motion spec launch");
#else
	(void)st; (void)err;
	const double cabs = prec == 'f' ? motion_spec_cabs<float>(a, (const float *)coeffs) : motion_spec_cabs<double>(a, (const double *)coeffs);
	for (long long i = 0; i < n; i++) {
		if (prec == 'f') motion_spec_at<float>(a, (const float *)coeffs, pels, i, cabs); else motion_spec_at<double>(a, (const double *)coeffs, pels, i, cabs);
	}
	return true;
#endif
}

}  // namespace dsp
