// dct_core.cuh -- the CTA-level DCT-II / DCT-III engine (REDFT10 / REDFT01) for sm_100a.
//
// One CTA transforms a batch of real sequences held in shared memory:
//   copy-in (128-bit global loads, fused prologue op, Makhoul even/odd permutation for DCT-II)
//   -> [DCT-III: pre-twiddle of (k, n-k) pairs, in place]
//   -> in-place mixed-radix DIF complex FFT (two real sequences ride in one complex sequence)
//   -> [DCT-II: post-twiddle of (k, n-k) pairs, in place at their digit-reversed slots]
//   -> copy-out (digit-reversed gather, fused epilogue op, 128-bit global stores).
//
// The same source is compiled two ways:
//   * by nvcc for sm_100a (the product: libdspdct.so), and
//   * by g++ with -DDSP_EMULATE as a sequential SIMT emulation used ONLY by tests/emu to check the index
//     arithmetic on machines without a GPU.  The emulation is never part of the product library.
//
// Every "phase" below is written so that a thread touches only smem slots it owns within the phase; phases are
// separated by DSP_SYNC().  That is what lets the emulation run threads one after another.
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__) && !defined(DSP_EMULATE)
#define DSP_GPU 1
#define DSP_DEV __device__ __forceinline__
#define DSP_DEVM __device__ __forceinline__
#define DSP_HDM __host__ __device__ __forceinline__
#define DSP_SYNC() __syncthreads()
#define DSP_LDG(p) __ldg(p)
#else
#define DSP_GPU 0
#define DSP_DEV static inline
#define DSP_DEVM inline
#define DSP_HDM inline
#define DSP_SYNC() ((void)0)
#define DSP_LDG(p) (*(p))
#endif

#define DSP_MAX_FAC 14
#define DSP_KIND_REDFT01 4 /* same numeric values as FFTW's fftw_r2r_kind */
#define DSP_KIND_REDFT10 5

namespace dsp {

// ------------------------------------------------------------------------------------------------ small types
template <class T> struct alignas(2 * sizeof(T)) C2 { T x, y; };

template <class T> struct VecOf;               // 16-byte global access unit
template <> struct VecOf<float>  { enum { N = 4 }; struct alignas(16) type { float v[4]; }; };
template <> struct VecOf<double> { enum { N = 2 }; struct alignas(16) type { double v[2]; }; };

// unsigned division by a runtime constant: q = n / d for n < 2^31 (same scheme as CUTLASS FastDivmod)
struct FastDiv { uint32_t mul, shr, d; };
DSP_DEV uint32_t fd_div(uint32_t n, const FastDiv &f) {
#if DSP_GPU
	return f.d == 1 ? n : (__umulhi(n, f.mul) >> f.shr);
#else
	return f.d == 1 ? n : (uint32_t)((((uint64_t)n * f.mul) >> 32) >> f.shr);
#endif
}

// Bank-skew padding of a complex index: every hex (f32) / octal (f64) digit of the index is added into the low
// digit, so any access pattern whose lanes differ by a power-of-two stride lands in distinct banks.
template <class T> struct Pad;
template <> struct Pad<float>  { DSP_HDM static int of(int e) { return e + (e >> 4) + (e >> 8) + (e >> 12); } };
template <> struct Pad<double> { DSP_HDM static int of(int e) { return e + (e >> 3) + (e >> 6) + (e >> 9) + (e >> 12); } };

// ------------------------------------------------------------------------------------------------ descriptors
struct FftDesc {
	int n;                       // transform length
	int nfac;                    // number of DIF passes
	int fac[DSP_MAX_FAC];        // radix of each pass
	int npad;                    // sequence stride in smem, complex elements
	const void *tw;              // C2<T>[n]      W_n^k = (cos 2 pi k/n, -sin 2 pi k/n)
	const void *om;              // C2<T>[n/2+1]  (cos pi k/2n, sin pi k/2n)
	const uint16_t *pos2;        // [n] smem slot that holds frequency k after the DIF passes
	const uint16_t *pos3;        // [n] smem slot that holds DCT-III output sample j: pos2[perm(j)]
	FastDiv dM[DSP_MAX_FAC];     // divide by M_p = L_p / r_p
	FastDiv dNb[DSP_MAX_FAC];    // divide by n / r_p (butterflies per sequence in pass p)
	FastDiv dHalf;               // divide by n/2+1
	// dense fallback (lengths with a prime factor > 13): direct O(n^2) evaluation of the definition
	int dense;                   // 1: no FFT; input at [0, half), output at [half, 2*half) of each sequence
	int half;                    // complex elements per half
	const void *ctab;            // T[4n]  cos(pi m / 2n)
	FastDiv dN;                  // divide by n
};

// Up to four outer loop levels around a pass (level 0 fastest).  slot: which coordinate the level feeds
// (0..2 = logical plan index i0..i2, 3 = channel, 4 = batch, -1 = none).
struct Outer {
	int cnt[4];
	long long is[4], os[4];
	int slot[4];
	FastDiv d0, d01, d012;       // divide by cnt[0], cnt[0]*cnt[1], cnt[0]*cnt[1]*cnt[2]
	int base_slot, base;         // chunked launches (L2-resident schedule): the launch covers indices base.. of the outermost
	                             // level; pointers are pre-offset by the host, coordinates get the base back here
};

// Logical coordinates of an element.  Named fields + select-based set(): a dynamically indexed array would be
// demoted to local memory.
struct Coord {
	int i0, i1, i2, ch, b;
	DSP_HDM void set(int slot, int v) {
		i0 = slot == 0 ? v : i0; i1 = slot == 1 ? v : i1; i2 = slot == 2 ? v : i2;
		ch = slot == 3 ? v : ch; b = slot == 4 ? v : b;
	}
};

// ------------------------------------------------------------------------------------------------ complex helpers
template <class T> DSP_DEV C2<T> cadd(C2<T> a, C2<T> b) { return C2<T>{a.x + b.x, a.y + b.y}; }
template <class T> DSP_DEV C2<T> csub(C2<T> a, C2<T> b) { return C2<T>{a.x - b.x, a.y - b.y}; }
#if defined(__CUDA_ARCH__) && !defined(DSP_NO_F32X2)
// sm_100 packed fp32: one FADD2 per complex add / subtract (same rounding per component as FADD; the negation of
// the subtrahend folds into the instruction's operand modifier).  The butterflies are issue-bound, and adds are
// ~80% of their arithmetic, so this is the cheapest way to retire fewer instructions per sample.
template <> DSP_DEV C2<float> cadd<float>(C2<float> a, C2<float> b) {
	const float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
	return C2<float>{r.x, r.y};
}
template <> DSP_DEV C2<float> csub<float>(C2<float> a, C2<float> b) {
	const float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(-b.x, -b.y));
	return C2<float>{r.x, r.y};
}
#endif
template <class T> DSP_DEV C2<T> cmul(C2<T> a, C2<T> b) { return C2<T>{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
template <class T> DSP_DEV C2<T> cmulc(C2<T> a, T br, T bi) { return C2<T>{a.x * br - a.y * bi, a.x * bi + a.y * br}; }
template <class T> DSP_DEV C2<T> mul_mi(C2<T> a) { return C2<T>{a.y, -a.x}; }   // a * (-i)

// streaming 16-byte global load that does not allocate in L1 (keeps the twiddle / slot tables resident there)
#if DSP_GPU
DSP_DEV VecOf<float>::type ldg_stream(const VecOf<float>::type *p) {
	VecOf<float>::type r;
	asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]) : "l"(p));
	return r;
}
DSP_DEV VecOf<double>::type ldg_stream(const VecOf<double>::type *p) {
	VecOf<double>::type r;
	asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.v[0]), "=d"(r.v[1]) : "l"(p));
	return r;
}
#endif
template <class V> DSP_DEV V ldg_stream(const V *p) { return *p; }      // any other vector type: plain load

// L2 prefetch of the 128-byte line holding p (no data returned; the transfer overlaps whatever runs next)
DSP_DEV void prefetch_l2(const void *p) {
#if DSP_GPU
	asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
	(void)p;
#endif
}

// bulk L2 prefetch of `bytes` (multiple of 16) starting at the 16-byte aligned p: one instruction handed to the copy
// engine instead of one LSU request per 128-byte line
DSP_DEV void prefetch_l2_bulk(const void *p, uint32_t bytes) {
#if DSP_GPU
	asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes));
#else
	(void)p; (void)bytes;
#endif
}

// read-only (LDG) load of a complex table entry
#if DSP_GPU
DSP_DEV C2<float> ldg_c2(const C2<float> *p) { const float2 t = __ldg((const float2 *)p); return C2<float>{t.x, t.y}; }
DSP_DEV C2<double> ldg_c2(const C2<double> *p) { const double2 t = __ldg((const double2 *)p); return C2<double>{t.x, t.y}; }
#else
template <class T> DSP_DEV C2<T> ldg_c2(const C2<T> *p) { return *p; }
#endif

// ------------------------------------------------------------------------------------------------ small DFTs
// All are forward transforms: y[m] = sum_j x[j] exp(-2 pi i j m / R), in place on a register array.
template <class T> DSP_DEV void dft2(C2<T> &a, C2<T> &b) { C2<T> t = a; a = cadd(t, b); b = csub(t, b); }

template <class T> DSP_DEV void dft4(C2<T> &a0, C2<T> &a1, C2<T> &a2, C2<T> &a3) {
	C2<T> s0 = cadd(a0, a2), d0 = csub(a0, a2), s1 = cadd(a1, a3), d1 = mul_mi(csub(a1, a3));
	a0 = cadd(s0, s1); a2 = csub(s0, s1); a1 = cadd(d0, d1); a3 = csub(d0, d1);
}

template <class T, int R> struct Dft;

template <class T> struct Dft<T, 2> { DSP_DEVM static void run(C2<T> *v) { dft2(v[0], v[1]); } };
template <class T> struct Dft<T, 4> { DSP_DEVM static void run(C2<T> *v) { dft4(v[0], v[1], v[2], v[3]); } };

template <class T> struct Dft<T, 8> {
	DSP_DEVM static void run(C2<T> *v) {
		const T h = (T)0.70710678118654752440084436210484903928;
		// j = 2a + b : radix-4 over a for b = 0,1, twiddle W8^{b q}, radix-2 over b; output m = q + 4p
		dft4(v[0], v[2], v[4], v[6]);
		dft4(v[1], v[3], v[5], v[7]);
		v[3] = cmulc(v[3], h, -h);            // W8^1
		v[5] = mul_mi(v[5]);                  // W8^2
		v[7] = cmulc(v[7], -h, -h);           // W8^3
		C2<T> y0 = v[0], y1 = v[2], y2 = v[4], y3 = v[6];
		C2<T> z0 = v[1], z1 = v[3], z2 = v[5], z3 = v[7];
		v[0] = cadd(y0, z0); v[4] = csub(y0, z0);
		v[1] = cadd(y1, z1); v[5] = csub(y1, z1);
		v[2] = cadd(y2, z2); v[6] = csub(y2, z2);
		v[3] = cadd(y3, z3); v[7] = csub(y3, z3);
	}
};

template <class T> struct Dft<T, 16> {
	DSP_DEVM static void run(C2<T> *v) {
		const T c1 = (T)0.92387953251128675612818318939678828682;   // cos(pi/8)
		const T s1 = (T)0.38268343236508977172845998403039886676;   // sin(pi/8)
		const T h  = (T)0.70710678118654752440084436210484903928;
		// j = 4a + b : radix-4 over a (stride 4) for each b; t[b][q] sits in v[4q + b]
		dft4(v[0], v[4], v[8],  v[12]);
		dft4(v[1], v[5], v[9],  v[13]);
		dft4(v[2], v[6], v[10], v[14]);
		dft4(v[3], v[7], v[11], v[15]);
		// twiddle W16^{b q}
		v[5]  = cmulc(v[5],  c1, -s1);   // b=1,q=1 : W16^1
		v[6]  = cmulc(v[6],  h,  -h);    // b=2,q=1 : W16^2
		v[7]  = cmulc(v[7],  s1, -c1);   // b=3,q=1 : W16^3
		v[9]  = cmulc(v[9],  h,  -h);    // b=1,q=2 : W16^2
		v[10] = mul_mi(v[10]);           // b=2,q=2 : W16^4
		v[11] = cmulc(v[11], -h, -h);    // b=3,q=2 : W16^6
		v[13] = cmulc(v[13], s1, -c1);   // b=1,q=3 : W16^3
		v[14] = cmulc(v[14], -h, -h);    // b=2,q=3 : W16^6
		v[15] = cmulc(v[15], -c1, s1);   // b=3,q=3 : W16^9
		// radix-4 over b for each q: outputs m = q + 4p land in v[4q + p]; then transpose to natural order
		dft4(v[0],  v[1],  v[2],  v[3]);
		dft4(v[4],  v[5],  v[6],  v[7]);
		dft4(v[8],  v[9],  v[10], v[11]);
		dft4(v[12], v[13], v[14], v[15]);
		C2<T> t;
#define DSP_SWAP(a, b) t = v[a]; v[a] = v[b]; v[b] = t;
		DSP_SWAP(1, 4) DSP_SWAP(2, 8) DSP_SWAP(3, 12) DSP_SWAP(6, 9) DSP_SWAP(7, 13) DSP_SWAP(11, 14)
#undef DSP_SWAP
	}
};

template <class T> struct Dft<T, 3> {
	DSP_DEVM static void run(C2<T> *v) {
		const T s = (T)0.86602540378443864676372317075293618347;    // sin(2 pi/3)
		C2<T> a = cadd(v[1], v[2]), b = csub(v[1], v[2]);
		C2<T> m = C2<T>{v[0].x - (T)0.5 * a.x, v[0].y - (T)0.5 * a.y};
		C2<T> r = C2<T>{s * b.y, -s * b.x};                          // -i s b
		v[0] = cadd(v[0], a); v[1] = cadd(m, r); v[2] = csub(m, r);
	}
};

// odd prime radix, direct symmetric form.  cs[k] = cos(2 pi k/R), sn[k] = sin(2 pi k/R), k < R.
template <class T, int R> struct DftOdd {
	DSP_DEVM static void run(C2<T> *v, const T *cs, const T *sn) {
		C2<T> a[(R - 1) / 2], b[(R - 1) / 2], y[R];
#pragma unroll
		for (int j = 1; j <= (R - 1) / 2; j++) { a[j - 1] = cadd(v[j], v[R - j]); b[j - 1] = csub(v[j], v[R - j]); }
		y[0] = v[0];
#pragma unroll
		for (int j = 0; j < (R - 1) / 2; j++) y[0] = cadd(y[0], a[j]);
#pragma unroll
		for (int m = 1; m <= (R - 1) / 2; m++) {
			T pr = v[0].x, pi = v[0].y, qr = 0, qi = 0;
#pragma unroll
			for (int j = 1; j <= (R - 1) / 2; j++) {
				const T c = cs[(j * m) % R], s = sn[(j * m) % R];
				pr += c * a[j - 1].x; pi += c * a[j - 1].y;
				qr += s * b[j - 1].y; qi -= s * b[j - 1].x;         // -i s b
			}
			y[m] = C2<T>{pr + qr, pi + qi};
			y[R - m] = C2<T>{pr - qr, pi - qi};
		}
#pragma unroll
		for (int m = 0; m < R; m++) v[m] = y[m];
	}
};

// composite radix R = R1 R2 in registers: x[R2 a + b] -> R1-point DFTs over a, twiddle W_R^{bq}, R2-point DFTs over
// b -> y[q + R1 p].  Merging 3*5 and 3*3 saves one full shared-memory pass each (1080 = 15*9*8, 1920 = 15*8*16).
template <class T, int R1, int R2> struct DftComp {
	DSP_DEVM static void run(C2<T> *v, const T *cs, const T *sn) {
		const int R = R1 * R2;
		C2<T> t[R2][R1];
#pragma unroll
		for (int b = 0; b < R2; b++) {
			C2<T> u[R1];
#pragma unroll
			for (int a = 0; a < R1; a++) u[a] = v[R2 * a + b];
			Dft<T, R1>::run(u);
#pragma unroll
			for (int q = 0; q < R1; q++) t[b][q] = u[q];
		}
#pragma unroll
		for (int b = 1; b < R2; b++)
#pragma unroll
			for (int q = 1; q < R1; q++) t[b][q] = cmulc(t[b][q], cs[(b * q) % R], -sn[(b * q) % R]);
#pragma unroll
		for (int q = 0; q < R1; q++) {
			C2<T> u[R2];
#pragma unroll
			for (int b = 0; b < R2; b++) u[b] = t[b][q];
			Dft<T, R2>::run(u);
#pragma unroll
			for (int p = 0; p < R2; p++) v[q + R1 * p] = u[p];
		}
	}
};
#define DSP_DEFINE_COMP(R, R1, R2, CS, SN)                                            \
	template <class T> struct Dft<T, R> {                                             \
		DSP_DEVM static void run(C2<T> *v) {                                           \
			const T cs[R] = CS;                                                       \
			const T sn[R] = SN;                                                       \
			DftComp<T, R1, R2>::run(v, cs, sn);                                       \
		}                                                                             \
	};

#define DSP_DEFINE_ODD(R, CS, SN)                                                     \
	template <class T> struct Dft<T, R> {                                             \
		DSP_DEVM static void run(C2<T> *v) {                                           \
			const T cs[R] = CS;                                                       \
			const T sn[R] = SN;                                                       \
			DftOdd<T, R>::run(v, cs, sn);                                             \
		}                                                                             \
	};
#define DSP_L(...) { __VA_ARGS__ }
#include "dct_oddtabs.inc"
#undef DSP_L
#undef DSP_DEFINE_ODD
#undef DSP_DEFINE_COMP

// ------------------------------------------------------------------------------------------------ DIF passes
// In-place decimation-in-frequency pass p: sub-length L, radix R, M = L/R.  Butterfly (blk, i) reads
// e = blk*L + i + j*M, computes the R-point DFT, multiplies output m by W_L^{i m} and stores it back to
// e = blk*L + i + m*M.  After the last pass frequency k sits at the mixed-radix digit-reversed slot pos2[k].
template <class T, int R>
DSP_DEV void radix_pass(C2<T> *s, int nseq, const FftDesc &f, int p, int L, int tid, int nthr) {
	const int M = L / R;
	const uint32_t nb = (uint32_t)(f.n / R);
	const uint32_t total = (uint32_t)nseq * nb;
	const int twstep = f.n / L;
	const C2<T> *tw = (const C2<T> *)f.tw;
	for (uint32_t g = (uint32_t)tid; g < total; g += (uint32_t)nthr) {
		const uint32_t seq = fd_div(g, f.dNb[p]);
		const uint32_t b = g - seq * nb;
		const uint32_t blk = fd_div(b, f.dM[p]);
		const int i = (int)(b - blk * (uint32_t)M);
		C2<T> *base = s + (size_t)seq * (size_t)f.npad;
		const int e0 = (int)blk * L + i;
		C2<T> v[R];
#pragma unroll
		for (int j = 0; j < R; j++) v[j] = base[Pad<T>::of(e0 + j * M)];
		Dft<T, R>::run(v);
		if (M > 1) {
			const int ti = i * twstep;
#pragma unroll
			for (int m = 1; m < R; m++) v[m] = cmul(v[m], ldg_c2(tw + ti * m));
		}
#pragma unroll
		for (int m = 0; m < R; m++) base[Pad<T>::of(e0 + m * M)] = v[m];
	}
}

template <class T>
DSP_DEV void fft_dif(C2<T> *s, int nseq, const FftDesc &f, int t0, int t1, int nthr) {
	int L = f.n;
	for (int p = 0; p < f.nfac; p++) {
		const int r = f.fac[p];
		for (int tid = t0; tid < t1; tid++) {
			switch (r) {
			case 2:  radix_pass<T, 2>(s, nseq, f, p, L, tid, nthr); break;
			case 3:  radix_pass<T, 3>(s, nseq, f, p, L, tid, nthr); break;
			case 4:  radix_pass<T, 4>(s, nseq, f, p, L, tid, nthr); break;
			case 5:  radix_pass<T, 5>(s, nseq, f, p, L, tid, nthr); break;
			case 7:  radix_pass<T, 7>(s, nseq, f, p, L, tid, nthr); break;
			case 8:  radix_pass<T, 8>(s, nseq, f, p, L, tid, nthr); break;
			case 9:  radix_pass<T, 9>(s, nseq, f, p, L, tid, nthr); break;
			case 15: radix_pass<T, 15>(s, nseq, f, p, L, tid, nthr); break;
			case 11: radix_pass<T, 11>(s, nseq, f, p, L, tid, nthr); break;
			case 13: radix_pass<T, 13>(s, nseq, f, p, L, tid, nthr); break;
			case 16: radix_pass<T, 16>(s, nseq, f, p, L, tid, nthr); break;
			default: break;
			}
		}
		DSP_SYNC();
		L /= r;
	}
}

// ------------------------------------------------------------------------------------------------ pair phases
// DCT-III pre-twiddle.  Slot k holds (XA[k], XB[k]) (natural order).  Writes conj(Z[k]) with
// Z = UA + i UB, U[k] = e^{+i pi k/2n} (X[k] - i X[n-k]), so that REDFT01(X)[perm^-1] = conj(FFT(conj Z)).
template <class T>
DSP_DEV void dct3_pre(C2<T> *s, int nseq, const FftDesc &f, int tid, int nthr) {
	const int n = f.n;
	const uint32_t half = (uint32_t)(n / 2 + 1);
	const uint32_t total = (uint32_t)nseq * half;
	const C2<T> *om = (const C2<T> *)f.om;
	for (uint32_t g = (uint32_t)tid; g < total; g += (uint32_t)nthr) {
		const uint32_t seq = fd_div(g, f.dHalf);
		const int k = (int)(g - seq * half);
		C2<T> *base = s + (size_t)seq * (size_t)f.npad;
		const int pk = Pad<T>::of(k);
		const C2<T> zk = base[pk];
		if (k == 0) { base[pk] = C2<T>{zk.x, -zk.y}; continue; }
		const C2<T> w = ldg_c2(om + k);
		const int pn = Pad<T>::of(n - k);
		const C2<T> zn = (2 * k == n) ? zk : base[pn];
		const T pa = w.x * zk.x + w.y * zn.x, qa = w.y * zk.x - w.x * zn.x;
		const T pb = w.x * zk.y + w.y * zn.y, qb = w.y * zk.y - w.x * zn.y;
		base[pk] = C2<T>{pa - qb, -(qa + pb)};
		if (2 * k != n) base[pn] = C2<T>{pa + qb, -(pb - qa)};
	}
}

// DCT-II post-twiddle.  Slot pos2[k] holds Z[k] = A[k] + i B[k] (A, B the Hermitian spectra of the two real
// sequences).  Writes (XA[k], XB[k]) back into slot pos2[k] and (XA[n-k], XB[n-k]) into slot pos2[n-k].
template <class T>
DSP_DEV void dct2_post(C2<T> *s, int nseq, const FftDesc &f, int tid, int nthr) {
	const int n = f.n;
	const uint32_t half = (uint32_t)(n / 2 + 1);
	const uint32_t total = (uint32_t)nseq * half;
	const C2<T> *om = (const C2<T> *)f.om;
	for (uint32_t g = (uint32_t)tid; g < total; g += (uint32_t)nthr) {
		const uint32_t seq = fd_div(g, f.dHalf);
		const int k = (int)(g - seq * half);
		C2<T> *base = s + (size_t)seq * (size_t)f.npad;
		const int pk = Pad<T>::of((int)DSP_LDG(f.pos2 + k));
		const int pn = Pad<T>::of((int)DSP_LDG(f.pos2 + (k == 0 ? 0 : n - k)));
		const C2<T> z = base[pk], y = base[pn];
		const C2<T> w = ldg_c2(om + k);
		const T ar = z.x + y.x, ai = z.y - y.y, br = z.y + y.y, bi = y.x - z.x;
		base[pk] = C2<T>{w.x * ar + w.y * ai, w.x * br + w.y * bi};
		if (k != 0 && 2 * k != n) base[pn] = C2<T>{w.y * ar - w.x * ai, w.y * br - w.x * bi};
	}
}

// ------------------------------------------------------------------------------------------------ dense fallback
// Y_k = 2 sum_j x_j cos(pi (2j+1) k / 2n)            (REDFT10)
// Y_k = x_0 + 2 sum_{j>=1} x_j cos(pi j (2k+1) / 2n)  (REDFT01)
// evaluated directly on the (A, B) complex pairs; the cosine comes from a 4n-periodic table.
template <class T>
DSP_DEV void dense_dct(C2<T> *s, int nseq, const FftDesc &f, bool fwd, int tid, int nthr) {
	const int n = f.n, p4 = 4 * n;
	const T *ct = (const T *)f.ctab;
	const uint32_t total = (uint32_t)nseq * (uint32_t)n;
	for (uint32_t g = (uint32_t)tid; g < total; g += (uint32_t)nthr) {
		const uint32_t seq = fd_div(g, f.dN);
		const int k = (int)(g - seq * (uint32_t)n);
		const C2<T> *in = s + (size_t)seq * (size_t)f.npad;
		T ar = 0, ai = 0;
		if (fwd) {
			int idx = k % p4;
			const int step = (2 * k) % p4;
			for (int j = 0; j < n; j++) {
				const T c = DSP_LDG(ct + idx);
				const C2<T> x = in[Pad<T>::of(j)];
				ar += x.x * c; ai += x.y * c;
				idx += step; if (idx >= p4) idx -= p4;
			}
			ar *= 2; ai *= 2;
		} else {
			const int step = (2 * k + 1) % p4;
			int idx = step;
			for (int j = 1; j < n; j++) {
				const T c = DSP_LDG(ct + idx);
				const C2<T> x = in[Pad<T>::of(j)];
				ar += x.x * c; ai += x.y * c;
				idx += step; if (idx >= p4) idx -= p4;
			}
			const C2<T> x0 = in[Pad<T>::of(0)];
			ar = x0.x + 2 * ar; ai = x0.y + 2 * ai;
		}
		s[(size_t)seq * (size_t)f.npad + f.half + Pad<T>::of(k)] = C2<T>{ar, ai};
	}
}

// the transform phases between copy-in and copy-out
template <class T>
DSP_DEV void transform_phases(C2<T> *s, int nseq, const FftDesc &f, bool fwd, int t0, int t1, int nthr) {
	if (f.dense) {
		for (int tid = t0; tid < t1; tid++) dense_dct<T>(s, nseq, f, fwd, tid, nthr);
		DSP_SYNC();
		return;
	}
	if (!fwd) {
		for (int tid = t0; tid < t1; tid++) dct3_pre<T>(s, nseq, f, tid, nthr);
		DSP_SYNC();
	}
	fft_dif<T>(s, nseq, f, t0, t1, nthr);
	if (fwd) {
		for (int tid = t0; tid < t1; tid++) dct2_post<T>(s, nseq, f, tid, nthr);
		DSP_SYNC();
	}
}

// ------------------------------------------------------------------------------------------------ fused ops
// A load op maps the value read from global memory to the value entering the transform; a store op maps the
// transform output to the value written.  Both see the element's logical coordinates.
// v * f: the lean kernels' only pointwise stage (f = 1 when nothing is fused; dsp_dct_fuse_scale sets it)
template <class T> struct OpMul {
	enum { kNeedsCoord = 0 };
	T f;
	DSP_DEVM T operator()(T v, const Coord &) const { return v * f; }
};
struct OpNone {
	enum { kNeedsCoord = 0 };
	template <class T> DSP_DEVM T operator()(T v, const Coord &) const { return v; }
};
struct OpScale {                 // v * a   (e.g. scan's 1/(4wh), spec/ispec plain normalisations)
	enum { kNeedsCoord = 0 };
	double a;
	template <class T> DSP_DEVM T operator()(T v, const Coord &) const { return v * (T)a; }
};

DSP_DEV void outer_decode(const Outer &o, uint32_t l, long long &ioff, long long &ooff, Coord &c) {
	const uint32_t l3 = fd_div(l, o.d012);
	const uint32_t r3 = l - l3 * o.d012.d;
	const uint32_t l2 = fd_div(r3, o.d01);
	const uint32_t r = r3 - l2 * o.d01.d;
	const uint32_t l1 = fd_div(r, o.d0);
	const uint32_t l0 = r - l1 * o.d0.d;
	ioff = (long long)l0 * o.is[0] + (long long)l1 * o.is[1] + (long long)l2 * o.is[2] + (long long)l3 * o.is[3];
	ooff = (long long)l0 * o.os[0] + (long long)l1 * o.os[1] + (long long)l2 * o.os[2] + (long long)l3 * o.os[3];
	c.set(o.slot[0], (int)l0); c.set(o.slot[1], (int)l1); c.set(o.slot[2], (int)l2); c.set(o.slot[3], (int)l3);
	if (o.base) {
		c.i0 += o.base_slot == 0 ? o.base : 0; c.i1 += o.base_slot == 1 ? o.base : 0; c.i2 += o.base_slot == 2 ? o.base : 0;
		c.ch += o.base_slot == 3 ? o.base : 0; c.b += o.base_slot == 4 ? o.base : 0;
	}
}

// ------------------------------------------------------------------------------------------------ lean tile moves
DSP_DEV int ilog2(int v) { int l = 0; while ((1 << l) < v) l++; return l; }
// Makhoul's permutation: position of x[j] in the even/odd-reordered sequence
DSP_DEV int makhoul(int x, int n) { return (x & 1) ? n - 1 - (x >> 1) : (x >> 1); }

// W consecutive elements of T as one global access (W * sizeof(T) in {8, 16} bytes)
template <class T, int W> struct alignas(W * sizeof(T)) VecW { T v[W]; };
#if DSP_GPU
DSP_DEV VecW<float, 4> ldg_stream(const VecW<float, 4> *p) {
	VecW<float, 4> r;
	asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]) : "l"(p));
	return r;
}
DSP_DEV VecW<float, 2> ldg_stream(const VecW<float, 2> *p) {
	VecW<float, 2> r;
	asm volatile("ld.global.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.v[0]), "=f"(r.v[1]) : "l"(p));
	return r;
}
DSP_DEV VecW<double, 2> ldg_stream(const VecW<double, 2> *p) {
	VecW<double, 2> r;
	asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.v[0]), "=d"(r.v[1]) : "l"(p));
	return r;
}
#endif


// Lean tile move for the common case -- float, full 16-byte groups, a power-of-two number of groups per row that
// divides the thread count, and a pointwise stage that ignores coordinates (OpMul): each thread keeps one column
// group and walks down the rows, so the only per-row work is the row / slot mapping.
//   rowmap.off(r, rs) -> element offset of tile row r (rs = row stride)      slotmap(r) -> padded smem slot of row r
template <class T, bool IN, class Op, class RowMap, class SlotMap>
DSP_DEV void tile_move_lean(const T *gin, T *gout, long long rs, int nrows, int lg, const Op &op, bool negim, const RowMap &rowmap,
                            const SlotMap &slotmap, int npad, int tid, int nthr, C2<T> *s) {
	typedef VecW<T, 4> Vec;
	const int UNR = 8;
	const int cg = tid & ((1 << lg) - 1), dr = nthr >> lg;
	C2<T> *sq = s + (2 * cg) * npad;
	const T *gp = gin + 4 * cg;
	T *gq = gout + 4 * cg;
	const Coord cz = {0, 0, 0, 0, 0};
	for (int r0 = tid >> lg; r0 < nrows; r0 += dr * UNR) {
		Vec v[UNR];
		if (IN) {
#pragma unroll
			for (int u = 0; u < UNR; u++) {
				const int r = r0 + u * dr;
				if (r < nrows) v[u] = ldg_stream((const Vec *)(gp + rowmap.off(r, rs)));
			}
		}
#pragma unroll
		for (int u = 0; u < UNR; u++) {
			const int r = r0 + u * dr;
			if (r < nrows) {
				const int slot = slotmap(r);
				if (IN) {
					sq[slot] = C2<T>{op(v[u].v[0], cz), op(v[u].v[1], cz)};
					sq[npad + slot] = C2<T>{op(v[u].v[2], cz), op(v[u].v[3], cz)};
				} else {
					const C2<T> z0 = sq[slot], z1 = sq[npad + slot];
					Vec o;
					o.v[0] = op(z0.x, cz); o.v[1] = op(negim ? -z0.y : z0.y, cz);
					o.v[2] = op(z1.x, cz); o.v[3] = op(negim ? -z1.y : z1.y, cz);
					*(Vec *)(gq + rowmap.off(r, rs)) = o;
				}
			}
		}
	}
}
struct RowIdent { DSP_DEVM long long off(int r, long long rs) const { return (long long)r * rs; } };
// segmented output: rows [g S, (g+1) S) live at element offset seg[g] (from the pass's output pointer) with the same
// row stride -- the peer GPUs' buffers of the slab-sharded 3-D transform (ColArgs::seg_rows, dsp_dct_set_output_segments)
struct RowSeg {
	int S; FastDiv dS; const long long *seg;
	DSP_DEVM long long off(int r, long long rs) const {
		const int g = (int)fd_div((uint32_t)r, dS);
		return seg[g] + (long long)(r - g * S) * rs;
	}
};
template <class T> struct SlotNat { DSP_DEVM int operator()(int r) const { return Pad<T>::of(r); } };
DSP_DEV bool lean_ok(int ncl, int tc, int nthr, bool aligned) {
	const int gpr = tc / 4;
	return aligned && ncl == tc && (tc % 4) == 0 && (gpr & (gpr - 1)) == 0 && (nthr % gpr) == 0;
}

// slot maps of the generic engine: natural-order input of the DIF passes, digit-reversed output via pos2 / pos3
template <class T> struct SlotMakhoulNat { int n; DSP_DEVM int operator()(int r) const { return Pad<T>::of(makhoul(r, n)); } };
template <class T> struct SlotTab { const uint16_t *pos; DSP_DEVM int operator()(int r) const { return Pad<T>::of((int)DSP_LDG(pos + r)); } };

// Threads of a CTA over (line pair, vector group): `gp2` threads (a power of two >= the groups per line, capped at the
// CTA) share one pair, nthr / gp2 pairs are moved at a time -- short lines (e.g. the 8-point transforms of a block
// DCT) would otherwise leave all but a handful of threads idle in a pair-by-pair walk.
struct PairThreads {
	int gp2, ngrp, grp, ql;
	DSP_DEVM PairThreads(int gpl, int tid, int nthr) {
		gp2 = 1;
		while (gp2 < gpl && gp2 < nthr) gp2 <<= 1;
		ngrp = nthr / gp2; grp = tid / gp2; ql = tid - grp * gp2;
	}
};

// Row moves of the generic engine for planar float lines (d == 1, 16-byte access, whole line pairs, coordinate-free
// op): one vector group = x in [4q, 4q+4) of lines A and B.  IN && FWD: Makhoul scatter to natural-order slots;
// IN && !FWD: natural order; !IN: gather through pos (pos2 / pos3, one 8-byte table load per group).
template <class T, bool IN, bool FWD, class Op>
DSP_DEV void row_move_generic_lean(const T *gin, T *gout, long long ls, int line0, int npairs, int n, int npad, const uint16_t *pos,
                                   const Op &op, int tid, int nthr, C2<T> *s) {
	typedef VecW<T, 4> Vec;
	struct alignas(8) U4 { uint16_t v[4]; };
	const int gpl = n >> 2;
	const Coord cz = {0, 0, 0, 0, 0};
	const PairThreads pt(gpl, tid, nthr);
	for (int g = pt.grp; g < npairs; g += pt.ngrp) {
		const long long l = line0 + 2 * g;
		const Vec *pa = (const Vec *)(gin + l * ls), *pb = (const Vec *)(gin + (l + 1) * ls);
		Vec *qa = (Vec *)(gout + l * ls), *qb = (Vec *)(gout + (l + 1) * ls);
		C2<T> *sg = s + (size_t)g * (size_t)npad;
		for (int q = pt.ql; q < gpl; q += pt.gp2) {
			if (IN) {
				const Vec ta = ldg_stream(pa + q), tb = ldg_stream(pb + q);
				int sl[4];
				if (FWD) { sl[0] = Pad<T>::of(2 * q); sl[2] = Pad<T>::of(2 * q + 1); sl[1] = Pad<T>::of(n - 1 - 2 * q); sl[3] = Pad<T>::of(n - 2 - 2 * q); }
				else { sl[0] = Pad<T>::of(4 * q); sl[1] = Pad<T>::of(4 * q + 1); sl[2] = Pad<T>::of(4 * q + 2); sl[3] = Pad<T>::of(4 * q + 3); }
#pragma unroll
				for (int t = 0; t < 4; t++) sg[sl[t]] = C2<T>{op(ta.v[t], cz), op(tb.v[t], cz)};
			} else {
				const U4 pp = *(const U4 *)(pos + 4 * q);
				Vec ra, rb;
#pragma unroll
				for (int t = 0; t < 4; t++) {
					const C2<T> z = sg[Pad<T>::of((int)pp.v[t])];
					ra.v[t] = op(z.x, cz);
					rb.v[t] = op(FWD ? z.y : -z.y, cz);
				}
				qa[q] = ra;
				qb[q] = rb;
			}
		}
	}
}

// ------------------------------------------------------------------------------------------------ row pass
// Transform along the contiguous axis.  A "line" is n*d contiguous elements: d interleaved sequences of
// length n (d = 1 planar, d = 3 for dspfun's RGB images).  Lines (2g, 2g+1) of the CTA's range ride together.
struct RowArgs {
	FftDesc f;
	int kind;                    // DSP_KIND_REDFT10 / DSP_KIND_REDFT01
	int d;                       // interleave
	FastDiv dd;                  // divide by d
	int nlines, lines_per_cta;
	Outer o;                     // line index -> offsets / coords
	int ax_slot;                 // coordinate fed by the axis index
	int pf_dist;                 // > 0: prefetch the input lines of CTA (cta + pf_dist) into L2 while this one computes
	int simple;                  // line l sits at l * ls_in / l * ls_out (all outer levels collapse to one stride)
	long long ls_in, ls_out;
	const void *in;
	void *out;
	int vec_in, vec_out;         // 16-byte access legal
	int in_u8, out_u8;           // the global side holds unsigned 8-bit samples (motion's pels); implies vec_* == 0
};

// 8-bit store: the value has already been clamped and rounded by the store op (motion/motion.c:776)
template <class T> DSP_DEV unsigned char to_u8(T v) { return (unsigned char)(int)v; }

template <class T, class LoadOp, class StoreOp>
DSP_DEV void cta_row_pass(const RowArgs &a, const LoadOp &lop, const StoreOp &sop, int cta, int t0, int t1, int nthr,
                          C2<T> *s) {
	typedef typename VecOf<T>::type Vec;
	const T *gin = (const T *)a.in;
	T *gout = (T *)a.out;
	const int VN = VecOf<T>::N;
	const int n = a.f.n, d = a.d;
	const int line0 = cta * a.lines_per_cta;
	int nl = a.nlines - line0;
	if (nl > a.lines_per_cta) nl = a.lines_per_cta;
	const int npairs = (nl + 1) / 2;
	const int nseq = npairs * d;
	const int llen = n * d;                                  // elements per line
	const uint32_t gpl = (uint32_t)((llen + VN - 1) / VN);   // vector groups per line
	const bool fwd = a.kind == DSP_KIND_REDFT10;

	const bool lean = sizeof(T) == 4 && d == 1 && a.simple && a.vec_in && a.vec_out && !LoadOp::kNeedsCoord &&
	                  !StoreOp::kNeedsCoord && !a.f.dense && !(nl & 1) && (n & 3) == 0;

	// ---- copy-in
	for (int tid = t0; tid < t1; tid++) {
		if (lean) {
			if (fwd) row_move_generic_lean<T, true, true, LoadOp>(gin, gout, a.ls_in, line0, npairs, n, a.f.npad, a.f.pos2, lop, tid, nthr, s);
			else row_move_generic_lean<T, true, false, LoadOp>(gin, gout, a.ls_in, line0, npairs, n, a.f.npad, a.f.pos3, lop, tid, nthr, s);
			continue;
		}
		const PairThreads pt((int)gpl, tid, nthr);
		for (int g = pt.grp; g < npairs; g += pt.ngrp) {      // the line decode is per pair, not per vector
			const int la = line0 + 2 * g;
			const bool hasb = (2 * g + 1) < nl;
			Coord ca = {0, 0, 0, 0, 0}, cb = {0, 0, 0, 0, 0};
			long long ia, oa, ib = 0, ob = 0;
			outer_decode(a.o, (uint32_t)la, ia, oa, ca);
			if (hasb) outer_decode(a.o, (uint32_t)la + 1, ib, ob, cb);
		for (uint32_t q = (uint32_t)pt.ql; q < gpl; q += (uint32_t)pt.gp2) {
			const int e0 = (int)q * VN;
			T va[VecOf<T>::N], vb[VecOf<T>::N];
			if (a.vec_in && e0 + VN <= llen) {
				const Vec ta = *(const Vec *)(gin + ia + e0);
#pragma unroll
				for (int t = 0; t < VN; t++) va[t] = ta.v[t];
				if (hasb) {
					const Vec tb = *(const Vec *)(gin + ib + e0);
#pragma unroll
					for (int t = 0; t < VN; t++) vb[t] = tb.v[t];
				}
			} else {
				const unsigned char *g8 = (const unsigned char *)a.in;
#pragma unroll
				for (int t = 0; t < VN; t++) {
					if (a.in_u8) {
						va[t] = (e0 + t < llen) ? (T)g8[ia + e0 + t] : (T)0;
						vb[t] = (hasb && e0 + t < llen) ? (T)g8[ib + e0 + t] : (T)0;
					} else {
						va[t] = (e0 + t < llen) ? gin[ia + e0 + t] : (T)0;
						vb[t] = (hasb && e0 + t < llen) ? gin[ib + e0 + t] : (T)0;
					}
				}
			}
#pragma unroll
			for (int t = 0; t < VN; t++) {
				const int e = e0 + t;
				if (e < llen) {
					const int x = d == 1 ? e : (int)fd_div((uint32_t)e, a.dd), ch = e - x * d;
					ca.set(a.ax_slot, x); ca.ch = ch;
					cb.set(a.ax_slot, x); cb.ch = ch;
					const T pa = lop(va[t], ca);
					const T pb = hasb ? lop(vb[t], cb) : (T)0;
					const int slot = (fwd && !a.f.dense) ? ((x & 1) ? n - 1 - (x >> 1) : (x >> 1)) : x;
					s[(size_t)(g * d + ch) * (size_t)a.f.npad + Pad<T>::of(slot)] = C2<T>{pa, pb};
				}
			}
		}
		}
	}
	DSP_SYNC();

	// ---- transform
	transform_phases<T>(s, nseq, a.f, fwd, t0, t1, nthr);

	// ---- copy-out
	const uint16_t *pos = fwd ? a.f.pos2 : a.f.pos3;
	for (int tid = t0; tid < t1; tid++) {
		if (lean) {
			if (fwd) row_move_generic_lean<T, false, true, StoreOp>(gin, gout, a.ls_out, line0, npairs, n, a.f.npad, pos, sop, tid, nthr, s);
			else row_move_generic_lean<T, false, false, StoreOp>(gin, gout, a.ls_out, line0, npairs, n, a.f.npad, pos, sop, tid, nthr, s);
			continue;
		}
		const PairThreads pt((int)gpl, tid, nthr);
		for (int g = pt.grp; g < npairs; g += pt.ngrp) {
			const int la = line0 + 2 * g;
			const bool hasb = (2 * g + 1) < nl;
			Coord ca = {0, 0, 0, 0, 0}, cb = {0, 0, 0, 0, 0};
			long long ia, oa, ib = 0, ob = 0;
			outer_decode(a.o, (uint32_t)la, ia, oa, ca);
			if (hasb) outer_decode(a.o, (uint32_t)la + 1, ib, ob, cb);
		for (uint32_t q = (uint32_t)pt.ql; q < gpl; q += (uint32_t)pt.gp2) {
			const int e0 = (int)q * VN;
			Vec ra, rb;
#pragma unroll
			for (int t = 0; t < VN; t++) {
				const int e = e0 + t;
				ra.v[t] = 0; rb.v[t] = 0;
				if (e >= llen) continue;
				const int x = d == 1 ? e : (int)fd_div((uint32_t)e, a.dd), ch = e - x * d;
				const C2<T> z = s[(size_t)(g * d + ch) * (size_t)a.f.npad + (a.f.dense ? a.f.half : 0) + Pad<T>::of((int)DSP_LDG(pos + x))];
				ca.set(a.ax_slot, x); ca.ch = ch;
				cb.set(a.ax_slot, x); cb.ch = ch;
				ra.v[t] = sop(z.x, ca);
				rb.v[t] = sop((fwd || a.f.dense) ? z.y : -z.y, cb);
			}
			if (a.vec_out && e0 + VN <= llen) {
				*(Vec *)(gout + oa + e0) = ra;
				if (hasb) *(Vec *)(gout + ob + e0) = rb;
			} else {
				unsigned char *o8 = (unsigned char *)a.out;
#pragma unroll
				for (int t = 0; t < VN; t++)
					if (e0 + t < llen) {
						if (a.out_u8) {
							o8[oa + e0 + t] = to_u8(ra.v[t]);
							if (hasb) o8[ob + e0 + t] = to_u8(rb.v[t]);
						} else {
							gout[oa + e0 + t] = ra.v[t];
							if (hasb) gout[ob + e0 + t] = rb.v[t];
						}
					}
			}
		}
		}
	}
}

// ------------------------------------------------------------------------------------------------ column pass
// Transform along a strided axis.  The CTA owns a tile of `tc` adjacent columns (unit stride in memory) over all
// n positions of the axis; adjacent columns (2p, 2p+1) ride together as one complex sequence.
struct ColArgs {
	FftDesc f;
	int kind;
	int ncols, tc;               // columns in the contiguous run, columns per CTA (even, multiple of the vector width)
	int ntiles;                  // ceil(ncols / tc)
	FastDiv dtiles;
	int d;                       // interleave of the column index: col = x*d + ch
	FastDiv dd;
	long long ax_is, ax_os;      // element stride between successive axis positions (input, output)
	Outer o;                     // outer index (CTA / ntiles) -> offsets / coords
	int ax_slot, col_slot;       // coordinates fed by the axis index and by col / d
	const void *in;
	void *out;
	int vec_in, vec_out;
	int pf_dist;                 // > 0: prefetch the input tile of CTA (cta + pf_dist) into L2 while this one computes
	// segmented output (0 = off): axis positions [g seg_rows, (g+1) seg_rows) are stored at out + seg_off[g] (+ the
	// outer offset, + (position - g seg_rows) * ax_os): the last local pass of the slab-sharded 3-D transform writes
	// straight into the peer GPUs' buffers over NVLink (full aligned float tiles only; checked when it is set)
	int seg_rows;
	FastDiv dseg;
	long long seg_off[8];
};

template <class T, class LoadOp, class StoreOp>
DSP_DEV void cta_col_pass(const ColArgs &a, const LoadOp &lop, const StoreOp &sop, int cta, int t0, int t1, int nthr,
                          C2<T> *s) {
	typedef typename VecOf<T>::type Vec;
	const T *gin = (const T *)a.in;
	T *gout = (T *)a.out;
	const int VN = VecOf<T>::N;
	const int n = a.f.n;
	const uint32_t oidx = fd_div((uint32_t)cta, a.dtiles);
	const int tile = cta - (int)oidx * a.ntiles;
	const int col0 = tile * a.tc;
	int ncl = a.ncols - col0;
	if (ncl > a.tc) ncl = a.tc;
	const int nseq = (ncl + 1) / 2;
	const uint32_t gpr = (uint32_t)((ncl + VN - 1) / VN);    // vector groups per axis position
	int gsh = 0;                                             // groups are indexed on a power-of-two pitch >= gpr
	while ((1u << gsh) < gpr) gsh++;
	const bool fwd = a.kind == DSP_KIND_REDFT10;
	Coord cbase = {0, 0, 0, 0, 0};
	long long ibase, obase;
	outer_decode(a.o, oidx, ibase, obase, cbase);

	// full aligned float tiles with a coordinate-free op: the lean tile moves (one column group per thread)
	const bool lean = sizeof(T) == 4 && !LoadOp::kNeedsCoord && !StoreOp::kNeedsCoord && !a.f.dense &&
	                  lean_ok(ncl, a.tc, nthr, a.vec_in && a.vec_out);
	const int lg = lean ? ilog2(a.tc / 4) : 0;

	// ---- copy-in
	for (int tid = t0; tid < t1; tid++) {
		if (lean) {
			const T *gi = gin + ibase + col0;
			if (fwd) tile_move_lean<T, true, LoadOp>(gi, (T *)0, a.ax_is, n, lg, lop, false, RowIdent(), SlotMakhoulNat<T>{n}, a.f.npad, tid, nthr, s);
			else tile_move_lean<T, true, LoadOp>(gi, (T *)0, a.ax_is, n, lg, lop, false, RowIdent(), SlotNat<T>(), a.f.npad, tid, nthr, s);
			continue;
		}
		for (uint32_t idx = (uint32_t)tid; idx < (uint32_t)n << gsh; idx += (uint32_t)nthr) {
			const uint32_t r = idx >> gsh, cg = idx & ((1u << gsh) - 1u);
			if (cg >= gpr) continue;
			const int c0 = (int)cg * VN;                      // column within the tile
			const T *src = gin + ibase + (long long)r * a.ax_is + col0 + c0;
			T v[VecOf<T>::N];
			if (a.vec_in && c0 + VN <= ncl) {
				const Vec tv = *(const Vec *)src;
#pragma unroll
				for (int t = 0; t < VN; t++) v[t] = tv.v[t];
			} else {
#pragma unroll
				for (int t = 0; t < VN; t++) v[t] = (c0 + t < ncl) ? src[t] : (T)0;
			}
			Coord c = cbase;
			c.set(a.ax_slot, (int)r);
#pragma unroll
			for (int t = 0; t < VN; t++) {
				if (c0 + t < ncl) {
					const int col = col0 + c0 + t;
					const int x = (int)fd_div((uint32_t)col, a.dd);
					c.ch = col - x * a.d; c.set(a.col_slot, x);   // ch first: with a wide interleave col_slot IS the channel slot
					v[t] = lop(v[t], c);
				}
			}
			const int slot = Pad<T>::of((fwd && !a.f.dense) ? (((int)r & 1) ? n - 1 - ((int)r >> 1) : ((int)r >> 1)) : (int)r);
#pragma unroll
			for (int p = 0; p < VN / 2; p++)
				if (c0 + 2 * p < ncl)
					s[(size_t)(c0 / 2 + p) * (size_t)a.f.npad + slot] = C2<T>{v[2 * p], v[2 * p + 1]};
		}
	}
	DSP_SYNC();

	// ---- transform
	transform_phases<T>(s, nseq, a.f, fwd, t0, t1, nthr);

	// ---- copy-out
	const uint16_t *pos = fwd ? a.f.pos2 : a.f.pos3;
	for (int tid = t0; tid < t1; tid++) {
		if (lean) {
			if (a.seg_rows > 0) tile_move_lean<T, false, StoreOp>((const T *)0, gout + obase + col0, a.ax_os, n, lg, sop, !fwd, RowSeg{a.seg_rows, a.dseg, a.seg_off}, SlotTab<T>{pos}, a.f.npad, tid, nthr, s);
			else tile_move_lean<T, false, StoreOp>((const T *)0, gout + obase + col0, a.ax_os, n, lg, sop, !fwd, RowIdent(), SlotTab<T>{pos}, a.f.npad, tid, nthr, s);
			continue;
		}
		for (uint32_t idx = (uint32_t)tid; idx < (uint32_t)n << gsh; idx += (uint32_t)nthr) {
			const uint32_t r = idx >> gsh, cg = idx & ((1u << gsh) - 1u);
			if (cg >= gpr) continue;
			const int c0 = (int)cg * VN;
			const int slot = (a.f.dense ? a.f.half : 0) + Pad<T>::of((int)DSP_LDG(pos + r));
			Coord c = cbase;
			c.set(a.ax_slot, (int)r);
			Vec res;
#pragma unroll
			for (int p = 0; p < VN / 2; p++) {
				res.v[2 * p] = 0; res.v[2 * p + 1] = 0;
				if (c0 + 2 * p < ncl) {
					const C2<T> z = s[(size_t)(c0 / 2 + p) * (size_t)a.f.npad + slot];
					res.v[2 * p] = z.x;
					res.v[2 * p + 1] = (fwd || a.f.dense) ? z.y : -z.y;
				}
			}
#pragma unroll
			for (int t = 0; t < VN; t++) {
				if (c0 + t < ncl) {
					const int col = col0 + c0 + t;
					const int x = (int)fd_div((uint32_t)col, a.dd);
					c.ch = col - x * a.d; c.set(a.col_slot, x);   // ch first: with a wide interleave col_slot IS the channel slot
					res.v[t] = sop(res.v[t], c);
				}
			}
			T *dst = gout + obase + (long long)r * a.ax_os + col0 + c0;
			if (a.vec_out && c0 + VN <= ncl) {
				*(Vec *)dst = res;
			} else {
#pragma unroll
				for (int t = 0; t < VN; t++)
					if (c0 + t < ncl) dst[t] = res.v[t];
			}
		}
	}
}

}  // namespace dsp
