// dct_fast.cuh -- power-of-two fast path of the CTA-level DCT engine (n = r0 * 16^k, r0 in {1,2,4,8,16,32}).
//
// DCT-II  (REDFT10): decimation in time.   scatter-load (Makhoul permutation + digit reversal, one table lookup)
//                    -> contiguous radix-r0 pass -> radix-16 middle passes -> OUTER radix-16 pass fused with the
//                    (k, n-k) post-twiddle; the row kernel stores straight to global memory from registers.
// DCT-III (REDFT01): decimation in frequency, the mirror image: the OUTER radix-16 pass reads (k, n-k) pairs
//                    (row kernel: straight from global memory), pre-twiddles them in registers, then middle
//                    passes, the contiguous radix-r0 pass and a gather-store.
//
// In the outer pass one thread owns butterflies i and M-i (M = n/16): every (k, n-k) pair the twiddle step needs
// is then in its registers, and the two butterflies share one set of twiddles because
// W_n^{(M-i) j} = W_16^j conj(W_n^{i j}) and a W_16^j factor is an index rotation of the 16-point DFT.
//
// Same emulation discipline as dct_core.cuh: within a phase a thread touches only smem slots it owns.
#pragma once
#include "dct_core.cuh"

namespace dsp {

template <class T> DSP_DEV void tw_powers(const C2<T> *tw, int idx, C2<T> *w);

// Twiddle sources of the radix-16 passes.  The default reads the global tables through L1; the ring row kernel
// (dct_ring.cuh) overrides them with tables it keeps in shared memory.
template <class D> struct TwGlobal {
	template <class T> DSP_DEVM void tw_mid(int i, int twsh, C2<T> *w) const { tw_powers<T>((const C2<T> *)static_cast<const D *>(this)->tw, i << twsh, w); }
	template <class T> DSP_DEVM void tw_outer(int i, C2<T> *w) const { tw_powers<T>((const C2<T> *)static_cast<const D *>(this)->tw, i, w); }
	template <class T> DSP_DEVM C2<T> om_at(int i) const { return ldg_c2((const C2<T> *)static_cast<const D *>(this)->om + i); }
};

struct FastDesc : TwGlobal<FastDesc> {
	int n, M;                    // length, n/16
	int r0;                      // contiguous radix (first DIT pass / last DIF pass); 1 = none
	int nmid;                    // radix-16 middle passes between the contiguous and the outer pass
	int npad;                    // sequence stride in smem (complex elements)
	const void *tw;              // C2<T>[n]      W_n^k
	const void *om;              // C2<T>[n/2+1]  (cos, sin)(pi k / 2n)
	const uint16_t *sig;         // [n] padded smem slot of natural index e under the digit reversal
	int poff[4][16];             // poff[q][j] = Pad(j * Lprev_q) for radix-16 pass q (0..nmid-1 middle, nmid = outer)
	FastDiv dHalf;               // divide by M/2+1 (outer-pass units per sequence)
	// accessors shared with FastFixed (the compile-time variant below)
	enum { kFixed = 0, kN = 0 };
	DSP_HDM int N() const { return n; }
	DSP_HDM int Mq() const { return M; }
	DSP_HDM int R0() const { return r0; }
	DSP_HDM int NMID() const { return nmid; }
	DSP_HDM int NPAD() const { return npad; }
	DSP_HDM int PO(int q, int j) const { return poff[q][j]; }
	DSP_DEVM uint32_t divHalf(uint32_t u) const { return fd_div(u, dHalf); }
};

// The same description with the length fixed at compile time (float padding): every smem offset, loop bound and
// division of the passes folds to an immediate.  Kernels are instantiated for the common lengths; the runtime
// FastDesc serves the rest.
template <int LG> struct FastFixedBase {
	const void *tw, *om;
	const uint16_t *sig;
	enum { kFixed = 1, kN = 1 << LG };
	static constexpr int kA = LG - 4;
	static constexpr int kL0raw = kA % 4, kKraw = (kA - kL0raw) / 4;
	static constexpr int kL0 = (kL0raw == 0 && kKraw > 0) ? 4 : ((kL0raw == 1 && kKraw > 0) ? 5 : kL0raw);
	static constexpr int kK = ((kL0raw == 0 || kL0raw == 1) && kKraw > 0) ? kKraw - 1 : kKraw;
	DSP_HDM static constexpr int padc(int e) { return e + (e >> 4) + (e >> 8) + (e >> 12); }
	DSP_HDM constexpr int N() const { return 1 << LG; }
	DSP_HDM constexpr int Mq() const { return 1 << (LG - 4); }
	DSP_HDM constexpr int R0() const { return 1 << kL0; }
	DSP_HDM constexpr int NMID() const { return kK; }
	static constexpr int kNpad0 = (1 << LG) - 1 + (((1 << LG) - 1) >> 4) + (((1 << LG) - 1) >> 8) + (((1 << LG) - 1) >> 12) + 1;
	static constexpr int kNPAD = kNpad0 + ((17 - kNpad0 % 16) % 16);                    // == 1 (mod 16), as the planner pads
	DSP_HDM constexpr int NPAD() const { return kNPAD; }
	DSP_HDM constexpr int PO(int q, int j) const { return padc(j * ((1 << kL0) << (4 * q))); }
	DSP_DEVM uint32_t divHalf(uint32_t u) const { return u / (uint32_t)((1 << (LG - 4)) / 2 + 1); }
};
template <int LG> struct FastFixed : FastFixedBase<LG>, TwGlobal<FastFixed<LG>> {};

// ------------------------------------------------------------------------------------------------ radix 32
template <class T> struct Dft<T, 32> {
	DSP_DEVM static void run(C2<T> *v) {
		// j = 2a + b: 16-point DFTs of the even and odd inputs, twiddle W32^q on the odd half, radix-2 combine
		C2<T> e[16], o[16];
#pragma unroll
		for (int a = 0; a < 16; a++) { e[a] = v[2 * a]; o[a] = v[2 * a + 1]; }
		Dft<T, 16>::run(e);
		Dft<T, 16>::run(o);
		const T c[16] = {(T)1.0, (T)0.980785280403230449126182236134, (T)0.923879532511286756128183189397,
		                 (T)0.831469612302545237078788377618, (T)0.707106781186547524400844362105,
		                 (T)0.555570233019602224742830813949, (T)0.382683432365089771728459984030,
		                 (T)0.195090322016128267848284868477, (T)0.0, (T)-0.195090322016128267848284868477,
		                 (T)-0.382683432365089771728459984030, (T)-0.555570233019602224742830813949,
		                 (T)-0.707106781186547524400844362105, (T)-0.831469612302545237078788377618,
		                 (T)-0.923879532511286756128183189397, (T)-0.980785280403230449126182236134};
		const T s[16] = {(T)0.0, (T)0.195090322016128267848284868477, (T)0.382683432365089771728459984030,
		                 (T)0.555570233019602224742830813949, (T)0.707106781186547524400844362105,
		                 (T)0.831469612302545237078788377618, (T)0.923879532511286756128183189397,
		                 (T)0.980785280403230449126182236134, (T)1.0, (T)0.980785280403230449126182236134,
		                 (T)0.923879532511286756128183189397, (T)0.831469612302545237078788377618,
		                 (T)0.707106781186547524400844362105, (T)0.555570233019602224742830813949,
		                 (T)0.382683432365089771728459984030, (T)0.195090322016128267848284868477};
#pragma unroll
		for (int q = 0; q < 16; q++) {
			const C2<T> t = q == 0 ? o[0] : cmulc(o[q], c[q], -s[q]);     // W32^q = (cos, -sin)(2 pi q / 32)
			v[q] = cadd(e[q], t);
			v[q + 16] = csub(e[q], t);
		}
	}
};

// w[j] = W^{j}, j = 1..15, from w1, w2, w4, w8 (products of depth <= 3)
template <class T>
DSP_DEV void tw_powers(const C2<T> *tw, int idx, C2<T> *w) {
	w[1] = ldg_c2(tw + idx); w[2] = ldg_c2(tw + 2 * idx); w[4] = ldg_c2(tw + 4 * idx); w[8] = ldg_c2(tw + 8 * idx);
	w[3] = cmul(w[1], w[2]); w[5] = cmul(w[1], w[4]); w[6] = cmul(w[2], w[4]); w[7] = cmul(w[3], w[4]);
	w[9] = cmul(w[1], w[8]); w[10] = cmul(w[2], w[8]); w[11] = cmul(w[3], w[8]); w[12] = cmul(w[4], w[8]);
	w[13] = cmul(w[5], w[8]); w[14] = cmul(w[6], w[8]); w[15] = cmul(w[7], w[8]);
}
template <class T> DSP_DEV C2<T> cmul_conj(C2<T> a, C2<T> b) {   // a * conj(b)
	return C2<T>{a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y};
}

// ------------------------------------------------------------------------------------------------ inner passes

template <class T, int R, class F>
DSP_DEV void contig_pass_r(C2<T> *s, int nseq, const F &f, int tid, int nthr) {
	const int lnb = ilog2(f.N() / R);                           // butterflies per sequence (power of two)
	const int total = nseq << lnb;
	for (int g = tid; g < total; g += nthr) {
		const int seq = g >> lnb, b = g & ((1 << lnb) - 1);
		C2<T> *p = s + seq * f.NPAD() + Pad<T>::of(b * R);
		C2<T> v[R];
#pragma unroll
		for (int j = 0; j < R; j++) v[j] = p[Pad<T>::of(j)];      // b*R and j are bit-disjoint (R <= 32): Pad is additive
		Dft<T, R>::run(v);
#pragma unroll
		for (int j = 0; j < R; j++) p[Pad<T>::of(j)] = v[j];
	}
}

template <class T, class F>
DSP_DEV void contig_pass(C2<T> *s, int nseq, const F &f, int t0, int t1, int nthr) {
	if (f.R0() <= 1) return;
	for (int tid = t0; tid < t1; tid++) {
		switch (f.R0()) {
		case 2:  contig_pass_r<T, 2>(s, nseq, f, tid, nthr); break;
		case 4:  contig_pass_r<T, 4>(s, nseq, f, tid, nthr); break;
		case 8:  contig_pass_r<T, 8>(s, nseq, f, tid, nthr); break;
		case 16: contig_pass_r<T, 16>(s, nseq, f, tid, nthr); break;
		case 32: contig_pass_r<T, 32>(s, nseq, f, tid, nthr); break;
		default: break;
		}
	}
	DSP_SYNC();
}

// radix-16 middle pass q (Lprev = r0 * 16^q, L = 16 Lprev).  DIT: twiddle the inputs; DIF: twiddle the outputs.
template <class T, bool DIT, class F>
DSP_DEV void mid_pass(C2<T> *s, int nseq, const F &f, int q, int tid, int nthr) {
	const int lsh = ilog2(f.R0()) + 4 * q;                      // log2 Lprev
	const int lnb = ilog2(f.N()) - 4;                           // log2 (butterflies per sequence)
	const int total = nseq << lnb;
	const int twsh = ilog2(f.N()) - lsh - 4;                    // log2 (n / L)
	for (int g = tid; g < total; g += nthr) {
		const int seq = g >> lnb, b = g & ((1 << lnb) - 1);
		const int blk = b >> lsh, i = b & ((1 << lsh) - 1);
		C2<T> *p = s + seq * f.NPAD() + Pad<T>::of((blk << (lsh + 4)) + i);
		C2<T> v[16], w[16];
		if (i != 0) f.template tw_mid<T>(i, twsh, w);
#pragma unroll
		for (int j = 0; j < 16; j++) v[j] = p[f.PO(q, j)];
		if (DIT && i != 0) {
#pragma unroll
			for (int j = 1; j < 16; j++) v[j] = cmul(v[j], w[j]);
		}
		Dft<T, 16>::run(v);
		if (!DIT && i != 0) {
#pragma unroll
			for (int j = 1; j < 16; j++) v[j] = cmul(v[j], w[j]);
		}
#pragma unroll
		for (int j = 0; j < 16; j++) p[f.PO(q, j)] = v[j];
	}
}

// ------------------------------------------------------------------------------------------------ half-sample phases
// (cos, sin)(pi k / 2n) for the k the outer pass meets, derived from one table entry:
//   k = i + M m      -> angle(i) + pi m / 32          k = M - i + M m -> pi (m+1) / 32 - angle(i)
//   k = M m          -> pi m / 32 (constant)          k = M/2 + M m   -> pi (2m+1) / 64 (constant)
template <class T> struct OmConst {
	DSP_DEVM static C2<T> e(int m) {                            // (cos, sin)(pi m / 32), m = 0..8
		const T c[9] = {(T)1.0, (T)0.995184726672196886244836953109, (T)0.980785280403230449126182236134, (T)0.95694033573220886493579788698, (T)0.923879532511286756128183189397, (T)0.88192126434835502971275686366, (T)0.831469612302545237078788377618, (T)0.773010453362736960810906609758, (T)0.707106781186547524400844362105};
		const T s[9] = {(T)0.0, (T)0.0980171403295606019941955638886, (T)0.195090322016128267848284868477, (T)0.290284677254462367636192375817, (T)0.38268343236508977172845998403, (T)0.471396736825997648556387625905, (T)0.555570233019602224742830813949, (T)0.634393284163645498215171613225, (T)0.707106781186547524400844362105};
		return C2<T>{c[m], s[m]};
	}
	DSP_DEVM static C2<T> h(int m) {                            // (cos, sin)(pi (2m+1) / 64), m = 0..7
		const T c[8] = {(T)0.998795456205172392714771604759, (T)0.989176509964780973451673738016, (T)0.970031253194543992603984207286, (T)0.9415440651830207784125094026, (T)0.903989293123443331586200297231, (T)0.857728610000272069902269984285, (T)0.803207531480644909806676512963, (T)0.740951125354959091175616897495};
		const T s[8] = {(T)0.0490676743274180142549549769427, (T)0.146730474455361751658850129647, (T)0.242980179903263889948274162077, (T)0.336889853392220050689253212619, (T)0.427555093430282094320966856889, (T)0.514102744193221726593693838969, (T)0.59569930449243334346703652883, (T)0.671558954847018400625376850427};
		return C2<T>{c[m], s[m]};
	}
};
template <class T> DSP_DEV C2<T> om_plus(C2<T> wi, C2<T> e) { return C2<T>{wi.x * e.x - wi.y * e.y, wi.y * e.x + wi.x * e.y}; }   // angle(i) + angle(e)
template <class T> DSP_DEV C2<T> om_minus(C2<T> wi, C2<T> e) { return C2<T>{e.x * wi.x + e.y * wi.y, e.y * wi.x - e.x * wi.y}; }  // angle(e) - angle(i)

// ------------------------------------------------------------------------------------------------ butterfly storage
// Where the outer pass finds (DCT-II) or leaves (DCT-III) element j of butterfly i.
// SmemBf: slot Pad(i) + Pad(j M) of a sequence in shared memory.
template <class T, class F> struct SmemBf {
	C2<T> *base;
	const F *f;                  // slot of element j: Pad(i) + PO(nmid, j) = Pad(i) + Pad(j * M)
	struct Row {
		C2<T> *p; const F *f;
		DSP_DEVM C2<T> get(int j) const { return p[f->PO(f->NMID(), j)]; }
		DSP_DEVM void put(int j, C2<T> v) const { p[f->PO(f->NMID(), j)] = v; }
	};
	DSP_DEVM Row row(int i) const { return Row{base + Pad<T>::of(i), f}; }
	DSP_DEVM Row rowb(int i) const { return row(i); }       // second butterfly of a pair (register-staged storage tells them apart)
};
// GlobBf: row (j M + i) of a [16 M][...] global scratch, one complex (= one column pair) per row
template <class T> struct GlobBf {
	C2<T> *base;                 // scratch + column pair
	long long rs;                // row stride in complex elements
	int M;
	struct Row {
		C2<T> *p; long long js;
		DSP_DEVM C2<T> get(int j) const { return p[j * js]; }
		DSP_DEVM void put(int j, C2<T> v) const { p[j * js] = v; }
	};
	DSP_DEVM Row row(int i) const { return Row{base + (long long)i * rs, (long long)M * rs}; }
	DSP_DEVM Row rowb(int i) const { return row(i); }
};

// ------------------------------------------------------------------------------------------------ DCT-II outer pass
// pair (k, n-k), k <= n/2: Z = spectrum value at k, Y at n-k, w = (cos, sin)(pi k / 2n).
template <class T, class Sink>
DSP_DEV void dct2_pair(C2<T> w, int k, int n, C2<T> z, C2<T> y, bool self, Sink &sink) {
	const T ar = z.x + y.x, ai = z.y - y.y, br = z.y + y.y, bi = y.x - z.x;
	sink.put(k, w.x * ar + w.y * ai, w.x * br + w.y * bi);
	if (!self) sink.put(n - k, w.y * ar - w.x * ai, w.y * br - w.x * bi);
}

template <class T, class Bf, class Sink, class F>
DSP_DEV void dct2_outer_unit(const Bf &bf, const F &f, int i, Sink &sink) {
	const int n = f.N(), M = f.Mq();
	C2<T> a[16], w[16];
	if (i == 0) {
		const typename Bf::Row p = bf.row(0);
#pragma unroll
		for (int j = 0; j < 16; j++) a[j] = p.get(j);
		Dft<T, 16>::run(a);                                       // a[m] = Z[M m]
		dct2_pair<T>(OmConst<T>::e(0), 0, n, a[0], a[0], true, sink);
#pragma unroll
		for (int m = 1; m < 8; m++) dct2_pair<T>(OmConst<T>::e(m), M * m, n, a[m], a[16 - m], false, sink);
		dct2_pair<T>(OmConst<T>::e(8), 8 * M, n, a[8], a[8], true, sink);
		return;
	}
	f.template tw_outer<T>(i, w);
	if (2 * i == M) {
		const typename Bf::Row p = bf.row(i);
#pragma unroll
		for (int j = 0; j < 16; j++) a[j] = p.get(j);
#pragma unroll
		for (int j = 1; j < 16; j++) a[j] = cmul(a[j], w[j]);
		Dft<T, 16>::run(a);                                       // a[m] = Z[M/2 + M m]; partner of m is 15-m
#pragma unroll
		for (int m = 0; m < 8; m++) dct2_pair<T>(OmConst<T>::h(m), i + M * m, n, a[m], a[15 - m], false, sink);
		return;
	}
	const C2<T> wi = f.template om_at<T>(i);
	C2<T> b[16];
	{
		const typename Bf::Row p = bf.row(i);
#pragma unroll
		for (int j = 0; j < 16; j++) a[j] = p.get(j);
	}
	{
		const typename Bf::Row p = bf.rowb(M - i);
#pragma unroll
		for (int j = 0; j < 16; j++) b[j] = p.get(j);
	}
#pragma unroll
	for (int j = 1; j < 16; j++) { a[j] = cmul(a[j], w[j]); b[j] = cmul_conj(b[j], w[j]); }
	Dft<T, 16>::run(a);                                           // a[m] = Z[i + M m]
	Dft<T, 16>::run(b);                                           // b[(m+1)&15] = Z[M-i + M m]
#pragma unroll
	for (int m = 0; m < 8; m++) {
		dct2_pair<T>(om_plus<T>(wi, OmConst<T>::e(m)), i + M * m, n, a[m], b[(16 - m) & 15], false, sink);               // partner: Z[M-i + M(15-m)]
		dct2_pair<T>(om_minus<T>(wi, OmConst<T>::e(m + 1)), M - i + M * m, n, b[(m + 1) & 15], a[15 - m], false, sink);  // partner: Z[i + M(15-m)]
	}
}

// ------------------------------------------------------------------------------------------------ DCT-III outer pass
// pair (k, n-k), k <= n/2: xk = (XA[k], XB[k]), xn = (XA[n-k], XB[n-k]).  Produces conj(Z[k]) and conj(Z[n-k]).
template <class T>
DSP_DEV void dct3_pair(C2<T> w, C2<T> xk, C2<T> xn, C2<T> &wk, C2<T> &wn) {
	const T pa = w.x * xk.x + w.y * xn.x, qa = w.y * xk.x - w.x * xn.x;
	const T pb = w.x * xk.y + w.y * xn.y, qb = w.y * xk.y - w.x * xn.y;
	wk = C2<T>{pa - qb, -(qa + pb)};
	wn = C2<T>{pa + qb, -(pb - qa)};
}

template <class T, class Bf, class Source, class F>
DSP_DEV void dct3_outer_unit(const Bf &bf, const F &f, int i, Source &src) {
	const int M = f.Mq();
	C2<T> a[16], w[16];
	if (i == 0) {
		// k = M j; partner of j is 16-j; j = 0 and j = 8 are their own partners
		C2<T> x[16];
#pragma unroll
		for (int j = 0; j < 16; j++) x[j] = src.get(M * j);
		a[0] = C2<T>{x[0].x, -x[0].y};
#pragma unroll
		for (int j = 1; j < 8; j++) dct3_pair<T>(OmConst<T>::e(j), x[j], x[16 - j], a[j], a[16 - j]);
		{ C2<T> dummy; dct3_pair<T>(OmConst<T>::e(8), x[8], x[8], a[8], dummy); }
		Dft<T, 16>::run(a);
		const typename Bf::Row p = bf.row(0);
#pragma unroll
		for (int j = 0; j < 16; j++) p.put(j, a[j]);
		return;
	}
	if (2 * i == M) {
		C2<T> x[16];
#pragma unroll
		for (int j = 0; j < 16; j++) x[j] = src.get(i + M * j);
#pragma unroll
		for (int j = 0; j < 8; j++) dct3_pair<T>(OmConst<T>::h(j), x[j], x[15 - j], a[j], a[15 - j]);
		Dft<T, 16>::run(a);
		f.template tw_outer<T>(i, w);
#pragma unroll
		for (int j = 1; j < 16; j++) a[j] = cmul(a[j], w[j]);
		const typename Bf::Row p = bf.row(i);
#pragma unroll
		for (int j = 0; j < 16; j++) p.put(j, a[j]);
		return;
	}
	const C2<T> wi = f.template om_at<T>(i);
	C2<T> b[16];
	{
		C2<T> xa[16], xb[16];
#pragma unroll
		for (int j = 0; j < 16; j++) { xa[j] = src.get(i + M * j); xb[j] = src.get(M - i + M * j); }
		// butterfly i input j (k = i + M j) pairs with butterfly M-i input 15-j.  b is built rotated by one:
		// b[(j+1)&15] = input j of butterfly M-i, which turns the W_16^m output factor into the plain DFT.
#pragma unroll
		for (int j = 0; j < 8; j++) {
			dct3_pair<T>(om_plus<T>(wi, OmConst<T>::e(j)), xa[j], xb[15 - j], a[j], b[(16 - j) & 15]);
			dct3_pair<T>(om_minus<T>(wi, OmConst<T>::e(j + 1)), xb[j], xa[15 - j], b[(j + 1) & 15], a[15 - j]);
		}
	}
	Dft<T, 16>::run(a);
	Dft<T, 16>::run(b);
	f.template tw_outer<T>(i, w);
#pragma unroll
	for (int j = 1; j < 16; j++) { a[j] = cmul(a[j], w[j]); b[j] = cmul_conj(b[j], w[j]); }
	{
		const typename Bf::Row p = bf.row(i);
#pragma unroll
		for (int j = 0; j < 16; j++) p.put(j, a[j]);
	}
	{
		const typename Bf::Row p = bf.rowb(M - i);
#pragma unroll
		for (int j = 0; j < 16; j++) p.put(j, b[j]);
	}
}

// sources / sinks ------------------------------------------------------------------------------------------
// smem natural order: slot Pad(k) holds (XA[k], XB[k])
template <class T> struct SmemNat {
	C2<T> *base;
	DSP_DEVM void put(int k, T xa, T xb) { base[Pad<T>::of(k)] = C2<T>{xa, xb}; }
	DSP_DEVM C2<T> get(int k) const { return base[Pad<T>::of(k)]; }
};
// two global lines (rows A and B of the pair), element k of channel ch at [k*d + ch]
// DD > 0: the interleave d == DD and both lines present, known at compile time -- with a fixed-length F every k is
// "i + constant", so the accesses become [base + immediate] instead of a 64-bit address computation each
template <class T, class Op, int DD = 0> struct GlobalRows {
	T *pa, *pb;                  // line bases (+ channel); pb = nullptr when the pair has no second line
	int d, ax_slot;
	Coord ca, cb;
	const Op *op;
	DSP_DEVM void put(int k, T xa, T xb) {
		if (DD) { pa[k * DD] = (*op)(xa, ca); pb[k * DD] = (*op)(xb, cb); return; }
		ca.set(ax_slot, k); cb.set(ax_slot, k);
		pa[k * d] = (*op)(xa, ca);
		if (pb) pb[k * d] = (*op)(xb, cb);
	}
	DSP_DEVM C2<T> get(int k) {
		if (DD) return C2<T>{(*op)(pa[k * DD], ca), (*op)(pb[k * DD], cb)};
		ca.set(ax_slot, k); cb.set(ax_slot, k);
		const T xa = (*op)(pa[k * d], ca);
		const T xb = pb ? (*op)(pb[k * d], cb) : (T)0;
		return C2<T>{xa, xb};
	}
};


// The outer-pass units of `nseq` sequences: per sequence i = 0 .. M/2 (i = 0 and i = M/2 are half-cost special
// units, the others butterfly pairs i, M-i), flat over the CTA's threads.  One call site: fn is large.
// (Tried: general units first, special units on lane 0 of successive warps so that no warp serialises three code
// paths -- no gain at n = 8192 with two CTAs per SM, a loss at n = 1024 where it adds a second round.)
template <class Fn>
DSP_DEV void for_outer_units(int nseq, int M, int tid, int nthr, const Fn &fn) {
	const uint32_t upseq = (uint32_t)(M / 2 + 1);
	for (uint32_t u = (uint32_t)tid; u < (uint32_t)nseq * upseq; u += (uint32_t)nthr) {
		const uint32_t seq = u / upseq;
		fn((int)seq, (int)(u - seq * upseq));
	}
}

// ------------------------------------------------------------------------------------------------ row pass (fast)
// Moves the CTA's lines between global memory and smem in vector groups, `UNR` groups per thread in flight.
// FWD (DCT-II scatter-load):  global -> lop -> s[sig[makhoul(x)]]
// !FWD (DCT-III gather-store): s[sig[makhoul(x)]] -> (re, -im) -> sop -> global
template <class T, bool FWD, class Op, class F>
DSP_DEV void row_move(const RowArgs &a, const F &f, const Op &op, int line0, int nl, int tid, int nthr, C2<T> *s) {
	typedef typename VecOf<T>::type Vec;
	const int VN = VecOf<T>::N;
	const int UNR = 4;
	const T *gin = (const T *)a.in;
	T *gout = (T *)a.out;
	const int n = f.N(), d = a.d;
	const int llen = n * d;
	const int gpl = (llen + VN - 1) / VN;
	const int npairs = (nl + 1) / 2;
	const bool vec = (FWD ? a.vec_in : a.vec_out) && (llen % VN) == 0;
	const PairThreads pt(gpl, tid, nthr);
	for (int g = pt.grp; g < npairs; g += pt.ngrp) {
		const int la = line0 + 2 * g;
		const bool hasb = (2 * g + 1) < nl;
		Coord ca = {0, 0, 0, 0, 0}, cb = {0, 0, 0, 0, 0};
		long long ia, oa, ib = 0, ob = 0;
		outer_decode(a.o, (uint32_t)la, ia, oa, ca);
		if (hasb) outer_decode(a.o, (uint32_t)la + 1, ib, ob, cb);
		const T *pa = gin + ia, *pb = gin + ib;
		T *qa = gout + oa, *qb = gout + ob;
		C2<T> *sg = s + g * d * f.NPAD();
		for (int q0 = pt.ql; q0 < gpl; q0 += pt.gp2 * UNR) {
			T va[UNR][VecOf<T>::N], vb[UNR][VecOf<T>::N];
			if (FWD) {
				// ---- all loads of the batch first
#pragma unroll
				for (int u = 0; u < UNR; u++) {
					const int e0 = (q0 + u * pt.gp2) * VN;
					if (q0 + u * pt.gp2 < gpl) {
						if (vec) {
							const Vec ta = ldg_stream((const Vec *)(pa + e0));
#pragma unroll
							for (int t = 0; t < VN; t++) va[u][t] = ta.v[t];
							if (hasb) {
								const Vec tb = ldg_stream((const Vec *)(pb + e0));
#pragma unroll
								for (int t = 0; t < VN; t++) vb[u][t] = tb.v[t];
							}
						} else if (a.in_u8) {
							const unsigned char *g8 = (const unsigned char *)a.in;
#pragma unroll
							for (int t = 0; t < VN; t++) {
								va[u][t] = (e0 + t < llen) ? (T)g8[ia + e0 + t] : (T)0;
								vb[u][t] = (hasb && e0 + t < llen) ? (T)g8[ib + e0 + t] : (T)0;
							}
						} else {
#pragma unroll
							for (int t = 0; t < VN; t++) {
								va[u][t] = (e0 + t < llen) ? pa[e0 + t] : (T)0;
								vb[u][t] = (hasb && e0 + t < llen) ? pb[e0 + t] : (T)0;
							}
						}
					}
				}
			}
#pragma unroll
			for (int u = 0; u < UNR; u++) {
				const int e0 = (q0 + u * pt.gp2) * VN;
				if (q0 + u * pt.gp2 < gpl) {
#pragma unroll
					for (int t = 0; t < VN; t++) {
						const int e = e0 + t;
						if (e < llen) {
							int x = e, ch = 0;
							if (d != 1) { x = (int)fd_div((uint32_t)e, a.dd); ch = e - x * d; }
							ca.set(a.ax_slot, x); ca.ch = ch;
							cb.set(a.ax_slot, x); cb.ch = ch;
							C2<T> *slot = sg + ch * f.NPAD() + (int)DSP_LDG(f.sig + makhoul(x, n));
							if (FWD) {
								*slot = C2<T>{op(va[u][t], ca), hasb ? op(vb[u][t], cb) : (T)0};
							} else {
								const C2<T> z = *slot;
								va[u][t] = op(z.x, ca);
								vb[u][t] = op(-z.y, cb);
							}
						}
					}
					if (!FWD) {
						if (vec) {
							Vec ra, rb;
#pragma unroll
							for (int t = 0; t < VN; t++) { ra.v[t] = va[u][t]; rb.v[t] = vb[u][t]; }
							*(Vec *)(qa + e0) = ra;
							if (hasb) *(Vec *)(qb + e0) = rb;
						} else if (a.out_u8) {
							unsigned char *o8 = (unsigned char *)a.out;
#pragma unroll
							for (int t = 0; t < VN; t++)
								if (e0 + t < llen) {
									o8[oa + e0 + t] = to_u8(va[u][t]);
									if (hasb) o8[ob + e0 + t] = to_u8(vb[u][t]);
								}
						} else {
#pragma unroll
							for (int t = 0; t < VN; t++)
								if (e0 + t < llen) {
									qa[e0 + t] = va[u][t];
									if (hasb) qb[e0 + t] = vb[u][t];
								}
						}
					}
				}
			}
		}
	}
}

// The common case of row_move_planar4 (below) with all addressing hoisted: lines at a single stride, both lines of
// every pair present, no coordinates wanted by the op.  Per group of 8 samples: 2 x LDG.128 (or STG.128), two slot
// lookups, 4 x STS.64 (LDS.64) -- the general version spends ~100 instructions on the same work (ncu, profiles/).
template <class T, bool FWD, bool FULL, class Op, class F>
DSP_DEV void row_move_planar4_lean(const RowArgs &a, const F &f, const Op &op, int line0, int npairs, int tid, int nthr, C2<T> *s) {
	typedef typename VecOf<T>::type Vec;
	const int UNR = 8;
	const int n = f.N(), gpl = n >> 2;
	const int padM = f.PO(f.NMID(), 1);
	const Coord c0 = {0, 0, 0, 0, 0};
	const PairThreads pt(gpl, tid, nthr);                    // short lines: several pairs at a time
	const int ql = pt.ql, gp2 = pt.gp2;
	const uint16_t *sigA = f.sig + 2 * ql;                   // sig[2q],     q = ql + k gp2
	const uint16_t *sigB = f.sig + (n - 2) - 2 * ql;         // sig[n-2-2q]
	for (int g = pt.grp; g < npairs; g += pt.ngrp) {
		const long long l = line0 + 2 * g;
		const Vec *pa = (const Vec *)((const T *)a.in + l * a.ls_in) + ql;
		const Vec *pb = (const Vec *)((const T *)a.in + (l + 1) * a.ls_in) + ql;
		Vec *qa = (Vec *)((T *)a.out + l * a.ls_out) + ql;
		Vec *qb = (Vec *)((T *)a.out + (l + 1) * a.ls_out) + ql;
		C2<T> *sg = s + g * f.NPAD();
		for (int q0 = 0; q0 < gpl; q0 += gp2 * UNR) {
			Vec ta[UNR], tb[UNR];
			if (FWD) {
#pragma unroll
				for (int u = 0; u < UNR; u++) {
					const int qq = q0 + u * gp2;
					if (FULL || qq + ql < gpl) { ta[u] = ldg_stream(pa + qq); tb[u] = ldg_stream(pb + qq); }
				}
			}
#pragma unroll
			for (int u = 0; u < UNR; u++) {
				const int qq = q0 + u * gp2;
				if (FULL || qq + ql < gpl) {
					C2<T> *b0 = sg + (int)DSP_LDG(sigA + 2 * qq), *b1 = sg + (int)DSP_LDG(sigB - 2 * qq);
					if (FWD) {
						b0[0]    = C2<T>{op(ta[u].v[0], c0), op(tb[u].v[0], c0)};        // x = 4q
						b1[padM] = C2<T>{op(ta[u].v[1], c0), op(tb[u].v[1], c0)};        // x = 4q + 1
						b0[padM] = C2<T>{op(ta[u].v[2], c0), op(tb[u].v[2], c0)};        // x = 4q + 2
						b1[0]    = C2<T>{op(ta[u].v[3], c0), op(tb[u].v[3], c0)};        // x = 4q + 3
					} else {
						const C2<T> z0 = b0[0], z1 = b1[padM], z2 = b0[padM], z3 = b1[0];
						Vec ra, rb;
						ra.v[0] = op(z0.x, c0); ra.v[1] = op(z1.x, c0); ra.v[2] = op(z2.x, c0); ra.v[3] = op(z3.x, c0);
						rb.v[0] = op(-z0.y, c0); rb.v[1] = op(-z1.y, c0); rb.v[2] = op(-z2.y, c0); rb.v[3] = op(-z3.y, c0);
						qa[qq] = ra;
						qb[qq] = rb;
					}
				}
			}
		}
	}
}

template <class T, bool FWD, class Op, class F>
DSP_DEV void row_move_planar4_pick(const RowArgs &a, const F &f, const Op &op, int line0, int npairs, int tid, int nthr, C2<T> *s) {
	if (((f.N() >> 2) % (nthr * 8)) == 0) row_move_planar4_lean<T, FWD, true, Op, F>(a, f, op, line0, npairs, tid, nthr, s);
	else row_move_planar4_lean<T, FWD, false, Op, F>(a, f, op, line0, npairs, tid, nthr, s);
}

// planar float lines (d == 1, 16-byte access legal): one vector group = x in [4q, 4q+4) of lines A and B =
// elements 2q, 2q+1 (even x) and n-2-2q, n-1-2q (odd x) of the permuted sequence.  sig[e+1] = sig[e] + Pad(M)
// for even e, so two table lookups place (or fetch) all four complex values.
template <class T, bool FWD, class Op, class F>
DSP_DEV void row_move_planar4(const RowArgs &a, const F &f, const Op &op, int line0, int nl, int tid, int nthr, C2<T> *s) {
	typedef typename VecOf<T>::type Vec;
	const int UNR = 8;
	const T *gin = (const T *)a.in;
	T *gout = (T *)a.out;
	const int n = f.N();
	const int lgq = ilog2(n) - 2;                             // log2 (vector groups per line)
	const int npairs = (nl + 1) / 2;
	if (a.simple && !Op::kNeedsCoord && !(nl & 1)) {
		row_move_planar4_pick<T, FWD, Op, F>(a, f, op, line0, npairs, tid, nthr, s);
		return;
	}
	const int total = npairs << lgq;
	const int padM = f.PO(f.NMID(), 1);
	for (int i0 = tid; i0 < total; i0 += nthr * UNR) {
		Vec ta[UNR], tb[UNR];
		if (FWD) {
#pragma unroll
			for (int u = 0; u < UNR; u++) {
				const int idx = i0 + u * nthr;
				if (idx < total) {
					const int g = idx >> lgq, q = idx & ((1 << lgq) - 1);
					const int la = line0 + 2 * g;
					long long ia, ib, oa;
					Coord c = {0, 0, 0, 0, 0};
					if (a.simple && !Op::kNeedsCoord) { ia = la * a.ls_in; ib = ia + a.ls_in; }
					else { outer_decode(a.o, (uint32_t)la, ia, oa, c); ib = 0; if (2 * g + 1 < nl) outer_decode(a.o, (uint32_t)la + 1, ib, oa, c); }
					ta[u] = ldg_stream((const Vec *)(gin + ia + 4 * q));
					if (2 * g + 1 < nl) tb[u] = ldg_stream((const Vec *)(gin + ib + 4 * q));
					else { tb[u].v[0] = 0; tb[u].v[1] = 0; tb[u].v[2] = 0; tb[u].v[3] = 0; }
				}
			}
		}
#pragma unroll
		for (int u = 0; u < UNR; u++) {
			const int idx = i0 + u * nthr;
			if (idx < total) {
				const int g = idx >> lgq, q = idx & ((1 << lgq) - 1);
				const int la = line0 + 2 * g;
				const bool hasb = 2 * g + 1 < nl;
				Coord ca = {0, 0, 0, 0, 0}, cb = {0, 0, 0, 0, 0};
				long long ia, ib = 0, oa, ob = 0;
				if (a.simple && !Op::kNeedsCoord) { oa = la * a.ls_out; ob = oa + a.ls_out; }
				else { outer_decode(a.o, (uint32_t)la, ia, oa, ca); if (hasb) outer_decode(a.o, (uint32_t)la + 1, ib, ob, cb); }
				C2<T> *sg = s + g * f.NPAD();
				const int s0 = (int)DSP_LDG(f.sig + 2 * q), s1 = (int)DSP_LDG(f.sig + n - 2 - 2 * q);
				C2<T> *slot[4] = {sg + s0, sg + s1 + padM, sg + s0 + padM, sg + s1};       // x = 4q, 4q+1, 4q+2, 4q+3
				if (FWD) {
#pragma unroll
					for (int t = 0; t < 4; t++) {
						ca.set(a.ax_slot, 4 * q + t); cb.set(a.ax_slot, 4 * q + t);
						*slot[t] = C2<T>{op(ta[u].v[t], ca), hasb ? op(tb[u].v[t], cb) : (T)0};
					}
				} else {
					Vec ra, rb;
#pragma unroll
					for (int t = 0; t < 4; t++) {
						const C2<T> z = *slot[t];
						ca.set(a.ax_slot, 4 * q + t); cb.set(a.ax_slot, 4 * q + t);
						ra.v[t] = op(z.x, ca);
						rb.v[t] = op(-z.y, cb);
					}
					*(Vec *)(gout + oa + 4 * q) = ra;
					if (hasb) *(Vec *)(gout + ob + 4 * q) = rb;
				}
			}
		}
	}
}

// channel-interleaved float lines (D = 2..4 channels, 16-byte access legal, n % 4 == 0): one group = x in [4q, 4q+4)
// of lines A and B = D vectors per line; the slot pattern per channel is the planar one.
template <class T, int D, bool FWD, class Op, class F, bool SIMPLE = false>
DSP_DEV void row_move_inter4(const RowArgs &a, const F &f, const Op &op, int line0, int nl, int tid, int nthr, C2<T> *s) {
	typedef typename VecOf<T>::type Vec;
	const T *gin = (const T *)a.in;
	T *gout = (T *)a.out;
	const int n = f.N();
	const int lgq = ilog2(n) - 2;                             // log2 (groups per line)
	const int npairs = (nl + 1) / 2;
	const int total = npairs << lgq;
	const int padM = f.PO(f.NMID(), 1);
	for (int idx = tid; idx < total; idx += nthr) {
		const int g = idx >> lgq, q = idx & ((1 << lgq) - 1);
		const int la = line0 + 2 * g;
		const bool hasb = SIMPLE || 2 * g + 1 < nl;
		Coord ca = {0, 0, 0, 0, 0}, cb = {0, 0, 0, 0, 0};
		long long ia, ib = 0, oa, ob = 0;
		if (SIMPLE || (a.simple && !Op::kNeedsCoord)) { ia = la * a.ls_in; ib = ia + a.ls_in; oa = la * a.ls_out; ob = oa + a.ls_out; }
		else { outer_decode(a.o, (uint32_t)la, ia, oa, ca); if (hasb) outer_decode(a.o, (uint32_t)la + 1, ib, ob, cb); }
		T va[4 * D], vb[4 * D];
		if (FWD) {
#pragma unroll
			for (int v = 0; v < D; v++) {
				const Vec ta = ldg_stream((const Vec *)(gin + ia + 4 * D * q + 4 * v));
#pragma unroll
				for (int t = 0; t < 4; t++) va[4 * v + t] = ta.v[t];
				if (hasb) {
					const Vec tb = ldg_stream((const Vec *)(gin + ib + 4 * D * q + 4 * v));
#pragma unroll
					for (int t = 0; t < 4; t++) vb[4 * v + t] = tb.v[t];
				} else {
#pragma unroll
					for (int t = 0; t < 4; t++) vb[4 * v + t] = 0;
				}
			}
		}
		const int s0 = (int)DSP_LDG(f.sig + 2 * q), s1 = (int)DSP_LDG(f.sig + n - 2 - 2 * q);
		const int slot[4] = {s0, s1 + padM, s0 + padM, s1};       // x = 4q, 4q+1, 4q+2, 4q+3
		C2<T> *sg = s + g * D * f.NPAD();
#pragma unroll
		for (int t = 0; t < 4; t++) {
#pragma unroll
			for (int ch = 0; ch < D; ch++) {
				ca.set(a.ax_slot, 4 * q + t); ca.ch = ch; cb.set(a.ax_slot, 4 * q + t); cb.ch = ch;
				C2<T> *p = sg + ch * f.NPAD() + slot[t];
				if (FWD) *p = C2<T>{op(va[t * D + ch], ca), hasb ? op(vb[t * D + ch], cb) : (T)0};
				else { const C2<T> z = *p; va[t * D + ch] = op(z.x, ca); vb[t * D + ch] = op(-z.y, cb); }
			}
		}
		if (!FWD) {
#pragma unroll
			for (int v = 0; v < D; v++) {
				Vec ra, rb;
#pragma unroll
				for (int t = 0; t < 4; t++) { ra.v[t] = va[4 * v + t]; rb.v[t] = vb[4 * v + t]; }
				*(Vec *)(gout + oa + 4 * D * q + 4 * v) = ra;
				if (hasb) *(Vec *)(gout + ob + 4 * D * q + 4 * v) = rb;
			}
		}
	}
}

template <class T, bool FWD, class Op, class F>
DSP_DEV void row_move_any(const RowArgs &a, const F &f, const Op &op, int line0, int nl, int tid, int nthr, C2<T> *s) {
	const bool al = (FWD ? a.vec_in : a.vec_out) && f.N() >= 4;
	if (sizeof(T) == 4 && a.d == 1 && al) row_move_planar4<T, FWD, Op>(a, f, op, line0, nl, tid, nthr, s);
	else if (sizeof(T) == 4 && a.d == 3 && al) row_move_inter4<T, 3, FWD, Op>(a, f, op, line0, nl, tid, nthr, s);
	else if (sizeof(T) == 4 && a.d == 2 && al) row_move_inter4<T, 2, FWD, Op>(a, f, op, line0, nl, tid, nthr, s);
	else if (sizeof(T) == 4 && a.d == 4 && al) row_move_inter4<T, 4, FWD, Op>(a, f, op, line0, nl, tid, nthr, s);
	else row_move<T, FWD, Op>(a, f, op, line0, nl, tid, nthr, s);
}

// DD > 0 (chosen at launch, lean ops only): interleave d == DD (1 planar, 3 RGB) at compile time, lines at one
// stride, every CTA owns whole line pairs, 16-byte access legal -- the moves and the fused outer pass then run with
// all addressing hoisted.
template <class T, bool FWD, class LoadOp, class StoreOp, class F, int DD = 0>
DSP_DEV void cta_row_fast(const RowArgs &a, const F &f, const LoadOp &lop, const StoreOp &sop, int cta, int t0, int t1,
                          int nthr, C2<T> *s) {
	const T *gin = (const T *)a.in;
	T *gout = (T *)a.out;
	const int d = DD ? DD : a.d;
	const int line0 = cta * a.lines_per_cta;
	int nl = a.nlines - line0;
	if (nl > a.lines_per_cta) nl = a.lines_per_cta;
	const int npairs = (nl + 1) / 2;
	const int nseq = npairs * d;

	// just-in-time L2 prefetch: the lines of the CTA that will take this SM's place (about one resident wave ahead)
	if (a.pf_dist > 0 && a.simple) {
		const int pl0 = (cta + a.pf_dist) * a.lines_per_cta;
		const int lbytes = f.N() * d * (int)sizeof(T);
		if (a.vec_in && (lbytes & 15) == 0) {
			// 16-byte aligned lines: bulk prefetches of up to 8 KB, one thread each
			const int chunks = (lbytes + 8191) / 8192;
			for (int tid = t0; tid < t1; tid++)
				for (int i = tid; i < a.lines_per_cta * chunks; i += nthr) {
					const int l = pl0 + i / chunks, c = i % chunks;
					const int nb = lbytes - c * 8192 < 8192 ? lbytes - c * 8192 : 8192;
					if (l < a.nlines) prefetch_l2_bulk((const char *)gin + ((long long)l * a.ls_in) * (long long)sizeof(T) + (long long)c * 8192, (uint32_t)nb);
				}
		} else {
			const int lines128 = (lbytes + 127) / 128;
			for (int tid = t0; tid < t1; tid++)
				for (int i = tid; i < a.lines_per_cta * lines128; i += nthr) {
					const int l = pl0 + i / lines128;
					if (l < a.nlines) prefetch_l2((const char *)gin + ((long long)l * a.ls_in) * (long long)sizeof(T) + (long long)(i % lines128) * 128);
				}
		}
	}

	if (FWD) {
		for (int tid = t0; tid < t1; tid++) {
			if (DD == 1) row_move_planar4_pick<T, true, LoadOp>(a, f, lop, line0, npairs, tid, nthr, s);
			else if (DD > 1) row_move_inter4<T, (DD > 1 ? DD : 2), true, LoadOp, F, true>(a, f, lop, line0, nl, tid, nthr, s);
			else row_move_any<T, true, LoadOp>(a, f, lop, line0, nl, tid, nthr, s);
		}
		DSP_SYNC();
		contig_pass<T>(s, nseq, f, t0, t1, nthr);
		for (int q = 0; q < f.NMID(); q++) {
			for (int tid = t0; tid < t1; tid++) mid_pass<T, true>(s, nseq, f, q, tid, nthr);
			DSP_SYNC();
		}
		// ---- outer pass + post-twiddle + direct global store
		for (int tid = t0; tid < t1; tid++) {
			if (DD) {
				for_outer_units(nseq, f.Mq(), tid, nthr, [&](int seq, int i) {
					GlobalRows<T, StoreOp, DD> sink;
					const int g = DD == 1 ? seq : seq / (DD ? DD : 1), ch = DD == 1 ? 0 : seq - g * DD;
					sink.ca = Coord{0, 0, 0, 0, 0}; sink.cb = sink.ca;
					sink.pa = gout + (long long)(line0 + 2 * g) * a.ls_out + ch; sink.pb = sink.pa + a.ls_out;
					sink.d = DD; sink.ax_slot = a.ax_slot; sink.op = &sop;
					dct2_outer_unit<T>(SmemBf<T, F>{s + seq * f.NPAD(), &f}, f, i, sink);
				});
				continue;
			}
			for_outer_units(nseq, f.Mq(), tid, nthr, [&](int seq, int i) {
				const int g = seq / d, ch = seq - g * d;
				const int la = line0 + 2 * g;
				const bool hasb = (2 * g + 1) < nl;
				GlobalRows<T, StoreOp> sink;
				long long ia, oa, ib = 0, ob = 0;
				sink.ca = Coord{0, 0, 0, 0, 0}; sink.cb = Coord{0, 0, 0, 0, 0};
				outer_decode(a.o, (uint32_t)la, ia, oa, sink.ca);
				if (hasb) outer_decode(a.o, (uint32_t)la + 1, ib, ob, sink.cb);
				sink.ca.ch = ch; sink.cb.ch = ch;
				sink.pa = gout + oa + ch; sink.pb = hasb ? gout + ob + ch : (T *)0;
				sink.d = d; sink.ax_slot = a.ax_slot; sink.op = &sop;
				dct2_outer_unit<T>(SmemBf<T, F>{s + seq * f.NPAD(), &f}, f, i, sink);
			});
		}
		return;
	}

	// ---- DCT-III: outer pass reads the (k, n-k) pairs straight from global memory
	for (int tid = t0; tid < t1; tid++) {
		if (DD) {
			for_outer_units(nseq, f.Mq(), tid, nthr, [&](int seq, int i) {
				GlobalRows<T, LoadOp, DD> src;
				const int g = DD == 1 ? seq : seq / (DD ? DD : 1), ch = DD == 1 ? 0 : seq - g * DD;
				src.ca = Coord{0, 0, 0, 0, 0}; src.cb = src.ca;
				src.pa = (T *)gin + (long long)(line0 + 2 * g) * a.ls_in + ch; src.pb = src.pa + a.ls_in;
				src.d = DD; src.ax_slot = a.ax_slot; src.op = &lop;
				dct3_outer_unit<T>(SmemBf<T, F>{s + seq * f.NPAD(), &f}, f, i, src);
			});
			continue;
		}
		for_outer_units(nseq, f.Mq(), tid, nthr, [&](int seq, int i) {
			const int g = seq / d, ch = seq - g * d;
			const int la = line0 + 2 * g;
			const bool hasb = (2 * g + 1) < nl;
			GlobalRows<T, LoadOp> src;
			long long ia, oa, ib = 0, ob = 0;
			src.ca = Coord{0, 0, 0, 0, 0}; src.cb = Coord{0, 0, 0, 0, 0};
			outer_decode(a.o, (uint32_t)la, ia, oa, src.ca);
			if (hasb) outer_decode(a.o, (uint32_t)la + 1, ib, ob, src.cb);
			src.ca.ch = ch; src.cb.ch = ch;
			src.pa = (T *)gin + ia + ch; src.pb = hasb ? (T *)gin + ib + ch : (T *)0;
			src.d = d; src.ax_slot = a.ax_slot; src.op = &lop;
			dct3_outer_unit<T>(SmemBf<T, F>{s + seq * f.NPAD(), &f}, f, i, src);
		});
	}
	DSP_SYNC();
	for (int q = f.NMID() - 1; q >= 0; q--) {
		for (int tid = t0; tid < t1; tid++) mid_pass<T, false>(s, nseq, f, q, tid, nthr);
		DSP_SYNC();
	}
	contig_pass<T>(s, nseq, f, t0, t1, nthr);
	for (int tid = t0; tid < t1; tid++) {
		if (DD == 1) row_move_planar4_pick<T, false, StoreOp>(a, f, sop, line0, npairs, tid, nthr, s);
		else if (DD > 1) row_move_inter4<T, (DD > 1 ? DD : 2), false, StoreOp, F, true>(a, f, sop, line0, nl, tid, nthr, s);
		else row_move_any<T, false, StoreOp>(a, f, sop, line0, nl, tid, nthr, s);
	}
}

// ------------------------------------------------------------------------------------------------ column pass (fast)
// Moves the CTA's column tile between global memory and smem in groups of W columns, UNR groups per thread in
// flight.  IN: global -> lop -> slot ; !IN: slot -> (re, +-im) -> sop -> global.  `scatter` selects
// sig[makhoul(r)] vs Pad(r).  `vec`: W-wide accesses are legal (alignment + the tile is a whole number of groups).
template <class T, int W, bool IN, class Op, class F>
DSP_DEV void col_move(const ColArgs &a, const F &f, const Op &op, bool scatter, bool negim, bool vec, int col0, int ncl,
                      long long gbase, const Coord &cbase, int tid, int nthr, C2<T> *s) {
	typedef VecW<T, W> Vec;
	const int UNR = 8;
	const T *gin = (const T *)a.in;
	T *gout = (T *)a.out;
	const int n = f.N();
	const int gpr = (ncl + W - 1) / W;                        // groups per axis position
	const int total = n * gpr;
	const long long axs = IN ? a.ax_is : a.ax_os;
	const int lg = (gpr & (gpr - 1)) == 0 ? ilog2(gpr) : -1;
	for (int i0 = tid; i0 < total; i0 += nthr * UNR) {
		T v[UNR][W];
		if (IN) {
#pragma unroll
			for (int u = 0; u < UNR; u++) {
				const int idx = i0 + u * nthr;
				if (idx < total) {
					const int r = lg >= 0 ? idx >> lg : idx / gpr, cg = idx - r * gpr;
					const int c0 = cg * W;
					const T *src = gin + gbase + (long long)r * axs + col0 + c0;
					if (vec) {
						const Vec tv = ldg_stream((const Vec *)src);
#pragma unroll
						for (int t = 0; t < W; t++) v[u][t] = tv.v[t];
					} else {
#pragma unroll
						for (int t = 0; t < W; t++) v[u][t] = (c0 + t < ncl) ? src[t] : (T)0;
					}
				}
			}
		}
#pragma unroll
		for (int u = 0; u < UNR; u++) {
			const int idx = i0 + u * nthr;
			if (idx < total) {
				const int r = lg >= 0 ? idx >> lg : idx / gpr, cg = idx - r * gpr;
				const int c0 = cg * W;
				const int slot = scatter ? (int)DSP_LDG(f.sig + makhoul(r, n)) : Pad<T>::of(r);
				Coord c = cbase;
				c.set(a.ax_slot, r);
				if (!IN) {
#pragma unroll
					for (int p = 0; p < W / 2; p++) {
						v[u][2 * p] = 0; v[u][2 * p + 1] = 0;
						if (c0 + 2 * p < ncl) {
							const C2<T> z = s[(c0 / 2 + p) * f.NPAD() + slot];
							v[u][2 * p] = z.x;
							v[u][2 * p + 1] = negim ? -z.y : z.y;
						}
					}
				}
#pragma unroll
				for (int t = 0; t < W; t++) {
					if (c0 + t < ncl) {
						const int col = col0 + c0 + t;
						int x = col, ch = 0;
						if (a.d != 1) { x = (int)fd_div((uint32_t)col, a.dd); ch = col - x * a.d; }
						c.ch = ch; c.set(a.col_slot, x);
						v[u][t] = op(v[u][t], c);
					}
				}
				if (IN) {
#pragma unroll
					for (int p = 0; p < W / 2; p++)
						if (c0 + 2 * p < ncl) s[(c0 / 2 + p) * f.NPAD() + slot] = C2<T>{v[u][2 * p], v[u][2 * p + 1]};
				} else {
					T *dst = gout + gbase + (long long)r * axs + col0 + c0;
					if (vec) {
						Vec res;
#pragma unroll
						for (int t = 0; t < W; t++) res.v[t] = v[u][t];
						*(Vec *)dst = res;
					} else {
#pragma unroll
						for (int t = 0; t < W; t++)
							if (c0 + t < ncl) dst[t] = v[u][t];
					}
				}
			}
		}
	}
}

struct SlotSigMakhoul {                                       // DCT-II input / DCT-III output of a whole axis
	const uint16_t *sig; int n;
	DSP_DEVM int operator()(int r) const { return (int)DSP_LDG(sig + makhoul(r, n)); }
};

// ---- the same move with the axis length fixed at compile time (F = FastFixed), 256 threads and a full TC-column
// tile: thread = (column group cg, row phase r0) walks rows r = r0 + DR u; the global row offsets, the natural-order
// slots Pad(r0) + nat_delta(u) and the offsets into the slot table are compile-time in u.
template <int TC> struct FixedTile {
	enum { GPR = TC / 4, DR = 256 / GPR };                               // column groups per row, rows per step
	// Pad(r0 + DR u) - Pad(r0) for r0 < DR <= 64, DR | 256, r0 + DR u < 4096
	DSP_HDM static constexpr int nat_delta(int u) { return u * (DR + DR / 16) + ((u * DR) >> 8); }
};
template <class T, int TC, bool IN, bool SCATTER, class Op, class F>
DSP_DEV void col_tile_fixed(const T *gin, T *gout, long long rs, const Op &op, bool negim, const F &f, int tid, C2<T> *s) {
	typedef VecW<T, 4> Vec;
	typedef FixedTile<TC> G;
	const int n = F::kN, U = n / G::DR, UNR = U < 8 ? U : 8;
	const int cg = tid & (G::GPR - 1), r0 = tid / G::GPR;
	C2<T> *sq = s + (2 * cg) * f.NPAD();
	const T *gp = gin + 4 * cg + (long long)r0 * rs;
	T *gq = gout + 4 * cg + (long long)r0 * rs;
	const long long step = (long long)G::DR * rs;
	// SCATTER: slot = sig[makhoul(r)]; r keeps the parity of r0, so the table index moves by +-DR/2 per step
	const uint16_t *sg = f.sig + makhoul(r0, n);
	const int sdir = (r0 & 1) ? -(G::DR / 2) : (G::DR / 2);
	const C2<T> *snat = sq + Pad<T>::of(r0);
	const Coord cz = {0, 0, 0, 0, 0};
#pragma unroll
	for (int ub = 0; ub < U; ub += UNR) {
		Vec v[UNR];
		if (IN) {
#pragma unroll
			for (int u = 0; u < UNR; u++) v[u] = ldg_stream((const Vec *)(gp + (ub + u) * step));
		}
#pragma unroll
		for (int u = 0; u < UNR; u++) {
			const int uu = ub + u;
			C2<T> *p0 = SCATTER ? sq + (int)DSP_LDG(sg + sdir * uu) : (C2<T> *)snat + G::nat_delta(uu);
			if (IN) {
				p0[0] = C2<T>{op(v[u].v[0], cz), op(v[u].v[1], cz)};
				p0[f.NPAD()] = C2<T>{op(v[u].v[2], cz), op(v[u].v[3], cz)};
			} else {
				const C2<T> z0 = p0[0], z1 = p0[f.NPAD()];
				Vec o;
				o.v[0] = op(z0.x, cz); o.v[1] = op(negim ? -z0.y : z0.y, cz);
				o.v[2] = op(z1.x, cz); o.v[3] = op(negim ? -z1.y : z1.y, cz);
				*(Vec *)(gq + uu * step) = o;
			}
		}
	}
}

// picks the access width: full 16-byte groups when the tile allows it, else 8-byte (float) pairs, else scalar
template <class T, bool IN, class Op, class F>
DSP_DEV void col_move_any(const ColArgs &a, const F &f, const Op &op, bool scatter, bool negim, int col0, int ncl,
                          long long gbase, const Coord &cbase, int tid, int nthr, C2<T> *s) {
	const int VN = VecOf<T>::N;
	const bool al = IN ? a.vec_in : a.vec_out;
	if (sizeof(T) == 4 && !Op::kNeedsCoord && lean_ok(ncl, a.tc, nthr, al)) {
		const T *gin = (const T *)a.in + gbase + col0;
		T *gout = (T *)a.out + gbase + col0;
		const long long rs = IN ? a.ax_is : a.ax_os;
		if (!IN && a.seg_rows > 0) {                          // segmented output rows (peer buffers): general lean move
			const int lgs = ilog2(a.tc / 4);
			if (scatter) tile_move_lean<T, IN, Op>(gin, gout, rs, f.N(), lgs, op, negim, RowSeg{a.seg_rows, a.dseg, a.seg_off}, SlotSigMakhoul{f.sig, f.N()}, f.NPAD(), tid, nthr, s);
			else tile_move_lean<T, IN, Op>(gin, gout, rs, f.N(), lgs, op, negim, RowSeg{a.seg_rows, a.dseg, a.seg_off}, SlotNat<T>(), f.NPAD(), tid, nthr, s);
			return;
		}
		if constexpr (F::kFixed != 0 && sizeof(T) == 4) {
			if (nthr == 256 && F::kN <= 4096) {
#define DSP_COL_FIXED(TC)                                                                                              \
	if (a.tc == TC && (F::kN % (256 / (TC / 4))) == 0) {                                                               \
		if (scatter) col_tile_fixed<T, TC, IN, true, Op>(gin, gout, rs, op, negim, f, tid, s);                         \
		else col_tile_fixed<T, TC, IN, false, Op>(gin, gout, rs, op, negim, f, tid, s);                                \
		return;                                                                                                        \
	}
				DSP_COL_FIXED(32) DSP_COL_FIXED(16)
#undef DSP_COL_FIXED
			}
		}
		const int lg = ilog2(a.tc / 4);
		if (scatter) tile_move_lean<T, IN, Op>(gin, gout, rs, f.N(), lg, op, negim, RowIdent(), SlotSigMakhoul{f.sig, f.N()}, f.NPAD(), tid, nthr, s);
		else tile_move_lean<T, IN, Op>(gin, gout, rs, f.N(), lg, op, negim, RowIdent(), SlotNat<T>(), f.NPAD(), tid, nthr, s);
		return;
	}
	if (al && (ncl % VN) == 0) col_move<T, VecOf<T>::N, IN, Op>(a, f, op, scatter, negim, true, col0, ncl, gbase, cbase, tid, nthr, s);
	else if (sizeof(T) == 4 && al && (ncl % 2) == 0 && (a.tc % 2) == 0) col_move<T, 2, IN, Op>(a, f, op, scatter, negim, true, col0, ncl, gbase, cbase, tid, nthr, s);
	else col_move<T, 2, IN, Op>(a, f, op, scatter, negim, false, col0, ncl, gbase, cbase, tid, nthr, s);
}

template <class T, bool FWD, class LoadOp, class StoreOp, class F>
DSP_DEV void cta_col_fast(const ColArgs &a, const F &f, const LoadOp &lop, const StoreOp &sop, int cta, int t0, int t1,
                          int nthr, C2<T> *s) {
	const uint32_t oidx = fd_div((uint32_t)cta, a.dtiles);
	const int tile = cta - (int)oidx * a.ntiles;
	const int col0 = tile * a.tc;
	int ncl = a.ncols - col0;
	if (ncl > a.tc) ncl = a.tc;
	const int nseq = (ncl + 1) / 2;
	Coord cbase = {0, 0, 0, 0, 0};
	long long ibase, obase;
	outer_decode(a.o, oidx, ibase, obase, cbase);

	// just-in-time L2 prefetch of the tile that will follow on this SM: one 128-byte line per axis position covers
	// 32 columns, so only every (32 / tc)-th tile issues it
	if (a.pf_dist > 0) {
		const int pcta = cta + a.pf_dist;
		const int tpl = 32 / a.tc > 0 ? 32 / a.tc : 1;
		const uint32_t po = fd_div((uint32_t)pcta, a.dtiles);
		const int ptile = pcta - (int)po * a.ntiles;
		if (po == oidx && ptile < a.ntiles && (ptile % tpl) == 0) {
			const char *pb = (const char *)a.in + (ibase + (long long)ptile * a.tc) * (long long)sizeof(T);
			for (int tid = t0; tid < t1; tid++)
				for (int r = tid; r < f.N(); r += nthr) prefetch_l2(pb + (long long)r * a.ax_is * (long long)sizeof(T));
		}
	}

	// ---- copy-in: DCT-II scatters through sig, DCT-III keeps natural order
	for (int tid = t0; tid < t1; tid++) col_move_any<T, true, LoadOp>(a, f, lop, FWD, false, col0, ncl, ibase, cbase, tid, nthr, s);
	DSP_SYNC();

	if (FWD) {
		contig_pass<T>(s, nseq, f, t0, t1, nthr);
		for (int q = 0; q < f.NMID(); q++) {
			for (int tid = t0; tid < t1; tid++) mid_pass<T, true>(s, nseq, f, q, tid, nthr);
			DSP_SYNC();
		}
		for (int tid = t0; tid < t1; tid++) {
			for_outer_units(nseq, f.Mq(), tid, nthr, [&](int seq, int i) {
				SmemNat<T> sink;
				sink.base = s + seq * f.NPAD();
				dct2_outer_unit<T>(SmemBf<T, F>{sink.base, &f}, f, i, sink);
			});
		}
		DSP_SYNC();
	} else {
		for (int tid = t0; tid < t1; tid++) {
			for_outer_units(nseq, f.Mq(), tid, nthr, [&](int seq, int i) {
				SmemNat<T> src;
				src.base = s + seq * f.NPAD();
				dct3_outer_unit<T>(SmemBf<T, F>{src.base, &f}, f, i, src);
			});
		}
		DSP_SYNC();
		for (int q = f.NMID() - 1; q >= 0; q--) {
			for (int tid = t0; tid < t1; tid++) mid_pass<T, false>(s, nseq, f, q, tid, nthr);
			DSP_SYNC();
		}
		contig_pass<T>(s, nseq, f, t0, t1, nthr);
	}

	// ---- copy-out: DCT-II results sit in natural order, DCT-III results at their digit-reversed slots
	for (int tid = t0; tid < t1; tid++) col_move_any<T, false, StoreOp>(a, f, sop, !FWD, !FWD, col0, ncl, obase, cbase, tid, nthr, s);
}

}  // namespace dsp
