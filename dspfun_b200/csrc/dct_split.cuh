// dct_split.cuh -- long strided (column) axes as two L2-resident sub-passes, n = 16 M (power of two, float).
//
// A column tile of the full length does not leave room for a second tile in shared memory (n = 8192: 4 columns x
// 8192 x 8.5 B = 140 KB), so the single-kernel column pass cannot overlap its load / compute / store phases and only
// ever touches 16 B per row.  Here the length-n transform is split along its outermost radix-16 stage:
//
//   DCT-II :  A) 16 independent M-point FFTs per column pair (rows e = 16 e' + j of the permuted sequence), tiles of
//                32 columns x M rows (70 KB at M = 512; several CTAs per SM; 128 B row segments), results written to
//                a panel-sized scratch [16 M][P] that stays in L2;
//             B) the outer radix-16 butterflies + (k, n-k) post-twiddle straight from / to global memory, one thread
//                per (column pair, butterfly pair i | M-i), lanes along columns (256 B per warp instruction), no smem.
//   DCT-III:  B') pre-twiddle + outer radix-16 into the scratch, then A') the M-point FFTs and the un-permuting store.
//
// The plane is processed in panels of P columns (n * P * 4 B ~ 32 MB), A then B per panel, so the scratch traffic
// never reaches HBM: per sample the pass still reads 4 B and writes 4 B of DRAM.
#pragma once
#include "dct_fast.cuh"

namespace dsp {

struct SplitArgs {
	int n, M;                    // n = 16 M
	int kind;
	int d;                       // interleave of the column index (col = x*d + ch), for coordinates only
	FastDiv dd;
	long long ax_is, ax_os;      // row strides of in / out (elements)
	long long ax_ss;             // row stride of the scratch (elements, even)
	int ax_slot, col_slot;
	const void *in;              // already offset to the outer index (batch / frame) of this launch
	void *out;
	void *scratch;
	Coord cbase;                 // coordinates of the outer index
	int pcol0, pcols;            // panel column range [pcol0, pcol0 + pcols)
	int tc, ntiles;              // sub-pass A: columns per CTA, tiles per panel
	int ngroups;                 // sub-pass B: 32-pair column groups per panel
	int pf_warps;                // inverse sub-pass B': L2 prefetch distance in warps (0 = off)
	int tci, ntilesi;            // DIT-style inverse, tile sub-pass: columns per CTA (16), tiles per panel
};

DSP_DEV int split_row(int e, int n) { return e < n / 2 ? 2 * e : 2 * (n - 1 - e) + 1; }   // inverse Makhoul: row holding v[e]

// ------------------------------------------------------------------------------------------------ sub-pass A
// idx -> (global row, smem slot) for the four moves of sub-pass A
template <class T, int W, bool IN, bool GLOBAL_SIDE_IS_IMAGE, class Op, class F>
DSP_DEV void split_move(const SplitArgs &a, const F &fM, const Op &op, int j, int col0, int ncl, bool negim, int tid,
                        int nthr, C2<T> *s) {
	typedef VecW<T, W> Vec;
	const int UNR = 8;
	const int M = a.M, n = a.n;
	const int gpr = (ncl + W - 1) / W;
	const int total = M * gpr;
	const bool vec = (ncl % W) == 0;
	const int lg = (gpr & (gpr - 1)) == 0 ? ilog2(gpr) : -1;
	const T *gin = (const T *)(GLOBAL_SIDE_IS_IMAGE ? a.in : a.scratch);
	T *gout = (T *)(GLOBAL_SIDE_IS_IMAGE ? a.out : a.scratch);
	const long long rs = GLOBAL_SIDE_IS_IMAGE ? (IN ? a.ax_is : a.ax_os) : a.ax_ss;
	const int cofs = GLOBAL_SIDE_IS_IMAGE ? col0 : col0 - a.pcol0;
	for (int i0 = tid; i0 < total; i0 += nthr * UNR) {
		T v[UNR][W];
		if (IN) {
#pragma unroll
			for (int u = 0; u < UNR; u++) {
				const int idx = i0 + u * nthr;
				if (idx < total) {
					const int r = lg >= 0 ? idx >> lg : idx / gpr, cg = idx - r * gpr;
					const int c0 = cg * W;
					const long long grow = GLOBAL_SIDE_IS_IMAGE ? split_row(16 * r + j, n) : (long long)j * M + r;
					const T *src = gin + grow * rs + cofs + c0;
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
				const int grow = GLOBAL_SIDE_IS_IMAGE ? split_row(16 * r + j, n) : j * M + r;
				// image side: element e' of sub-FFT j lives at the digit-reversed slot; scratch side: natural order
				const int slot = GLOBAL_SIDE_IS_IMAGE ? (int)DSP_LDG(fM.sig + r) : Pad<T>::of(r);
				Coord c = a.cbase;
				c.set(a.ax_slot, grow);
				if (!IN) {
#pragma unroll
					for (int p = 0; p < W / 2; p++) {
						v[u][2 * p] = 0; v[u][2 * p + 1] = 0;
						if (c0 + 2 * p < ncl) {
							const C2<T> z = s[(c0 / 2 + p) * fM.NPAD() + slot];
							v[u][2 * p] = z.x;
							v[u][2 * p + 1] = negim ? -z.y : z.y;
						}
					}
				}
				if (GLOBAL_SIDE_IS_IMAGE) {
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
				}
				if (IN) {
#pragma unroll
					for (int p = 0; p < W / 2; p++)
						if (c0 + 2 * p < ncl) s[(c0 / 2 + p) * fM.NPAD() + slot] = C2<T>{v[u][2 * p], v[u][2 * p + 1]};
				} else {
					T *dst = gout + (long long)grow * rs + cofs + c0;
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

// ---- tile moves of sub-pass A with the sub-FFT length M fixed at compile time (F = FastFixed), 256 threads and
// full TC-column tiles.  Thread = (column group cg, row phase r0); it walks rows r = r0 + DR u.  Everything that
// depends on u is a compile-time constant: the general tile_move_lean spends ~20 instructions per row on the row
// map, a 64-bit multiply and the predicate (ncu source view, profiles/r01_ncu_current_summary.md).
// image rows of sub-FFT j (Makhoul order: element e = 16 r + j sits in row 2e for e < n/2, else 2(n-1-e)+1)
// -> digit-reversed slots.  First half of the rows ascends in steps of 32 DR image rows, second half descends.
template <class T, int TC, class Op, class F>
DSP_DEV void split_image_load_fixed(const T *gtile, long long rs, int j, const Op &op, const F &fM, int tid, C2<T> *s) {
	typedef VecW<T, 4> Vec;
	typedef FixedTile<TC> G;
	const int M = F::kN, U = M / G::DR, UNR = U < 8 ? U : 8;
	const int n = 16 * M;
	const int cg = tid & (G::GPR - 1), r0 = tid / G::GPR;
	C2<T> *sq = s + (2 * cg) * fM.NPAD();
	const uint16_t *sg = fM.sig + r0;
	const T *p1 = gtile + 4 * cg + (long long)(2 * (16 * r0 + j)) * rs;                                  // u = 0
	const T *p2 = gtile + 4 * cg + (long long)(2 * (n - 1 - (16 * (r0 + G::DR * (U / 2)) + j)) + 1) * rs; // u = U/2
	const long long step = (long long)(32 * G::DR) * rs;
	const Coord cz = {0, 0, 0, 0, 0};
#pragma unroll
	for (int ub = 0; ub < U; ub += UNR) {
		Vec v[UNR];
#pragma unroll
		for (int u = 0; u < UNR; u++) {
			const int uu = ub + u;
			v[u] = ldg_stream((const Vec *)(uu < U / 2 ? p1 + uu * step : p2 - (uu - U / 2) * step));
		}
#pragma unroll
		for (int u = 0; u < UNR; u++) {
			const int slot = (int)DSP_LDG(sg + G::DR * (ub + u));
			sq[slot] = C2<T>{op(v[u].v[0], cz), op(v[u].v[1], cz)};
			sq[fM.NPAD() + slot] = C2<T>{op(v[u].v[2], cz), op(v[u].v[3], cz)};
		}
	}
}

// natural-order smem rows -> scratch rows rowbase + r (rowbase = j M)
template <class T, int TC, class F>
DSP_DEV void split_scratch_store_fixed(T *gtile, long long ss, int rowbase, const F &fM, int tid, const C2<T> *s) {
	typedef VecW<T, 4> Vec;
	typedef FixedTile<TC> G;
	const int M = F::kN, U = M / G::DR;
	const int cg = tid & (G::GPR - 1), r0 = tid / G::GPR;
	const C2<T> *sq = s + (2 * cg) * fM.NPAD() + Pad<T>::of(r0);
	T *p = gtile + 4 * cg + (long long)(rowbase + r0) * ss;
	const long long step = (long long)G::DR * ss;
#pragma unroll
	for (int u = 0; u < U; u++) {
		const C2<T> z0 = sq[G::nat_delta(u)], z1 = sq[fM.NPAD() + G::nat_delta(u)];
		Vec o;
		o.v[0] = z0.x; o.v[1] = z0.y; o.v[2] = z1.x; o.v[3] = z1.y;
		*(Vec *)(p + u * step) = o;
	}
}

// sub-pass A'' load of a paired CTA (sub-sequences ja = jj and jb = 16 - jj, 0 < jj < 8) with M fixed: element kp of
// ja pairs with element kb = M-1-kp of jb (rows k = 16 kp + ja and n - k); pre-twiddle in registers, both results to
// the digit-reversed slots.  kp = k0 + DR u: rows, table entries and the k <= n/2 case are compile-time in u.
template <class T, int TC, class Op, class F>
DSP_DEV void split_inv_load_fixed(const T *gtile, long long rs, int ja, int jb, const Op &lop, const F &fM, const C2<T> *om,
                                  int tid, C2<T> *sA, C2<T> *sB) {
	typedef VecW<T, 4> Vec;
	typedef FixedTile<TC> G;
	const int M = F::kN, U = M / G::DR, UNR = U < 4 ? U : 4;
	const int n = 16 * M;
	const int cg = tid & (G::GPR - 1), k0 = tid / G::GPR;
	C2<T> *qa = sA + (2 * cg) * fM.NPAD(), *qb = sB + (2 * cg) * fM.NPAD();
	const T *pa = gtile + 4 * cg + (long long)(16 * k0 + ja) * rs;
	const T *pb = gtile + 4 * cg + (long long)(16 * (M - 1 - k0) + jb) * rs;
	const long long step = (long long)(16 * G::DR) * rs;
	const uint16_t *sgA = fM.sig + k0, *sgB = fM.sig + (M - 1 - k0);
	const C2<T> *omA = om + (16 * k0 + ja), *omB = om + (n - 16 * k0 - ja);
	const Coord cz = {0, 0, 0, 0, 0};
#pragma unroll
	for (int ub = 0; ub < U; ub += UNR) {
		Vec va[UNR], vb[UNR];
#pragma unroll
		for (int u = 0; u < UNR; u++) {
			va[u] = ldg_stream((const Vec *)(pa + (ub + u) * step));
			vb[u] = ldg_stream((const Vec *)(pb - (ub + u) * step));
		}
#pragma unroll
		for (int u = 0; u < UNR; u++) {
			const int uu = ub + u;
			const int sk = (int)DSP_LDG(sgA + G::DR * uu), sn = (int)DSP_LDG(sgB - G::DR * uu);
			const bool low = uu < U / 2;                              // k <= n - k
			const C2<T> w = low ? ldg_c2(omA + 16 * G::DR * uu) : ldg_c2(omB - 16 * G::DR * uu);
#pragma unroll
			for (int p = 0; p < 2; p++) {
				const C2<T> xa = C2<T>{lop(va[u].v[2 * p], cz), lop(va[u].v[2 * p + 1], cz)};
				const C2<T> xb = C2<T>{lop(vb[u].v[2 * p], cz), lop(vb[u].v[2 * p + 1], cz)};
				C2<T> wk, wn;
				if (low) { dct3_pair<T>(w, xa, xb, wk, wn); qa[p * fM.NPAD() + sk] = wk; qb[p * fM.NPAD() + sn] = wn; }
				else { dct3_pair<T>(w, xb, xa, wk, wn); qb[p * fM.NPAD() + sn] = wk; qa[p * fM.NPAD() + sk] = wn; }
			}
		}
	}
}

struct RowSplitImage { int j, n; DSP_DEVM long long off(int r, long long rs) const { return (long long)split_row(16 * r + j, n) * rs; } };
struct RowSplitScratch { int base; DSP_DEVM long long off(int r, long long rs) const { return (long long)(base + r) * rs; } };
struct SlotSig { const uint16_t *sig; DSP_DEVM int operator()(int r) const { return (int)DSP_LDG(sig + r); } };

// image-side and scratch-side moves of sub-pass A with the lean path when the tile is full and aligned
template <class T, bool IN, bool IMAGE, class Op, class F>
DSP_DEV void split_move_any(const SplitArgs &a, const F &fM, const Op &op, int j, int col0, int ncl, bool negim, int tid,
                            int nthr, C2<T> *s) {
	const int W = VecOf<T>::N;
	if (sizeof(T) == 4 && !Op::kNeedsCoord && lean_ok(ncl, a.tc, nthr, true)) {
		const int lg = ilog2(a.tc / 4);
		if (IMAGE) {
			const T *gin = (const T *)a.in + col0;
			T *gout = (T *)a.out + col0;
			tile_move_lean<T, IN, Op>(gin, gout, IN ? a.ax_is : a.ax_os, fM.N(), lg, op, negim, RowSplitImage{j, a.n}, SlotSig{fM.sig}, fM.NPAD(), tid, nthr, s);
		} else {
			const T *g = (const T *)a.scratch + (col0 - a.pcol0);
			tile_move_lean<T, IN, Op>(g, (T *)g, a.ax_ss, fM.N(), lg, op, negim, RowSplitScratch{j * fM.N()}, SlotNat<T>(), fM.NPAD(), tid, nthr, s);
		}
		return;
	}
	split_move<T, W, IN, IMAGE, Op>(a, fM, op, j, col0, ncl, negim, tid, nthr, s);
}

// CTA = (column tile, sub-FFT j).  FWD: image rows -> M-point DIT FFT -> scratch block j.
// !FWD: scratch block j -> M-point DIF FFT -> image rows (un-permuted).
template <class T, bool FWD, class LoadOp, class StoreOp, class F>
DSP_DEV void cta_split_fft(const SplitArgs &a, const F &fM, const LoadOp &lop, const StoreOp &sop, int cta, int t0, int t1,
                           int nthr, C2<T> *s) {
	const int j = cta / a.ntiles, tile = cta - j * a.ntiles;
	const int col0 = a.pcol0 + tile * a.tc;
	int ncl = a.pcol0 + a.pcols - col0;
	if (ncl > a.tc) ncl = a.tc;
	const int nseq = (ncl + 1) / 2;
	// fixed-length lean moves: compile-time M, 256 threads, full aligned 32-column tile, coordinate-free op
	const bool fixed = F::kFixed && sizeof(T) == 4 && nthr == 256 && a.tc == 32 && ncl == 32 && !LoadOp::kNeedsCoord &&
	                   (F::kN % 64) == 0;
	if (FWD) {
		bool done = false;
		if constexpr (F::kFixed != 0) {
			if (fixed) {
				for (int tid = t0; tid < t1; tid++) split_image_load_fixed<T, 32, LoadOp>((const T *)a.in + col0, a.ax_is, j, lop, fM, tid, s);
				done = true;
			}
		}
		if (!done) for (int tid = t0; tid < t1; tid++) split_move_any<T, true, true, LoadOp>(a, fM, lop, j, col0, ncl, false, tid, nthr, s);
		DSP_SYNC();
		contig_pass<T>(s, nseq, fM, t0, t1, nthr);
		for (int q = 0; q <= fM.NMID(); q++) {
			for (int tid = t0; tid < t1; tid++) mid_pass<T, true>(s, nseq, fM, q, tid, nthr);
			DSP_SYNC();
		}
		done = false;
		if constexpr (F::kFixed != 0) {
			if (fixed) {
				for (int tid = t0; tid < t1; tid++) split_scratch_store_fixed<T, 32>((T *)a.scratch + (col0 - a.pcol0), a.ax_ss, j * F::kN, fM, tid, s);
				done = true;
			}
		}
		if (!done) for (int tid = t0; tid < t1; tid++) split_move_any<T, false, false, OpNone>(a, fM, OpNone(), j, col0, ncl, false, tid, nthr, s);
	} else {
		for (int tid = t0; tid < t1; tid++) split_move_any<T, true, false, OpNone>(a, fM, OpNone(), j, col0, ncl, false, tid, nthr, s);
		DSP_SYNC();
		for (int q = fM.NMID(); q >= 0; q--) {
			for (int tid = t0; tid < t1; tid++) mid_pass<T, false>(s, nseq, fM, q, tid, nthr);
			DSP_SYNC();
		}
		contig_pass<T>(s, nseq, fM, t0, t1, nthr);
		for (int tid = t0; tid < t1; tid++) split_move_any<T, false, true, StoreOp>(a, fM, sop, j, col0, ncl, true, tid, nthr, s);
	}
}

// ------------------------------------------------------------------------------------------------ sub-pass B
// one column pair of the image: rows k of two adjacent columns, through the fused op
// LEAN (chosen at launch): every pair has both columns, 8-byte access is legal, the op ignores coordinates and the
// image spans < 2^31 elements, so an access is one 32-bit multiply-add on the pointer instead of a branch, a
// select and a 64-bit address computation.
template <class T, class Op, bool LEAN = false> struct GlobalCols {
	T *p;                        // image + column
	long long rs;                // row stride (elements)
	bool hasb, vec;              // second column exists; 8-byte access legal
	int ax_slot;
	Coord ca, cb;
	const Op *op;
	DSP_DEVM void put(int k, T xa, T xb) {
		if (LEAN) { *(C2<T> *)(p + (uint32_t)k * (uint32_t)rs) = C2<T>{(*op)(xa, ca), (*op)(xb, cb)}; return; }
		ca.set(ax_slot, k); cb.set(ax_slot, k);
		const T ya = (*op)(xa, ca), yb = hasb ? (*op)(xb, cb) : (T)0;
		T *q = p + (long long)k * rs;
		if (vec) *(C2<T> *)q = C2<T>{ya, yb};
		else { q[0] = ya; if (hasb) q[1] = yb; }
	}
	DSP_DEVM C2<T> get(int k) {
		if (LEAN) { const C2<T> v = *(const C2<T> *)(p + (uint32_t)k * (uint32_t)rs); return C2<T>{(*op)(v.x, ca), (*op)(v.y, cb)}; }
		ca.set(ax_slot, k); cb.set(ax_slot, k);
		const T *q = p + (long long)k * rs;
		C2<T> v;
		if (vec) v = *(const C2<T> *)q;
		else { v.x = q[0]; v.y = hasb ? q[1] : (T)0; }
		return C2<T>{(*op)(v.x, ca), hasb ? (*op)(v.y, cb) : (T)0};
	}
};

// thread = (column pair, unit i): lanes run along the columns, one warp per unit
template <class T, bool FWD, class LoadOp, class StoreOp, bool LEAN = false>
DSP_DEV void split_outer_thread(const SplitArgs &a, const FastDesc &fN, const LoadOp &lop, const StoreOp &sop, int gwarp, int lane) {
	const int group = gwarp % a.ngroups, i = gwarp / a.ngroups;
	if (!FWD && a.pf_warps > 0) {
		// the inverse reads 32 strided image rows per thread straight from DRAM: prefetch, one resident wave of warps
		// ahead, the 2 x 16 row segments (256 B each) the warp `gwarp + pf_warps` is going to load
		const int pw = gwarp + a.pf_warps;
		const int pg = pw % a.ngroups, pi = pw / a.ngroups;
		if (pi <= a.M / 2) {
			const int prow = (lane < 16 ? pi : (pi == 0 ? 0 : a.M - pi)) + a.M * (lane & 15);
			const char *pp = (const char *)((const T *)a.in + (long long)prow * a.ax_is + a.pcol0 + 64 * pg);
			prefetch_l2(pp);
			prefetch_l2(pp + 128);
		}
	}
	if (i > a.M / 2) return;
	const int pair = group * 32 + lane;
	const int col = a.pcol0 + 2 * pair;
	if (col >= a.pcol0 + a.pcols) return;
	const bool hasb = col + 1 < a.pcol0 + a.pcols;
	GlobBf<T> bf;
	bf.base = (C2<T> *)a.scratch + pair;
	bf.rs = a.ax_ss / 2;
	bf.M = a.M;
	Coord ca = a.cbase, cb = a.cbase;
	{
		int x = col, ch = 0;
		if (a.d != 1) { x = (int)fd_div((uint32_t)col, a.dd); ch = col - x * a.d; }
		ca.ch = ch; ca.set(a.col_slot, x);
		x = col + 1; ch = 0;
		if (a.d != 1) { x = (int)fd_div((uint32_t)(col + 1), a.dd); ch = col + 1 - x * a.d; }
		cb.ch = ch; cb.set(a.col_slot, x);
	}
	if (FWD) {
		GlobalCols<T, StoreOp, LEAN> sink;
		sink.p = (T *)a.out + col; sink.rs = a.ax_os; sink.hasb = hasb; sink.vec = hasb && (a.ax_os % 2) == 0 && (((size_t)a.out / sizeof(T) + col) % 2) == 0;
		sink.ax_slot = a.ax_slot; sink.ca = ca; sink.cb = cb; sink.op = &sop;
		dct2_outer_unit<T>(bf, fN, i, sink);
	} else {
		GlobalCols<T, LoadOp, LEAN> src;
		src.p = (T *)a.in + col; src.rs = a.ax_is; src.hasb = hasb; src.vec = hasb && (a.ax_is % 2) == 0 && (((size_t)a.in / sizeof(T) + col) % 2) == 0;
		src.ax_slot = a.ax_slot; src.ca = ca; src.cb = cb; src.op = &lop;
		dct3_outer_unit<T>(bf, fN, i, src);
	}
}

// ================================================================================================ DIT-style inverse
// DCT-III with the *tile* kernel first and the register kernel last (the efficient order, like the forward split):
//   A'') per column pair, sub-sequences w_j[k'] = W[16 k' + j] of the pre-twiddled spectrum W (dct3_pair) and their
//        M-point DIT FFTs G_j.  The pre-twiddle pairs k with n-k, i.e. element k' of sub-sequence j with element
//        M-1-k' of sub-sequence 16-j (j = 0: k' with M-k'; j = 8: k' with M-1-k'), so one CTA owns the sub-sequence
//        pair (j, 16-j) of a 16-column tile: 2 x 8 sequences x M.  G_j goes to scratch block j.
//   B'') F[i + M m] = sum_j W_n^{ij} W_16^{jm} G_j[i]: one plain radix-16 DIT butterfly per thread, lanes along the
//        columns, output sample e = i + M m stored (re, -im) to image row split_row(e).

// pre-twiddle of one (k, n-k) pair held in two smem slots (pk holds X[k], pn holds X[n-k]; any order of k vs n/2)
template <class T>
DSP_DEV void split_pretw(const C2<T> *om, int k, int n, C2<T> *pk, C2<T> *pn) {
	if (2 * k <= n) {
		C2<T> wk, wn;
		dct3_pair<T>(ldg_c2(om + k), *pk, *pn, wk, wn);
		*pk = wk; *pn = wn;
	} else {
		C2<T> wk, wn;
		dct3_pair<T>(ldg_c2(om + (n - k)), *pn, *pk, wk, wn);
		*pn = wk; *pk = wn;
	}
}

template <class T, class LoadOp, class F>
DSP_DEV void cta_split_inv_fft(const SplitArgs &a, const F &fM, const FastDesc &fN, const LoadOp &lop, int cta, int t0,
                               int t1, int nthr, C2<T> *s) {
	const int M = fM.N(), n = 16 * fM.N();
	const int jj = cta / a.ntilesi, tile = cta - jj * a.ntilesi;          // jj = 0..8
	const int col0 = a.pcol0 + tile * a.tci;
	const int nq = a.tci / 2;                                             // complex sequences per sub-sequence
	const bool paired = jj != 0 && jj != 8;
	const int ja = jj, jb = 16 - jj;
	const int nseq = paired ? 2 * nq : nq;
	const int lg = ilog2(a.tci / 4);
	const C2<T> *om = (const C2<T> *)fN.om;
	C2<T> *sA = s, *sB = s + nq * fM.NPAD();
	const T *gin = (const T *)a.in + col0;
	// ---- load both rows of every (k, n-k) pair, pre-twiddle in registers, store to the digit-reversed slots
	//      (element k' of sub-sequence ja pairs with element kb of sub-sequence jb; for the single sub-sequences 0 and 8
	//      both live in the same sequence set)
	typedef VecW<T, 4> Vec;
	const int UNR = 4;
	const int npair = paired ? M : (jj == 0 ? M / 2 + 1 : M / 2);
	C2<T> *sBB = paired ? sB : sA;
	bool fixed = false;
	if constexpr (F::kFixed != 0) fixed = sizeof(T) == 4 && nthr == 256 && a.tci == 16 && (F::kN % 128) == 0 && !LoadOp::kNeedsCoord;
	if constexpr (F::kFixed != 0) {
		if (fixed && paired)
			for (int tid = t0; tid < t1; tid++) split_inv_load_fixed<T, 16, LoadOp>(gin, a.ax_is, ja, jb, lop, fM, om, tid, sA, sB);
	}
	for (int tid = t0; tid < t1; tid++) {
		if (fixed && paired) break;
		const int cg = tid & ((1 << lg) - 1), dk = nthr >> lg;
		const T *gp = gin + 4 * cg;
		C2<T> *qa = sA + (2 * cg) * fM.NPAD(), *qb = sBB + (2 * cg) * fM.NPAD();
		for (int k0 = tid >> lg; k0 < npair; k0 += dk * UNR) {
			Vec va[UNR], vb[UNR];
#pragma unroll
			for (int u = 0; u < UNR; u++) {
				const int kp = k0 + u * dk;
				if (kp < npair) {
					const int kb = paired ? M - 1 - kp : (jj == 0 ? (kp == 0 ? 0 : M - kp) : M - 1 - kp);
					va[u] = ldg_stream((const Vec *)(gp + (long long)(16 * kp + ja) * a.ax_is));
					vb[u] = ldg_stream((const Vec *)(gp + (long long)(16 * kb + (paired ? jb : ja)) * a.ax_is));
				}
			}
#pragma unroll
			for (int u = 0; u < UNR; u++) {
				const int kp = k0 + u * dk;
				if (kp < npair) {
					const int kb = paired ? M - 1 - kp : (jj == 0 ? (kp == 0 ? 0 : M - kp) : M - 1 - kp);
					const int k = 16 * kp + ja;                       // k <= n - k may or may not hold for paired CTAs
					const int sk = (int)DSP_LDG(fM.sig + kp), sn = (int)DSP_LDG(fM.sig + kb);
					const Coord cz = {0, 0, 0, 0, 0};
					const bool self = !paired && kb == kp;            // k = 0 (jj = 0, kp = 0) or k = n/2 (jj = 0, kp = M/2)
					const C2<T> w = ldg_c2(om + (2 * k <= n ? k : n - k));
#pragma unroll
					for (int p = 0; p < 2; p++) {
						const C2<T> xa = C2<T>{lop(va[u].v[2 * p], cz), lop(va[u].v[2 * p + 1], cz)};
						const C2<T> xb = C2<T>{lop(vb[u].v[2 * p], cz), lop(vb[u].v[2 * p + 1], cz)};
						C2<T> wk, wn;
						if (self && k == 0) { qa[p * fM.NPAD() + sk] = C2<T>{xa.x, -xa.y}; continue; }
						if (2 * k <= n) { dct3_pair<T>(w, xa, xb, wk, wn); qa[p * fM.NPAD() + sk] = wk; if (!self) qb[p * fM.NPAD() + sn] = wn; }
						else { dct3_pair<T>(w, xb, xa, wk, wn); qb[p * fM.NPAD() + sn] = wk; qa[p * fM.NPAD() + sk] = wn; }
					}
				}
			}
		}
	}
	DSP_SYNC();
	// ---- M-point DIT FFTs
	contig_pass<T>(s, nseq, fM, t0, t1, nthr);
	for (int q = 0; q <= fM.NMID(); q++) {
		for (int tid = t0; tid < t1; tid++) mid_pass<T, true>(s, nseq, fM, q, tid, nthr);
		DSP_SYNC();
	}
	// ---- natural-order results to scratch blocks j (and 16 - j)
	T *gs = (T *)a.scratch + (col0 - a.pcol0);
	if constexpr (F::kFixed != 0) {
		if (fixed) {
			for (int tid = t0; tid < t1; tid++) {
				split_scratch_store_fixed<T, 16>(gs, a.ax_ss, ja * M, fM, tid, sA);
				if (paired) split_scratch_store_fixed<T, 16>(gs, a.ax_ss, jb * M, fM, tid, sB);
			}
			return;
		}
	}
	for (int tid = t0; tid < t1; tid++) {
		tile_move_lean<T, false, OpNone>(gs, gs, a.ax_ss, M, lg, OpNone(), false, RowSplitScratch{ja * M}, SlotNat<T>(), fM.NPAD(), tid, nthr, sA);
		if (paired) tile_move_lean<T, false, OpNone>(gs, gs, a.ax_ss, M, lg, OpNone(), false, RowSplitScratch{jb * M}, SlotNat<T>(), fM.NPAD(), tid, nthr, sB);
	}
}

template <class T, class StoreOp, bool LEAN = false>
DSP_DEV void split_inv_outer_thread(const SplitArgs &a, const FastDesc &fN, const StoreOp &sop, int gwarp, int lane) {
	const int group = gwarp % a.ngroups, i = gwarp / a.ngroups;
	if (i >= a.M) return;
	const int pair = group * 32 + lane;
	const int col = a.pcol0 + 2 * pair;
	if (col >= a.pcol0 + a.pcols) return;
	const bool hasb = col + 1 < a.pcol0 + a.pcols;
	const C2<T> *g0 = (const C2<T> *)a.scratch + pair + (long long)i * (a.ax_ss / 2);
	const long long js = (long long)a.M * (a.ax_ss / 2);
	C2<T> v[16], w[16];
	if (LEAN) {                                                 // the scratch panel is far below 2^31 elements
		const uint32_t js32 = (uint32_t)a.M * (uint32_t)(a.ax_ss / 2);
#pragma unroll
		for (int j = 0; j < 16; j++) v[j] = g0[(uint32_t)j * js32];
	} else {
#pragma unroll
		for (int j = 0; j < 16; j++) v[j] = g0[j * js];
	}
	if (i != 0) {
		tw_powers<T>((const C2<T> *)fN.tw, i, w);
#pragma unroll
		for (int j = 1; j < 16; j++) v[j] = cmul(v[j], w[j]);
	}
	Dft<T, 16>::run(v);
	GlobalCols<T, StoreOp, LEAN> sink;
	sink.p = (T *)a.out + col; sink.rs = a.ax_os; sink.hasb = hasb;
	sink.vec = hasb && (a.ax_os % 2) == 0 && (((size_t)a.out / sizeof(T) + col) % 2) == 0;
	sink.ax_slot = a.ax_slot; sink.op = &sop;
	sink.ca = a.cbase; sink.cb = a.cbase;
	{
		int x = col, ch = 0;
		if (a.d != 1) { x = (int)fd_div((uint32_t)col, a.dd); ch = col - x * a.d; }
		sink.ca.ch = ch; sink.ca.set(a.col_slot, x);
		x = col + 1; ch = 0;
		if (a.d != 1) { x = (int)fd_div((uint32_t)(col + 1), a.dd); ch = col + 1 - x * a.d; }
		sink.cb.ch = ch; sink.cb.set(a.col_slot, x);
	}
#pragma unroll
	for (int m = 0; m < 16; m++) sink.put(split_row(i + a.M * m, a.n), v[m].x, -v[m].y);
}

}  // namespace dsp
