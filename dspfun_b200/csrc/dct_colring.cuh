// dct_colring.cuh -- the split column pass (dct_split.cuh: n = 16 M, sub-FFTs of length M into an L2-resident scratch,
// then the outer radix-16 stage) as persistent TMA-fed kernels, float, n = 4096 | 8192, full 32-column tiles.
//
// Round 1's sub-pass kernels move their tiles with LDG / STG through registers: ncu shows them at 17-30% of DRAM
// throughput, waiting on the long scoreboard, with 1.15 waves of CTAs per 32 MB panel.  Here every byte between global
// and shared memory moves by tensor copy (cp.async.bulk.tensor, UTMALDG / UTMASTG), through the same ring as the row
// kernel (dct_ring.cuh): one CTA per SM, two thread groups, three 68 KB buffers, loads two iterations ahead.
//
// One launch walks all panels of the pass: A(0), A(1), B(0), A(2), B(1), ... -- B of a panel follows A of the next, so
// that by the time an item of B(q) is loaded every CTA has long finished A(q).  That order is a dependency, not a
// hope: every finished item bumps its panel's counter in global memory, and a load of B(q) (or of A(q + 3), which
// reuses B(q)'s scratch panel) is issued only once the counter is full.
//
// DCT-II, per column panel (scratch = [16][M][P] floats, stays in L2):
//   A  iteration = (sub-FFT j, 32-column tile).  Load: the image viewed as [n/32][32][cols]; sub-sequence j of the
//      Makhoul-permuted column is the rows of phase 2j (ascending) followed by the rows of phase 31 - 2j (descending):
//      two boxes {32 cols, 1, M/2}.  First radix-r0 pass from the raw boxes (registers across the barrier), padded
//      slots, radix-16 pass with twiddles, results laid out [k][column pair] and stored as boxes to scratch block j.
//   B  iteration = (32-column tile, block of 16 butterflies i).  Load: scratch rows (j, i) for i in the block and for
//      the mirror block M - i: two boxes {32 cols, 16, 16}.  A thread = (column pair, i) runs the fused outer
//      radix-16 + (k, n-k) twiddle of butterflies i and M - i in place (outputs k = i + M m land where input j = m was),
//      and the two boxes go straight to the image viewed as [16][M][cols].
//
// DCT-III mirrors it (pre-twiddle + sub-FFTs into the scratch, then the outer DIT butterflies): see the inverse section.
#pragma once
#include "dct_ring.cuh"
#include "dct_split.cuh"

namespace dsp {

static const int kColRingScratch = 3;      // scratch panels in rotation: A(q+1) fills one while B(q) reads another
static const int kColRingMaxPanels = 96;   // panels (outer index x column panel) one launch walks

struct ColRingArgs {
	TmaDesc in_map;       // A load : image   [planes][n/32][32][cols]   box {32, 1, M/2, 1}
	TmaDesc sc_st_map;    // A store: scratch [3][16][M][P]              box {32, min(M, 256), 1, 1}
	TmaDesc sc_ld_map;    // B load : scratch                            box {32, 16, 16, 1}
	TmaDesc sc_ld1_map;   // B load : scratch, butterfly M/2             box {32, 1, 16, 1}
	TmaDesc out_map;      // B store: image   [planes][16][M][cols]      box {32, 16, 16, 1}
	TmaDesc out1_map;     // B store: image, butterfly M/2               box {32, 1, 16, 1}
	int nplanes, ppp;     // outer indices (planes / batch elements) and column panels per plane
	int P, ncols;         // panel width (the last panel of a plane may be narrower); columns per plane
	int reverse;          // walk the panels of a plane from the last to the first (inverse plans)
	int *done;            // [2][nplanes * ppp] items finished per panel: sub-pass A | sub-pass B.  Zero before the launch.
	int flags;            // bit 0: do not discard dead scratch lines (DSP_DCT_RING_NODISCARD, A/B measurements)
	long long *trace;     // debug (DSP_DCT_RING_TRACE): per item of CTA 0, 4 globaltimer stamps; nullptr = off
	float *scratch;       // [3][16][M][P]: sub-pass B discards the rows it has read (no write-back of dead scratch lines to HBM)
	float *out;           // sub-pass B stores its results from registers (STG.64, 128 B per half warp): image base,
	long long ax_os, plane_os;   // row and plane strides (elements)
	float lscale, sscale;
	const void *twM;      // C2<float>[M]      sub-FFT twiddles
	const uint16_t *sigM; // [M]               sub-FFT slot table
	const void *twN, *omN;// tables of the n-point transform (outer pass)
	const uint16_t *sigN;
};

template <int LGM> struct ColGeom {
	typedef FastFixedBase<LGM> FM;
	enum {
		M = 1 << LGM, N = 16 * M, R0 = 1 << FM::kL0, BB = M / R0,
		NSEQ = 16,                                       // column pairs per tile
		P1R = (NSEQ * BB) / kRingGroup,                  // first-pass butterflies per thread
		P2R = (NSEQ * (M / 16)) / kRingGroup,            // radix-16 butterflies per thread
		SROWS = M < 256 ? M : 256,                       // rows per store box
		NBLK = M / 32,                                   // blocks of 16 butterflies in sub-pass B
		NPAD = FM::kNPAD,
	};
	static_assert(FM::kK == 0 && BB == 16 && P1R == 1, "sub-FFT geometry: M = r0 * 16, r0 in {16, 32}");
};

// ------------------------------------------------------------------------------------------------ A: sub-FFT j of a tile
template <int LGM>
DSP_DEV void colA_iter(const ColRingArgs &a, const RingFixed<LGM> &fM, C2<float> *buf, int group, int t0, int t1) {
	typedef ColGeom<LGM> G;
	const int M = G::M, R0 = G::R0, BB = G::BB;
	const C2<float> *h0 = buf, *h1 = buf + (M / 2) * 16;
#if DSP_GPU
	C2<float> v[R0 > 16 * G::P2R ? R0 : 16 * G::P2R];
#else
	static thread_local C2<float> v_all[kRingGroupThreads][R0 > 16 * G::P2R ? R0 : 16 * G::P2R];
#endif
	// ---- first radix-r0 pass from the raw boxes: element e' of the sub-sequence = row e' of box 0 (e' < M/2) | row
	//      M-1-e' of box 1
	for (int tid = t0; tid < t1 && tid < kRingGroup; tid++) {
		const int cp = tid & 15, r = tid >> 4;
		C2<float> *vv = RING_REGS(v, tid);
#pragma unroll
		for (int j = 0; j < R0; j++) {
			const int e = r + j * BB;                                // e < M/2 exactly when j < R0/2 (r < BB = 16)
			const C2<float> z = j < R0 / 2 ? h0[e * 16 + cp] : h1[(M - 1 - e) * 16 + cp];
			vv[j] = C2<float>{z.x * a.lscale, z.y * a.lscale};
		}
		Dft<float, R0>::run(vv);
	}
	RING_SYNC(group);
	for (int tid = t0; tid < t1 && tid < kRingGroup; tid++) {
		const int cp = tid & 15, r = tid >> 4;
		C2<float> *p = buf + cp * G::NPAD + (int)fM.s_sig[r];
		const C2<float> *vv = RING_REGS(v, tid);
#pragma unroll
		for (int m = 0; m < R0; m++) p[Pad<float>::of(m)] = vv[m];
	}
	RING_SYNC(group);
	// ---- radix-16 DIT pass with twiddles W_M^{i j}; the results leave the padded slots for the dense [k][column pair]
	//      layout the store boxes take, so they wait in registers for the barrier
	for (int tid = t0; tid < t1 && tid < kRingGroup; tid++) {
		const int cp = tid & 15;
#pragma unroll
		for (int rd = 0; rd < G::P2R; rd++) {
			const int i = (tid >> 4) + 16 * rd;
			const C2<float> *p = buf + cp * G::NPAD + Pad<float>::of(i);
			C2<float> *vv = RING_REGS(v, tid) + 16 * rd;
			C2<float> w[16];
			if (i != 0) fM.template tw_mid<float>(i, 0, w);
#pragma unroll
			for (int j = 0; j < 16; j++) vv[j] = p[fM.PO(0, j)];
			if (i != 0) {
#pragma unroll
				for (int j = 1; j < 16; j++) vv[j] = cmul(vv[j], w[j]);
			}
			Dft<float, 16>::run(vv);
		}
	}
	RING_SYNC(group);
	for (int tid = t0; tid < t1 && tid < kRingGroup; tid++) {
		const int cp = tid & 15;
#pragma unroll
		for (int rd = 0; rd < G::P2R; rd++) {
			const int i = (tid >> 4) + 16 * rd;
			const C2<float> *vv = RING_REGS(v, tid) + 16 * rd;
#pragma unroll
			for (int m = 0; m < 16; m++) buf[(i + R0 * m) * 16 + cp] = vv[m];
		}
	}
	RING_PROXY_FENCE();                                          // the buffer goes to the copy engine (tensor store) next
	RING_SYNC(group);
}

// ------------------------------------------------------------------------------------------------ B: outer pass of a block
// butterflies of a block as they sit in the boxes: input j / output m of butterfly i at [j][ii][column pair]
struct BoxBf {
	C2<float> *pa, *pb;          // element 0 of butterfly i (box 1) and of butterfly M - i (box 2)
	int js;                      // stride between j (complex elements)
	struct Row {
		C2<float> *p; int js;
		DSP_DEVM C2<float> get(int j) const { return p[j * js]; }
		DSP_DEVM void put(int j, C2<float> v) const { p[j * js] = v; }
	};
	DSP_DEVM Row row(int) const { return Row{pa, js}; }
	DSP_DEVM Row rowb(int) const { return Row{pb, js}; }
};
// results of the outer pass back into the boxes: k = i' + M m with i' = i (box 1) or M - i (box 2)
template <int LGM> struct BoxSink {
	C2<float> *pa, *pb;
	int js, i;
	float m;
	DSP_DEVM void put(int k, float xa, float xb) const {
		C2<float> *p = ((k & ((1 << LGM) - 1)) == i) ? pa : pb;
		p[(k >> LGM) * js] = C2<float>{xa * m, xb * m};
	}
};

// The butterflies' inputs go to registers first and the buffer is handed back at once (`released`, called by every
// thread after the barrier): the outer pass works in registers and stores its results to the image directly, so a
// sub-pass B item keeps its buffer only for the flight of its load -- all three buffers of the ring can be in flight.
template <int LGM, class Released>
DSP_DEV void colB_iter(const ColRingArgs &a, const RingFixed<LGM + 4> &fN, C2<float> *buf, int blk, int col, int pcol0, int sc, int plane,
                       int group, int t0, int t1, const Released &released) {
	typedef ColGeom<LGM> G;
	const int M = G::M;
#if DSP_GPU
	C2<float> va[16], vb[16];
#else
	static thread_local C2<float> va_all[kRingGroupThreads][16], vb_all[kRingGroupThreads][16];
#endif
#if DSP_GPU
	// The scratch rows of this item are dead: written once (sub-pass A), read once (the boxes have landed), rewritten
	// three panels later.  Without the discard the L2 writes most of them back to HBM (ncu r02: 986 MB of DRAM writes per
	// pass against 537 MB of results).  One 128-byte line per (j, i) row of the boxes; issued and fenced BEFORE the group
	// barrier, i.e. before the item is published: a late discard must not meet the panel's next contents.
	if (t0 < 32 && !(a.flags & 1)) {                            // one warp: 16 or 17 lines per lane, one fence per lane
		const int tile_col = col - pcol0;
		const float *sbase = a.scratch + (size_t)sc * 16 * M * a.P;
		for (int l = t0; l < 512; l += 32) {
			const int j = (l >> 4) & 15, ii = l & 15;
			const int row = l < 256 ? 16 * blk + ii : M - 16 * blk - 15 + ii;
			if (row < M) discard_l2(sbase + ((size_t)j * M + row) * a.P + tile_col);
		}
		if (blk == 0 && t0 < 16) discard_l2(sbase + ((size_t)t0 * M + M / 2) * a.P + tile_col);
		__threadfence();
	}
#endif
	for (int tid = t0; tid < t1 && tid < kRingGroup; tid++) {
		const int cp = tid & 15, ii = tid >> 4, i = 16 * blk + ii;
		const C2<float> *b1 = buf + ii * 16 + cp, *b2 = i != 0 ? buf + 4096 + (15 - ii) * 16 + cp : buf + 8192 + cp;
		const int js2 = i != 0 ? 256 : 16;                           // (butterfly 0 travels with butterfly M/2, from the third box)
		C2<float> *xa = RING_REGS(va, tid), *xb = RING_REGS(vb, tid);
#pragma unroll
		for (int j = 0; j < 16; j++) { xa[j] = b1[j * 256]; xb[j] = b2[j * js2]; }
	}
	RING_PROXY_FENCE();                                          // the buffer goes back to the copy engine (refill) next
	RING_SYNC(group);
	released();
	const OpMul<float> sop = {a.sscale};
	for (int tid = t0; tid < t1 && tid < kRingGroup; tid++) {
		const int cp = tid & 15, ii = tid >> 4, i = 16 * blk + ii;
		GlobalCols<float, OpMul<float>, true> sink;
		sink.p = a.out + (long long)plane * a.plane_os + col + 2 * cp; sink.rs = a.ax_os; sink.hasb = true; sink.vec = true;
		sink.ax_slot = 0; sink.ca = Coord{0, 0, 0, 0, 0}; sink.cb = sink.ca; sink.op = &sop;
		C2<float> *xa = RING_REGS(va, tid), *xb = RING_REGS(vb, tid);
		if (i != 0) dct2_outer_unit<float>(RegBf{xa, xb}, fN, i, sink);
		else {
			dct2_outer_unit<float>(RegBf{xa, xa}, fN, 0, sink);
			dct2_outer_unit<float>(RegBf{xb, xb}, fN, M / 2, sink);
		}
	}
}


// ================================================================================================ DCT-III
// The DIT-style inverse of dct_split.cuh on the same ring.  Per panel:
//   A''  item = (16-column tile, sub-sequence pair jj): rows k = 16 k' + ja and k = 16 k' + jb (jb = 16 - ja) of the image
//        viewed as [M][16][cols], boxes {16 cols, 1, <= 256}.  The pre-twiddle pairs row k with row n - k, i.e. element
//        k' of sub-sequence ja with element M-1-k' of jb (jj = 0, 8: within the one sub-sequence); each first-pass
//        butterfly forms its own inputs from both boxes (the pair's other output belongs to another thread, which forms
//        it again: 12 flops against a 64-register exchange).  Then the M-point DIT FFTs as in sub-pass A, results to
//        scratch blocks ja and jb.
//   B''  item = (32-column tile, block of 32 butterflies i): one box {32 cols, 32, 16} of scratch rows (j, i); a plain
//        radix-16 DIT butterfly per (column pair, i) in registers, sample e = i + M m stored (re, -im) to image row 2e |
//        2(n-1-e)+1 directly (STG.64, 128 B per half warp).  The buffer goes back as soon as the box is in registers.
struct InvTables {               // half-sample phases of the pre-twiddle, factored: (cos, sin)(pi k / 2n), k = 16 e + j
	const C2<float> *om16;       // [M + 1]  (cos, sin)(pi e / 2M)
	const C2<float> *omj;        // [17]     (cos, sin)(pi j / 2n)
	DSP_DEVM C2<float> at(int e, int j) const { return om_plus<float>(om16[e], omj[j]); }
};

template <int LGM, class Hook>
DSP_DEV void colA_inv_iter(const ColRingArgs &a, const RingFixed<LGM> &fM, const InvTables &it, C2<float> *buf, int jj, int group, int t0, int t1,
                           const Hook &) {
	typedef ColGeom<LGM> G;
	const int M = G::M, R0 = G::R0, BB = G::BB;
	const bool paired = jj != 0 && jj != 8;
#if DSP_GPU
	C2<float> v[R0 > 16 * G::P2R ? R0 : 16 * G::P2R];
#else
	static thread_local C2<float> v_all[kRingGroupThreads][R0 > 16 * G::P2R ? R0 : 16 * G::P2R];
#endif
	// ---- pre-twiddle + first radix-r0 pass from the raw boxes: raw_s[e][cp] at buf + (s M + e) 8 + cp
	for (int tid = t0; tid < t1 && tid < kRingGroup; tid++) {
		const int s = tid >> 7, cp = tid & 7, r = (tid >> 3) & 15;
		if (!paired && s) continue;
		const C2<float> *own = buf + s * M * 8 + cp, *oth = buf + (paired ? (1 - s) : 0) * M * 8 + cp;
		const int js = s ? 16 - jj : jj;                              // phase of this thread's sub-sequence
		C2<float> *vv = RING_REGS(v, tid);
#pragma unroll
		for (int j = 0; j < R0; j++) {
			const int e = r + j * BB;
			// partner element: k = 16 e + js pairs with n - k = 16 ep + (16 - js)  (jj = 0: 16 (M - e) + 0)
			const int ep = jj == 0 ? (e == 0 ? 0 : M - e) : M - 1 - e;
			C2<float> x = own[e * 8], y = oth[(jj == 0 && e == 0 ? 0 : ep) * 8];
			x = C2<float>{x.x * a.lscale, x.y * a.lscale}; y = C2<float>{y.x * a.lscale, y.y * a.lscale};
			C2<float> wk, wn;
			if (jj == 0 && e == 0) { vv[j] = C2<float>{x.x, -x.y}; continue; }                        // k = 0
			const bool low = e < M / 2 || (e == M / 2 && js == 0);                                    // k <= n - k
			if (low) { dct3_pair<float>(it.at(e, js), x, y, wk, wn); vv[j] = wk; }
			else { dct3_pair<float>(it.at(ep, jj == 0 ? 0 : 16 - js), y, x, wk, wn); vv[j] = wn; }
		}
		Dft<float, R0>::run(vv);
	}
	RING_SYNC(group);
	for (int tid = t0; tid < t1 && tid < kRingGroup; tid++) {
		const int s = tid >> 7, cp = tid & 7, r = (tid >> 3) & 15;
		if (!paired && s) continue;
		C2<float> *p = buf + (s * 8 + cp) * G::NPAD + (int)fM.s_sig[r];
		const C2<float> *vv = RING_REGS(v, tid);
#pragma unroll
		for (int m = 0; m < R0; m++) p[Pad<float>::of(m)] = vv[m];
	}
	RING_SYNC(group);
	// ---- radix-16 DIT pass with twiddles; dense [k'][cp] per sub-sequence for the store boxes
	for (int tid = t0; tid < t1 && tid < kRingGroup; tid++) {
		const int seq = tid & 15;
		if (!paired && seq >= 8) continue;
#pragma unroll
		for (int rd = 0; rd < G::P2R; rd++) {
			const int i = (tid >> 4) + 16 * rd;
			const C2<float> *p = buf + seq * G::NPAD + Pad<float>::of(i);
			C2<float> *vv = RING_REGS(v, tid) + 16 * rd;
			C2<float> w[16];
			if (i != 0) fM.template tw_mid<float>(i, 0, w);
#pragma unroll
			for (int j = 0; j < 16; j++) vv[j] = p[fM.PO(0, j)];
			if (i != 0) {
#pragma unroll
				for (int j = 1; j < 16; j++) vv[j] = cmul(vv[j], w[j]);
			}
			Dft<float, 16>::run(vv);
		}
	}
	RING_SYNC(group);
	for (int tid = t0; tid < t1 && tid < kRingGroup; tid++) {
		const int seq = tid & 15, s = seq >> 3, cp = seq & 7;
		if (!paired && s) continue;
#pragma unroll
		for (int rd = 0; rd < G::P2R; rd++) {
			const int i = (tid >> 4) + 16 * rd;
			const C2<float> *vv = RING_REGS(v, tid) + 16 * rd;
#pragma unroll
			for (int m = 0; m < 16; m++) buf[(s * M + i + R0 * m) * 8 + cp] = vv[m];
		}
	}
	RING_PROXY_FENCE();
	RING_SYNC(group);
}

template <int LGM, class Released>
DSP_DEV void colB_inv_iter(const ColRingArgs &a, C2<float> *buf, int blk, int col, int pcol0, int sc, int plane, int group, int t0, int t1,
                           const Released &released) {
	typedef ColGeom<LGM> G;
	const int M = G::M, n = 16 * M;
#if DSP_GPU
	C2<float> v[32];
	if (t0 < kRingGroup && !(a.flags & 1)) {                    // the scratch rows of the box are dead once it has landed
		const float *sbase = a.scratch + (size_t)sc * 16 * M * a.P + (col - pcol0);
#pragma unroll
		for (int r = 0; r < 2; r++) {
			const int l = t0 + 256 * r, j = l >> 5, ii = l & 31;
			discard_l2(sbase + ((size_t)j * M + 32 * blk + ii) * a.P);
		}
		if (t0 < 32) __threadfence();
	}
#else
	static thread_local C2<float> v_all[kRingGroupThreads][32];
#endif
	for (int tid = t0; tid < t1 && tid < kRingGroup; tid++) {
		const int cp = tid & 15, ii0 = tid >> 4;
		C2<float> *vv = RING_REGS(v, tid);
#pragma unroll
		for (int rd = 0; rd < 2; rd++)
#pragma unroll
			for (int j = 0; j < 16; j++) vv[16 * rd + j] = buf[j * 512 + (ii0 + 16 * rd) * 16 + cp];
	}
	RING_PROXY_FENCE();
	RING_SYNC(group);
	released();
	const C2<float> *tw = (const C2<float> *)a.twN;
	for (int tid = t0; tid < t1 && tid < kRingGroup; tid++) {
		const int cp = tid & 15, ii0 = tid >> 4;
		float *q = a.out + (long long)plane * a.plane_os + col + 2 * cp;
#pragma unroll
		for (int rd = 0; rd < 2; rd++) {
			const int i = 32 * blk + ii0 + 16 * rd;
			C2<float> *vv = RING_REGS(v, tid) + 16 * rd;
			if (i != 0) {
				C2<float> w[16];
				tw_powers<float>(tw, i, w);
#pragma unroll
				for (int j = 1; j < 16; j++) vv[j] = cmul(vv[j], w[j]);
			}
			Dft<float, 16>::run(vv);
#pragma unroll
			for (int m = 0; m < 16; m++) {
				const int e = i + M * m;
				const long long row = e < n / 2 ? 2 * e : 2 * (n - 1 - e) + 1;
				*(C2<float> *)(q + row * a.ax_os) = C2<float>{vv[m].x * a.sscale, -vv[m].y * a.sscale};
			}
		}
	}
}

// ------------------------------------------------------------------------------------------------ work list
// Panel q = plane * ppp + column panel.  Segments in launch order: A(0); then A(q), B(q-1) for q = 1..Q-1; then B(Q-1).
struct ColItem {
	int sub_b;            // 0: sub-pass A, 1: sub-pass B
	int q;                // panel
	int local;            // item within the segment
	int plane, col0, ntiles;
};
template <int LGM, bool INV = false> struct ColWork {
	typedef ColGeom<LGM> G;
	enum { kItemsA = INV ? 9 : 16, kTileA = INV ? 16 : 32, kItemsB = INV ? (int)G::M / 32 : (int)G::NBLK };
	DSP_HDM static int panels(const ColRingArgs &a) { return a.nplanes * a.ppp; }
	DSP_HDM static int ntiles(const ColRingArgs &a, int q) {
		int p = q % a.ppp;
		if (a.reverse) p = a.ppp - 1 - p;
		const int left = a.ncols - p * a.P;
		return (left < a.P ? left : a.P) / 32;
	}
	DSP_HDM static int seg_panel(int s, int Q) { return s == 0 ? 0 : (s == 2 * Q - 1 ? Q - 1 : ((s & 1) ? (s + 1) / 2 : s / 2 - 1)); }
	DSP_HDM static int seg_is_b(int s, int Q) { return s == 0 ? 0 : (s == 2 * Q - 1 ? 1 : ((s & 1) ? 0 : 1)); }
	// ntiles() counts 32-column tiles; sub-pass A'' of the inverse works on 16-column tiles
	DSP_HDM static int seg_items(const ColRingArgs &a, int s, int Q) { return ntiles(a, seg_panel(s, Q)) * (seg_is_b(s, Q) ? (int)kItemsB : (int)kItemsA * (32 / (int)kTileA)); }
	// seg_start[s] = first global item of segment s (2Q + 1 entries, filled once per CTA)
	DSP_HDM static void decode(const ColRingArgs &a, const int *seg_start, int gi, int &cursor, ColItem &w) {
		const int Q = panels(a);
		while (gi < seg_start[cursor]) cursor--;                  // (a deferred first load is issued after later ones)
		while (gi >= seg_start[cursor + 1]) cursor++;
		w.sub_b = seg_is_b(cursor, Q);
		w.q = seg_panel(cursor, Q);
		w.local = gi - seg_start[cursor];
		w.plane = w.q / a.ppp;
		int p = w.q % a.ppp;
		if (a.reverse) p = a.ppp - 1 - p;
		w.col0 = p * a.P;
		w.ntiles = ntiles(a, w.q);
	}
};

// ------------------------------------------------------------------------------------------------ CTA bodies
template <int LGM> struct ColRingSmem {
	typedef RingSmem<LGM + 4> S;                                        // same buffers as the row ring at n = 16 M
	typedef RingFixed<LGM> FM;
	static constexpr size_t kBufBytes = S::kBufBytes;
	static constexpr int kBufStride = S::kBufStride;
	static constexpr size_t kTabNBytes = S::kTablesBytes;                                                // tables of the n-point outer pass
	static constexpr size_t kTabMBytes = ((size_t)(4 * (1 << FM::kL0) + 8) * sizeof(C2<float>) + 127) / 128 * 128;   // sub-FFT: s_mid, s_sig
	static constexpr size_t kSegBytes = ((size_t)(2 * kColRingMaxPanels + 2) * sizeof(int) + 127) / 128 * 128;
	static constexpr size_t kTablesBytes = kTabNBytes + kTabMBytes + kSegBytes;
	static constexpr size_t kTotal = kTablesBytes + kRingBufs * kBufBytes + 64;
	static_assert((size_t)ColGeom<LGM>::NSEQ * ColGeom<LGM>::NPAD * sizeof(C2<float>) <= kBufBytes, "tile does not fit the ring buffer");
	static_assert((size_t)(8192 + 16 * 16) * sizeof(C2<float>) <= kBufBytes, "outer-pass boxes do not fit the ring buffer");
	static_assert(kTotal <= 227 * 1024, "shared memory budget");
};

// tables of the M-point sub-FFT in shared memory (only the radix-16 pass's twiddles and the slot table are used)
template <int LGM>
DSP_DEV void colA_fill_tables(const ColRingArgs &a, C2<float> *tab, RingFixed<LGM> &f, int t0, int t1, int nthr) {
	typedef RingFixed<LGM> F;
	const int R0 = 1 << F::kL0, M = 1 << LGM;
	C2<float> *s_mid = tab;
	uint16_t *s_sig = (uint16_t *)(s_mid + 4 * R0);
	const C2<float> *tw = (const C2<float> *)a.twM;
	for (int tid = t0; tid < t1; tid++) {
		for (int i = tid; i < R0; i += nthr) {
#pragma unroll
			for (int p = 0; p < 4; p++) s_mid[p * R0 + i] = ldg_c2(tw + (((i << p) * (M / (16 * R0))) & (M - 1)));
		}
		for (int i = tid; i < M / R0; i += nthr) s_sig[i] = DSP_LDG(a.sigM + i);
	}
	f.tw = a.twM; f.om = nullptr; f.sig = a.sigM;
	f.s_out = nullptr; f.s_om = nullptr; f.s_mid = s_mid; f.s_sig = s_sig;
}

// factored half-sample phases of the inverse pre-twiddle, from the n-point table om[k] = (cos, sin)(pi k / 2n), k <= n/2:
// om16[e] = angle pi 16 e / 2n, e <= M: om[16 e] for e <= M/2, mirrored above (cos(pi/2 - x) = sin x)
template <int LGM>
DSP_DEV void col_inv_fill_tables(const ColRingArgs &a, C2<float> *tab, InvTables &t, int t0, int t1, int nthr) {
	const int M = 1 << LGM;
	const C2<float> *om = (const C2<float> *)a.omN;
	C2<float> *om16 = tab, *omj = tab + (M + 1);
	for (int tid = t0; tid < t1; tid++) {
		for (int e = tid; e <= M; e += nthr) {
			if (2 * e <= M) om16[e] = ldg_c2(om + 16 * e);
			else { const C2<float> m = ldg_c2(om + 16 * (M - e)); om16[e] = C2<float>{m.y, m.x}; }   // angle pi/2 - pi 16 (M - e) / 2n
		}
		for (int j = tid; j < 17; j += nthr) omj[j] = ldg_c2(om + j);
	}
	t.om16 = om16; t.omj = omj;
}

template <int LGM, bool INV>
DSP_DEV void col_fill_segments(const ColRingArgs &a, int *seg_start) {
	const int Q = ColWork<LGM, INV>::panels(a);
	int acc = 0;
	for (int s = 0; s < 2 * Q; s++) { seg_start[s] = acc; acc += ColWork<LGM, INV>::seg_items(a, s, Q); }
	seg_start[2 * Q] = acc;
}

// ---- loads / stores of an item
template <int LGM, bool INV, class Bar>
DSP_DEV void col_load(const ColRingArgs &a, const ColItem &w, C2<float> *buf, Bar *bar) {
	typedef ColGeom<LGM> G;
	const int sc = w.q % kColRingScratch;
	if (INV) {
		const auto once = l2_policy_evict_first();
		if (!w.sub_b) {                      // 16-column tile, sub-sequence pair jj: rows of phase ja (and jb) of the image [M][16][cols]
			const int nt = 2 * w.ntiles, jj = w.local / nt, tile = w.local - jj * nt;
			const int c = w.col0 + 16 * tile;
			for (int s = 0; s < (jj == 0 || jj == 8 ? 1 : 2); s++)
				for (int h = 0; h < G::M / G::SROWS; h++)
					tma_load4_hint(buf + (s * G::M + h * G::SROWS) * 8, &a.in_map, c, s ? 16 - jj : jj, h * G::SROWS, w.plane, bar, once);
		} else {                             // 32-column tile, block of 32 butterflies: scratch rows (j, i)
			const int blk = w.local / w.ntiles, tile = w.local - blk * w.ntiles;
			tma_load4_hint(buf, &a.sc_ld_map, 32 * tile, 32 * blk, 0, sc, bar, once);
		}
		return;
	}
	if (!w.sub_b) {                          // sub-FFT j of a tile: rows of phase 2j ascending, rows of phase 31 - 2j descending
		const int j = w.local / w.ntiles, tile = w.local - j * w.ntiles;
		const int c = w.col0 + 32 * tile;
		const auto stream = l2_policy_evict_first();                 // image data passes through once: leave L2 to the scratch
		tma_load4_hint(buf, &a.in_map, c, 2 * j, 0, w.plane, bar, stream);
		tma_load4_hint(buf + (G::M / 2) * 16, &a.in_map, c, 31 - 2 * j, 0, w.plane, bar, stream);
	} else {                                 // block of 16 butterflies i and the mirror block M - i (+ butterfly M/2 with block 0)
		const int blk = w.local / w.ntiles, tile = w.local - blk * w.ntiles;
		const auto last_use = l2_policy_evict_first();               // the scratch rows are dead once read
		tma_load4_hint(buf, &a.sc_ld_map, 32 * tile, 16 * blk, 0, sc, bar, last_use);
		tma_load4_hint(buf + 4096, &a.sc_ld_map, 32 * tile, G::M - 16 * blk - 15, 0, sc, bar, last_use);
		if (blk == 0) tma_load4_hint(buf + 8192, &a.sc_ld1_map, 32 * tile, G::M / 2, 0, sc, bar, last_use);
	}
}
template <int LGM, bool INV> DSP_DEV uint32_t col_load_bytes(const ColItem &w) {
	if (INV) {
		if (w.sub_b) return 65536u;
		const int jj = w.local / (2 * w.ntiles);
		return (uint32_t)((jj == 0 || jj == 8 ? 1 : 2) * ColGeom<LGM>::M * 8 * sizeof(C2<float>));
	}
	if (!w.sub_b) return (uint32_t)(ColGeom<LGM>::M * 16 * sizeof(C2<float>));
	return 2u * 32768u + ((w.local / w.ntiles) == 0 ? 2048u : 0u);
}
template <int LGM, bool INV>
DSP_DEV void col_store(const ColRingArgs &a, const ColItem &w, const C2<float> *buf) {
	typedef ColGeom<LGM> G;
	const int sc = w.q % kColRingScratch;
	if (INV) {
		if (w.sub_b) return;
		const int nt = 2 * w.ntiles, jj = w.local / nt, tile = w.local - jj * nt;
		const auto keep = l2_policy_evict_last();
		for (int s = 0; s < (jj == 0 || jj == 8 ? 1 : 2); s++)
			for (int h = 0; h < G::M / G::SROWS; h++)
				tma_store4_hint(&a.sc_st_map, buf + (s * G::M + h * G::SROWS) * 8, 16 * tile, h * G::SROWS, s ? 16 - jj : jj, sc, keep);
		return;
	}
	if (!w.sub_b) {
		const int j = w.local / w.ntiles, tile = w.local - j * w.ntiles;
		const auto keep = l2_policy_evict_last();                    // the scratch is read back a segment later: keep it in L2
		for (int h = 0; h < G::M / G::SROWS; h++) tma_store4_hint(&a.sc_st_map, buf + h * G::SROWS * 16, 32 * tile, h * G::SROWS, j, sc, keep);
	}
}
// what has to be finished before the item's boxes may be loaded: (counter index, count), or index -1
template <int LGM, bool INV> DSP_DEV void col_dependency(const ColRingArgs &a, const ColItem &w, int &idx, int &need) {
	const int Q = ColWork<LGM, INV>::panels(a);
	if (w.sub_b) { idx = w.q; need = (int)ColWork<LGM, INV>::kItemsA * (32 / (int)ColWork<LGM, INV>::kTileA) * w.ntiles; }                                      // all sub-FFTs of the panel are in the scratch
	else if (w.q >= kColRingScratch) { idx = Q + w.q - kColRingScratch; need = (int)ColWork<LGM, INV>::kItemsB * ColWork<LGM, INV>::ntiles(a, w.q - kColRingScratch); }   // the scratch panel's previous user is done with it
	else { idx = -1; need = 0; }
}

#if DSP_GPU
DSP_DEV int ld_acquire(const int *p) {
	int v;
	asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
DSP_DEV long long gtimer() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
DSP_DEV void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

template <int LGM, bool INV>
DSP_DEV void col_issue(const ColRingArgs &a, const int *seg_start, int gi, int &cursor, C2<float> *buf, uint64_t *bar) {
	ColItem w;
	ColWork<LGM, INV>::decode(a, seg_start, gi, cursor, w);
	int idx, need;
	col_dependency<LGM, INV>(a, w, idx, need);
	if (idx >= 0) {
		while (ld_acquire(a.done + idx) < need) __nanosleep(64);
		fence_proxy_async_all();                                 // the boxes are read by the async proxy: order it after the acquire
	}
	mbar_expect_tx(bar, col_load_bytes<LGM, INV>(w));
	col_load<LGM, INV>(a, w, buf, bar);
}

// The ring.  A sub-pass A item ends by handing its buffer to the copy engine (tensor store to the scratch), waits for
// the store, publishes the item in its panel's counter and refills the buffer with the item three ahead.  A sub-pass B
// item gives its buffer back as soon as its boxes are in registers (and publishes that the scratch rows are read).
// Why the counters cannot deadlock: an item is published by the thread that ran it, right at its end, whatever else
// happens; a load of item y is issued by the thread group of items y - 3 / y - 1 and waits for items of an EARLIER
// SEGMENT.  The launcher sizes the grid so that every segment holds at least one item of every CTA, so those items
// are y - 2 or older in this CTA: either already done by this thread, or being run by the other thread group, which
// never waits for this one.  Items 0 and 1 of a CTA belong to the first two segments, which have no dependency.
template <int LGM, bool INV>
DSP_DEV void colring_cta(const ColRingArgs &a, unsigned char *smem, int cta, int ncta, int tid) {
	typedef ColRingSmem<LGM> S;
	C2<float> *tabN = (C2<float> *)smem;
	C2<float> *tabM = (C2<float> *)(smem + S::kTabNBytes);
	int *seg_start = (int *)(smem + S::kTabNBytes + S::kTabMBytes);
	C2<float> *bufs = (C2<float> *)(smem + S::kTablesBytes);
	uint64_t *full = (uint64_t *)(smem + S::kTablesBytes + kRingBufs * S::kBufBytes);
	RingFixed<LGM> fM;
	RingFixed<LGM + 4> fN;
	const int nthr = kRingGroups * kRingGroupThreads;
	colA_fill_tables<LGM>(a, tabM, fM, tid, tid + 1, nthr);
	InvTables itab;
	if (INV) col_inv_fill_tables<LGM>(a, tabN, itab, tid, tid + 1, nthr);
	else {
		RingArgs ra;
		ra.tw = a.twN; ra.om = a.omN; ra.sig = a.sigN;
		ring_fill_tables<LGM + 4>(ra, tabN, fN, tid, tid + 1, nthr);
	}
	if (tid == 0) {
		col_fill_segments<LGM, INV>(a, seg_start);
		for (int b = 0; b < kRingBufs; b++) mbar_init(full + b, 1);
		mbar_fence_init();
	}
	__syncthreads();
	const int total = seg_start[2 * ColWork<LGM, INV>::panels(a)];
	const int iters = (total - cta + ncta - 1) / ncta;
	// The first items are loaded before anything is computed: only those without a dependency (with >= one item per CTA
	// in every segment that is items 0 and 1 at least); a third one that has to wait is issued at the top of its own turn.
	int deferred = -1;
	if (tid == 0) {
		int cur = 0;
		for (int it = 0; it < kRingBufs && it < iters; it++) {
			ColItem w;
			int c2 = cur, idx, need;
			ColWork<LGM, INV>::decode(a, seg_start, cta + it * ncta, c2, w);
			col_dependency<LGM, INV>(a, w, idx, need);
			if (idx >= 0 && it == kRingBufs - 1) { deferred = it; break; }
			col_issue<LGM, INV>(a, seg_start, cta + it * ncta, cur, bufs + (size_t)it * S::kBufStride, full + it);
		}
	}
	const int group = tid / kRingGroupThreads, gt = tid - group * kRingGroupThreads;
	int cursor = 0, issue_cursor = 0;
	for (int it = group; it < iters; it += kRingGroups) {
		const int b = it % kRingBufs, gi = cta + it * ncta;
		C2<float> *buf = bufs + (size_t)b * S::kBufStride;
		if (tid == 0 && it == deferred) col_issue<LGM, INV>(a, seg_start, gi, issue_cursor, buf, full + b);       // (tid 0 is thread 0 of group 0: item 2 is its own)
		const bool tr = a.trace && cta == 0 && gt == 0 && it < 4096;
		if (tr) a.trace[4 * it + 0] = gtimer();
		mbar_wait(full + b, (uint32_t)((it / kRingBufs) & 1));
		if (tr) a.trace[4 * it + 1] = gtimer();
		ColItem w;
		ColWork<LGM, INV>::decode(a, seg_start, gi, cursor, w);
		const bool more = it + kRingBufs < iters;
		if (!w.sub_b) {
			if (INV) colA_inv_iter<LGM>(a, fM, itab, buf, w.local / (2 * w.ntiles), group, gt, gt + 1, []() {});
			else colA_iter<LGM>(a, fM, buf, group, gt, gt + 1);
			// the item ended on a group barrier: every result is in the buffer.  Store it, and PUBLISH it as soon as its boxes
			// are in global memory: the thread waits for the store here (the rest of the group is already at the next item;
			// the pass waits on memory, not on this thread).  Publishing at once is what lets the panels be small enough for
			// their scratch to stay in L2: a load then only ever waits for items at least two places back in its own CTA.
			if (gt == 0) {
				if (tr) a.trace[4 * it + 2] = gtimer();
				fence_proxy_async();
				col_store<LGM, INV>(a, w, buf);
				tma_commit();
				tma_wait_all0();
				if (tr) a.trace[4 * it + 3] = gtimer();
				fence_proxy_async_all();                             // the boxes were written by the async proxy: order them before
				__threadfence();                                     // the (generic-proxy) release of the counter
				atomicAdd(a.done + w.q, 1);
				if (more) col_issue<LGM, INV>(a, seg_start, cta + (it + kRingBufs) * ncta, issue_cursor, buf, full + b);
			}
		} else {
			const int blk = w.local / w.ntiles, tile = w.local - blk * w.ntiles;
			auto released = [&]() {
				if (gt != 0) return;
				if (tr) a.trace[4 * it + 2] = -gtimer();                 // (negative: a sub-pass B item, stamp = buffer released)
				atomicAdd(a.done + ColWork<LGM, INV>::panels(a) + w.q, 1);     // the scratch rows of this item have been read
				if (more) {
					fence_proxy_async();
					col_issue<LGM, INV>(a, seg_start, cta + (it + kRingBufs) * ncta, issue_cursor, buf, full + b);
				}
			};
			if (INV) colB_inv_iter<LGM>(a, buf, blk, w.col0 + 32 * tile, w.col0, w.q % kColRingScratch, w.plane, group, gt, gt + 1, released);
			else colB_iter<LGM>(a, fN, buf, blk, w.col0 + 32 * tile, w.col0, w.q % kColRingScratch, w.plane, group, gt, gt + 1, released);
		}
	}
}

template <int LGM, bool INV>
__global__ void __launch_bounds__(kRingGroups *kRingGroupThreads, 1) k_col_ring(const __grid_constant__ ColRingArgs a) {
	extern __shared__ __align__(128) unsigned char ring_smem[];
	colring_cta<LGM, INV>(a, ring_smem, (int)blockIdx.x, (int)gridDim.x, (int)threadIdx.x);
}
#else
template <int LGM, bool INV>
static void colring_emulate(const ColRingArgs &a, int ncta) {
	typedef ColRingSmem<LGM> S;
	std::vector<unsigned char> smem(S::kTotal + 128);
	C2<float> *tabN = (C2<float> *)smem.data();
	C2<float> *tabM = (C2<float> *)(smem.data() + S::kTabNBytes);
	int *seg_start = (int *)(smem.data() + S::kTabNBytes + S::kTabMBytes);
	C2<float> *buf = (C2<float> *)(smem.data() + S::kTablesBytes);
	RingFixed<LGM> fM;
	RingFixed<LGM + 4> fN;
	colA_fill_tables<LGM>(a, tabM, fM, 0, kRingGroupThreads, kRingGroupThreads);
	InvTables itab;
	if (INV) col_inv_fill_tables<LGM>(a, tabN, itab, 0, kRingGroupThreads, kRingGroupThreads);
	else {
		RingArgs ra;
		ra.tw = a.twN; ra.om = a.omN; ra.sig = a.sigN;
		ring_fill_tables<LGM + 4>(ra, tabN, fN, 0, kRingGroupThreads, kRingGroupThreads);
	}
	col_fill_segments<LGM, INV>(a, seg_start);
	const int Q = ColWork<LGM, INV>::panels(a), total = seg_start[2 * Q];
	// the emulation runs the items in launch order (which satisfies every dependency) and checks the counters it would
	// have waited for
	int cursor = 0;
	for (int gi = 0; gi < total; gi++) {
		(void)ncta;
		ColItem w;
		ColWork<LGM, INV>::decode(a, seg_start, gi, cursor, w);
		int idx, need;
		col_dependency<LGM, INV>(a, w, idx, need);
		if (idx >= 0 && a.done[idx] < need) abort();           // launch order must satisfy the dependencies
		col_load<LGM, INV>(a, w, buf, (void *)nullptr);
		const int blk = w.local / w.ntiles, bcol = w.col0 + 32 * (w.local % w.ntiles);
		if (!w.sub_b) {
			if (INV) colA_inv_iter<LGM>(a, fM, itab, buf, w.local / (2 * w.ntiles), 0, 0, kRingGroupThreads, []() {});
			else colA_iter<LGM>(a, fM, buf, 0, 0, kRingGroupThreads);
		} else if (INV) colB_inv_iter<LGM>(a, buf, blk, bcol, w.col0, w.q % kColRingScratch, w.plane, 0, 0, kRingGroupThreads, []() {});
		else colB_iter<LGM>(a, fN, buf, blk, bcol, w.col0, w.q % kColRingScratch, w.plane, 0, 0, kRingGroupThreads, []() {});
		col_store<LGM, INV>(a, w, buf);
		a.done[(w.sub_b ? Q : 0) + w.q] += 1;
	}
}
#endif

}  // namespace dsp
