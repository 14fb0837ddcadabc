// dct_colring.cuh -- the split column pass (dct_split.cuh: n = 16 M, sub-FFTs of length M into an L2-resident scratch,
// then the outer radix-16 stage) as persistent TMA-fed kernels, float, n = 4096 | 8192, full 32-column tiles.
//
// Round 1's sub-pass kernels move their tiles with LDG / STG through registers: ncu shows them at 17-30% of DRAM
// throughput, waiting on the long scoreboard, with 1.15 waves of CTAs per 32 MB panel.  Here every byte between global
// and shared memory moves by tensor copy (cp.async.bulk.tensor, UTMALDG / UTMASTG), through the same ring as the row
// kernel (dct_ring.cuh): one CTA per SM, two thread groups, three 68 KB buffers, loads two iterations ahead.
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

namespace dsp {

struct ColRingArgs {
	TmaDesc in_map;       // A load : image   [n/32][32][cols]   box {32, 1, M/2}
	TmaDesc sc_st_map;    // A store: scratch [16][M][P]         box {32, min(M, 256), 1}
	TmaDesc sc_ld_map;    // B load : scratch [16][M][P]         box {32, 16, 16}
	TmaDesc sc_ld1_map;   // B load : scratch, butterfly M/2     box {32, 1, 16}
	TmaDesc out_map;      // B store: image   [16][M][cols]      box {32, 16, 16}
	TmaDesc out1_map;     // B store: image, butterfly M/2       box {32, 1, 16}
	int col0;             // first column of the panel (image maps; the scratch maps start at 0)
	int ntiles;           // 32-column tiles in the panel
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

template <int LGM>
DSP_DEV void colB_iter(const ColRingArgs &a, const RingFixed<LGM + 4> &fN, C2<float> *buf, int blk, int group, int t0, int t1) {
	typedef ColGeom<LGM> G;
	const int M = G::M;
	for (int tid = t0; tid < t1 && tid < kRingGroup; tid++) {
		const int cp = tid & 15, ii = tid >> 4, i = 16 * blk + ii;
		C2<float> *b1 = buf + ii * 16 + cp, *b2 = buf + 4096 + (15 - ii) * 16 + cp, *b3 = buf + 8192 + cp;
		if (i != 0) {
			const BoxSink<LGM> sink{b1, b2, 256, i, a.sscale};
			dct2_outer_unit<float>(BoxBf{b1, b2, 256}, fN, i, sink);
		} else {
			const BoxSink<LGM> s0{b1, b1, 256, 0, a.sscale};
			dct2_outer_unit<float>(BoxBf{b1, b1, 256}, fN, 0, s0);
			const BoxSink<LGM> sh{b3, b3, 16, M / 2, a.sscale};
			dct2_outer_unit<float>(BoxBf{b3, b3, 16}, fN, M / 2, sh);
		}
	}
	RING_SYNC(group);
}

// ------------------------------------------------------------------------------------------------ CTA bodies
template <int LGM> struct ColRingSmem {
	typedef RingSmem<LGM + 4> S;                                        // same buffers as the row ring at n = 16 M
	static constexpr size_t kTablesBytes = S::kTablesBytes, kBufBytes = S::kBufBytes, kTotal = S::kTotal;
	static constexpr int kBufStride = S::kBufStride;
	static_assert((size_t)ColGeom<LGM>::NSEQ * ColGeom<LGM>::NPAD * sizeof(C2<float>) <= kBufBytes, "tile does not fit the ring buffer");
	static_assert((size_t)(8192 + 16 * 16) * sizeof(C2<float>) <= kBufBytes, "outer-pass boxes do not fit the ring buffer");
};

// tables of the M-point sub-FFT in shared memory (only the radix-16 pass's twiddles and the slot table are used)
template <int LGM>
DSP_DEV void colA_fill_tables(const ColRingArgs &a, C2<float> *tab, RingFixed<LGM> &f, int t0, int t1, int nthr) {
	RingArgs ra;
	ra.tw = a.twM; ra.om = nullptr; ra.sig = a.sigM;
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

// iteration gi of sub-pass A: sub-FFT j = gi / ntiles, tile = gi % ntiles
template <int LGM, class Bar>
DSP_DEV void colA_load(const ColRingArgs &a, C2<float> *buf, int gi, Bar *bar) {
	typedef ColGeom<LGM> G;
	const int j = gi / a.ntiles, tile = gi - j * a.ntiles;
	const int c = a.col0 + 32 * tile;
	tma_load3(buf, &a.in_map, c, 2 * j, 0, bar);
	tma_load3(buf + (G::M / 2) * 16, &a.in_map, c, 31 - 2 * j, 0, bar);
}
template <int LGM>
DSP_DEV void colA_store(const ColRingArgs &a, const C2<float> *buf, int gi) {
	typedef ColGeom<LGM> G;
	const int j = gi / a.ntiles, tile = gi - j * a.ntiles;
	for (int h = 0; h < G::M / G::SROWS; h++) tma_store3(&a.sc_st_map, buf + h * G::SROWS * 16, 32 * tile, h * G::SROWS, j);
}
// iteration gi of sub-pass B: block = gi / ntiles, tile = gi % ntiles
template <int LGM, class Bar>
DSP_DEV void colB_load(const ColRingArgs &a, C2<float> *buf, int gi, Bar *bar) {
	typedef ColGeom<LGM> G;
	const int blk = gi / a.ntiles, tile = gi - blk * a.ntiles;
	tma_load3(buf, &a.sc_ld_map, 32 * tile, 16 * blk, 0, bar);
	tma_load3(buf + 4096, &a.sc_ld_map, 32 * tile, G::M - 16 * blk - 15, 0, bar);
	if (blk == 0) tma_load3(buf + 8192, &a.sc_ld1_map, 32 * tile, G::M / 2, 0, bar);
}
template <int LGM>
DSP_DEV void colB_store(const ColRingArgs &a, const C2<float> *buf, int gi) {
	typedef ColGeom<LGM> G;
	const int blk = gi / a.ntiles, tile = gi - blk * a.ntiles;
	const int c = a.col0 + 32 * tile;
	tma_store3(&a.out_map, buf, c, 16 * blk, 0);
	tma_store3(&a.out_map, buf + 4096, c, G::M - 16 * blk - 15, 0);
	if (blk == 0) tma_store3(&a.out1_map, buf + 8192, c, G::M / 2, 0);
}
template <int LGM> DSP_DEV uint32_t colB_bytes(const ColRingArgs &a, int gi) { return 2u * 32768u + ((gi / a.ntiles) == 0 ? 2048u : 0u); }

#if DSP_GPU
// The ring with tensor-copy stores: an iteration ends by handing its buffer to the copy engine; the refill of that
// buffer (iteration it + 3) waits until the engine has read it, which the issuing thread checks one barrier into its
// next iteration -- by then the store is long under way, and the load still lands before the other group needs it.
template <int LGM, bool SUBA>
DSP_DEV void colring_cta(const ColRingArgs &a, unsigned char *smem, int cta, int ncta, int tid) {
	typedef ColRingSmem<LGM> S;
	typedef ColGeom<LGM> G;
	C2<float> *tab = (C2<float> *)smem;
	C2<float> *bufs = (C2<float> *)(smem + S::kTablesBytes);
	uint64_t *full = (uint64_t *)(smem + S::kTablesBytes + kRingBufs * S::kBufBytes);
	RingFixed<LGM> fM;
	RingFixed<LGM + 4> fN;
	if (SUBA) colA_fill_tables<LGM>(a, tab, fM, tid, tid + 1, kRingGroups * kRingGroupThreads);
	else {
		RingArgs ra;
		ra.tw = a.twN; ra.om = a.omN; ra.sig = a.sigN;
		ring_fill_tables<LGM + 4>(ra, tab, fN, tid, tid + 1, kRingGroups * kRingGroupThreads);
	}
	const int total = a.ntiles * (SUBA ? 16 : G::NBLK);
	const int iters = (total - cta + ncta - 1) / ncta;
	if (tid == 0) {
		for (int b = 0; b < kRingBufs; b++) mbar_init(full + b, 1);
		mbar_fence_init();
	}
	__syncthreads();
	if (tid == 0) {
		for (int it = 0; it < kRingBufs && it < iters; it++) {
			const int gi = cta + it * ncta;
			C2<float> *buf = bufs + (size_t)it * S::kBufStride;
			if (SUBA) { mbar_expect_tx(full + it, (uint32_t)(G::M * 16 * sizeof(C2<float>))); colA_load<LGM>(a, buf, gi, full + it); }
			else { mbar_expect_tx(full + it, colB_bytes<LGM>(a, gi)); colB_load<LGM>(a, buf, gi, full + it); }
		}
	}
	const int group = tid / kRingGroupThreads, gt = tid - group * kRingGroupThreads;
	int refill_it = -1;                                          // iteration whose buffer waits for its store before the refill
	for (int it = group; it < iters; it += kRingGroups) {
		const int b = it % kRingBufs, gi = cta + it * ncta;
		C2<float> *buf = bufs + (size_t)b * S::kBufStride;
		mbar_wait(full + b, (uint32_t)((it / kRingBufs) & 1));
		if (refill_it >= 0) {
			// (the first barrier of the iteration sits inside *_iter; issuing here, before it, costs the thread only the
			// wait for a store that was handed over a whole mbarrier wait ago)
			if (gt == 0) {
				tma_wait_read0();
				const int nb = refill_it % kRingBufs, ngi = cta + (refill_it + kRingBufs) * ncta;
				C2<float> *nbuf = bufs + (size_t)nb * S::kBufStride;
				if (SUBA) { mbar_expect_tx(full + nb, (uint32_t)(G::M * 16 * sizeof(C2<float>))); colA_load<LGM>(a, nbuf, ngi, full + nb); }
				else { mbar_expect_tx(full + nb, colB_bytes<LGM>(a, ngi)); colB_load<LGM>(a, nbuf, ngi, full + nb); }
			}
			refill_it = -1;
		}
		if (SUBA) colA_iter<LGM>(a, fM, buf, group, gt, gt + 1);
		else colB_iter<LGM>(a, fN, buf, gi / a.ntiles, group, gt, gt + 1);
		// the iteration ended on a group barrier: every result is in the buffer
		if (gt == 0) {
			fence_proxy_async();
			if (SUBA) colA_store<LGM>(a, buf, gi); else colB_store<LGM>(a, buf, gi);
			tma_commit();
		}
		if (it + kRingBufs < iters) refill_it = it;
	}
	if (gt == 0) tma_wait_all0();                               // shared memory must outlive the engine's reads
}

template <int LGM, bool SUBA>
__global__ void __launch_bounds__(kRingGroups *kRingGroupThreads, 1) k_col_ring(const __grid_constant__ ColRingArgs a) {
	extern __shared__ __align__(128) unsigned char ring_smem[];
	colring_cta<LGM, SUBA>(a, ring_smem, (int)blockIdx.x, (int)gridDim.x, (int)threadIdx.x);
}
#else
template <int LGM, bool SUBA>
static void colring_emulate(const ColRingArgs &a, int ncta) {
	typedef ColRingSmem<LGM> S;
	typedef ColGeom<LGM> G;
	std::vector<unsigned char> smem(S::kTotal + 128);
	C2<float> *tab = (C2<float> *)smem.data();
	C2<float> *buf = (C2<float> *)(smem.data() + S::kTablesBytes);
	RingFixed<LGM> fM;
	RingFixed<LGM + 4> fN;
	if (SUBA) colA_fill_tables<LGM>(a, tab, fM, 0, kRingGroupThreads, kRingGroupThreads);
	else {
		RingArgs ra;
		ra.tw = a.twN; ra.om = a.omN; ra.sig = a.sigN;
		ring_fill_tables<LGM + 4>(ra, tab, fN, 0, kRingGroupThreads, kRingGroupThreads);
	}
	const int total = a.ntiles * (SUBA ? 16 : G::NBLK);
	for (int cta = 0; cta < ncta; cta++)
		for (int gi = cta; gi < total; gi += ncta) {
			if (SUBA) {
				colA_load<LGM>(a, buf, gi, (void *)nullptr);
				colA_iter<LGM>(a, fM, buf, 0, 0, kRingGroupThreads);
				colA_store<LGM>(a, buf, gi);
			} else {
				colB_load<LGM>(a, buf, gi, (void *)nullptr);
				colB_iter<LGM>(a, fN, buf, gi / a.ntiles, 0, 0, kRingGroupThreads);
				colB_store<LGM>(a, buf, gi);
			}
		}
}
#endif

}  // namespace dsp
