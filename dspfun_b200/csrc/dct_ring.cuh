// dct_ring.cuh -- contiguous-axis (row) pass as a persistent, TMA-fed kernel: planar float lines, n = 2^LG, 256 <= n <= 8192.
//
// The one-shot row kernel (dct_fast.cuh: cta_row_fast) loads its lines with LDG -> registers -> STS and only then
// starts the transform: ncu (profiles/r01_ncu_current_summary.md) shows that load phase waiting on the long scoreboard
// with 16 warps per SM.  Here the lines arrive by bulk async copy (cp.async.bulk, the 1-D form of TMA: a line is one
// contiguous run) while the previous lines are being transformed:
//
//   * one CTA per SM, 2 independent GROUPS of 256 + 32 threads (8 warps that share every phase, 1 warp for the special
//     butterflies of the outer pass), each group with its own named barrier;
//   * a ring of 3 buffers of NSEQ = 8192 / n line pairs (8192 complex points): 2 being transformed, 1 in flight.  A buffer
//     first holds the raw lines exactly as the copy engine wrote them, then -- once the group has pulled them into
//     registers -- the padded complex sequences of the FFT passes (same Pad<> skew as the one-shot kernels);
//   * the group that finishes iteration `it` refills its buffer with the lines of iteration it + 3 (mbarrier
//     expect_tx + bulk copies issued by one thread) and moves on to iteration it + 2;
//   * the twiddles of the radix-16 passes and the half-sample phases sit in shared memory (the ring leaves little L1).
//
// DCT-II : raw lines -> [registers] first radix-r0 DIT pass (Makhoul permutation and digit reversal absorbed in the
//          addressing) -> padded slots -> middle radix-16 pass -> outer radix-16 pass + (k, n-k) twiddle -> STG.
// DCT-III: raw lines -> [registers] outer pass (reads the (k, n-k) pairs from the raw lines) -> padded slots -> middle
//          pass -> last radix-r0 DIF pass -> un-permuting STG.
//
// Same emulation discipline as the other kernels: phases separated by a (group) barrier, threads of a phase touch only
// their own slots.  Values that live in registers across a barrier are per-thread arrays on the GPU and a per-thread
// table in the emulation.
#pragma once
#include "dct_fast.cuh"
#include "dct_tma.cuh"

namespace dsp {

static const int kRingGroup = 256;       // general threads per group: every phase's work divides over them
static const int kRingSpecial = 0;       // 32: one more warp per group that only runs the outer pass's special butterflies
                                         // (i = 0, M/2), which make their general warp take twice as long as the others (ncu r02:
                                         // 12% of the samples wait at the barrier that ends an iteration).  Measured dead end:
                                         // 18 warps cap the kernel at 96 registers (warp allocation granularity 4) and it spills
                                         // 500-800 bytes per thread; with 0 the first thread of each sequence runs them.
static const int kRingGroupThreads = kRingGroup + kRingSpecial;
static const int kRingGroups = 2;
static const int kRingBufs = 3;
static const int kRingPoints = 8192;     // complex points per buffer

#if DSP_GPU
DSP_DEV void group_sync(int group) { asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(kRingGroupThreads) : "memory"); }
#define RING_SYNC(g) group_sync(g)
#else
#define RING_SYNC(g) ((void)0)
#endif

// ------------------------------------------------------------------------------------------------ descriptor
// FastFixed with the twiddles in shared memory:
//   s_out[4][M/2 + 1]  W_n^{i}, W_n^{2i}, W_n^{4i}, W_n^{8i}, i <= M/2           (outer pass: butterflies i and M - i)
//   s_mid[4][R0]       W_L^{i}, ...^{2i}, ^{4i}, ^{8i}, L = 16 R0, i < R0          (the one middle pass of n >= 1024)
//   s_om[M/2 + 1]      (cos, sin)(pi i / 2n)
//   s_sig[B]           slot table of the first / last pass (uint16)
template <int LG> struct RingFixed : FastFixedBase<LG> {
	const C2<float> *s_out, *s_mid, *s_om;
	const uint16_t *s_sig;                  // [B] padded slot of e = r (first / last pass butterfly r), B = n / R0
	enum { kHalf = (1 << (LG - 4)) / 2 + 1 };
	template <class T> DSP_DEVM void tw_from(const C2<float> *t, int stride, int i, C2<T> *w) const {
		w[1] = t[i]; w[2] = t[stride + i]; w[4] = t[2 * stride + i]; w[8] = t[3 * stride + i];
		w[3] = cmul(w[1], w[2]); w[5] = cmul(w[1], w[4]); w[6] = cmul(w[2], w[4]); w[7] = cmul(w[3], w[4]);
		w[9] = cmul(w[1], w[8]); w[10] = cmul(w[2], w[8]); w[11] = cmul(w[3], w[8]); w[12] = cmul(w[4], w[8]);
		w[13] = cmul(w[5], w[8]); w[14] = cmul(w[6], w[8]); w[15] = cmul(w[7], w[8]);
	}
	template <class T> DSP_DEVM void tw_mid(int i, int, C2<T> *w) const { tw_from<T>(s_mid, this->R0(), i, w); }
	template <class T> DSP_DEVM void tw_outer(int i, C2<T> *w) const { tw_from<T>(s_out, kHalf, i, w); }
	template <class T> DSP_DEVM C2<T> om_at(int i) const { return s_om[i]; }
	static constexpr int kTableElems = 4 * kHalf + 4 * (1 << FastFixedBase<LG>::kL0) + kHalf + ((1 << LG) >> FastFixedBase<LG>::kL0) / 4 + 1;
};

struct RingArgs {
	const float *in;
	float *out;
	long long ls_in, ls_out;     // line strides (elements); lines are contiguous runs of n floats
	int nlines;                  // even
	float lscale, sscale;        // plain multiplies on load / store (dsp_dct_fuse_scale)
	const void *tw, *om;         // global tables (C2<float>[n], C2<float>[n/2+1])
	const uint16_t *sig;
};

// ------------------------------------------------------------------------------------------------ storage adaptors
// the two raw lines of a pair as the DCT-III outer pass's source: element k of line A / B
struct StageRows {
	const float *ra, *rb;
	float m;
	DSP_DEVM C2<float> get(int k) const { return C2<float>{ra[k] * m, rb[k] * m}; }
};
// butterflies of an outer-pass unit kept in registers across the barrier that frees the raw lines
struct RegBf {
	C2<float> *a, *b;
	struct Row {
		C2<float> *p;
		DSP_DEVM void put(int j, C2<float> v) const { p[j] = v; }
		DSP_DEVM C2<float> get(int j) const { return p[j]; }
	};
	DSP_DEVM Row row(int) const { return Row{a}; }
	DSP_DEVM Row rowb(int) const { return Row{b}; }
};

template <int LG> struct RingGeom {
	typedef RingFixed<LG> F;
	enum {
		N = 1 << LG, M = N / 16, R0 = 1 << F::kL0, B = N / R0,
		NSEQ = kRingPoints / N,                         // line pairs per buffer
		ROUNDS = (NSEQ * B) / kRingGroup,               // first / last pass butterflies per thread (= 32 / R0)
		UNITS = M / 2,                                  // outer-pass units per sequence: unit 0 = butterflies 0 and M/2, unit u = (u, M - u)
		UROUNDS = (NSEQ * UNITS) / kRingGroup,          // = 1
	};
	static_assert(ROUNDS * R0 == 32 && UROUNDS == 1, "ring geometry");
};

#if DSP_GPU
#define RING_REGS(name, tid) name
#else
#define RING_REGS(name, tid) name##_all[tid]
#endif

// ------------------------------------------------------------------------------------------------ DCT-II iteration
// buf: raw lines [NSEQ][2][n] floats on entry; line0: first global line of the iteration; npairs: valid pairs (<= NSEQ)
template <int LG>
DSP_DEV void ring_fwd_iter(const RingArgs &a, const RingFixed<LG> &f, C2<float> *buf, long long line0, int npairs, int group,
                           int t0, int t1) {
	typedef RingGeom<LG> G;
	const int n = G::N, R0 = G::R0, B = G::B, NPAD = f.NPAD();
	const float *raw = (const float *)buf;
#if DSP_GPU
	C2<float> v[G::ROUNDS][R0];
#else
	static thread_local C2<float> v_all[kRingGroupThreads][G::ROUNDS][R0];
#endif
	// ---- first radix-R0 DIT pass straight from the raw lines: butterfly r of a sequence combines v[r + j B], where
	//      v[e] = x[2e] (e < n/2) | x[2(n-1-e) + 1]: even samples ascending, odd samples descending
	for (int tid = t0; tid < t1 && tid < kRingGroup; tid++) {
#pragma unroll
		for (int rd = 0; rd < G::ROUNDS; rd++) {
			const int g = tid + rd * kRingGroup, seq = g / B, r = g - seq * B;
			const float *ra = raw + seq * 2 * n, *rb = ra + n;
			C2<float> *vv = RING_REGS(v, tid)[rd];
#pragma unroll
			for (int j = 0; j < R0 / 2; j++) {
				const int x = 2 * (r + j * B);
				vv[j] = C2<float>{ra[x] * a.lscale, rb[x] * a.lscale};
			}
#pragma unroll
			for (int j = R0 / 2; j < R0; j++) {
				const int x = 2 * ((B - 1 - r) + (R0 - 1 - j) * B) + 1;
				vv[j] = C2<float>{ra[x] * a.lscale, rb[x] * a.lscale};
			}
			Dft<float, R0>::run(vv);
		}
	}
	RING_SYNC(group);                                            // every raw sample is in a register: the buffer is free
	for (int tid = t0; tid < t1 && tid < kRingGroup; tid++) {
#pragma unroll
		for (int rd = 0; rd < G::ROUNDS; rd++) {
			const int g = tid + rd * kRingGroup, seq = g / B, r = g - seq * B;
			C2<float> *p = buf + seq * NPAD + (int)f.s_sig[r];          // slot of e = r: Pad(rev(r) R0); element m follows at + Pad(m)
			const C2<float> *vv = RING_REGS(v, tid)[rd];
#pragma unroll
			for (int m = 0; m < R0; m++) p[Pad<float>::of(m)] = vv[m];
		}
	}
	RING_SYNC(group);
	for (int q = 0; q < f.NMID(); q++) {
		for (int tid = t0; tid < t1 && tid < kRingGroup; tid++) mid_pass<float, true>(buf, G::NSEQ, f, q, tid, kRingGroup);
		RING_SYNC(group);
	}
	// ---- outer pass + (k, n-k) twiddle, results straight to global memory (128 B per warp instruction)
	const OpMul<float> sop = {a.sscale};
	for (int tid = t0; tid < t1; tid++) {
		GlobalRows<float, OpMul<float>, 1> sink;
		sink.ca = Coord{0, 0, 0, 0, 0}; sink.cb = sink.ca;
		sink.d = 1; sink.ax_slot = 2; sink.op = &sop;
		if (tid < kRingGroup) {
			const int seq = tid / G::UNITS, u = tid - seq * G::UNITS;
			if (seq < npairs) {                                      // unit u = butterflies u and M - u
				sink.pa = a.out + (line0 + 2 * seq) * a.ls_out; sink.pb = sink.pa + a.ls_out;
				const SmemBf<float, RingFixed<LG>> bf{buf + seq * NPAD, &f};
				if (u != 0) dct2_outer_unit<float>(bf, f, u, sink);
				else if (kRingSpecial == 0) {
					dct2_outer_unit<float>(bf, f, 0, sink);
					dct2_outer_unit<float>(bf, f, G::M / 2, sink);
				}
			}
		} else if (kRingSpecial != 0) {
			for (int sl = tid - kRingGroup; sl < 2 * G::NSEQ; sl += kRingSpecial) {
				const int seq = sl >> 1;
				if (seq < npairs) {
					sink.pa = a.out + (line0 + 2 * seq) * a.ls_out; sink.pb = sink.pa + a.ls_out;
					dct2_outer_unit<float>(SmemBf<float, RingFixed<LG>>{buf + seq * NPAD, &f}, f, (sl & 1) ? G::M / 2 : 0, sink);
				}
			}
		}
	}
	RING_PROXY_FENCE();
	RING_SYNC(group);                                            // all slots read: the buffer may be refilled
}

// ------------------------------------------------------------------------------------------------ DCT-III iteration
template <int LG>
DSP_DEV void ring_inv_iter(const RingArgs &a, const RingFixed<LG> &f, C2<float> *buf, long long line0, int npairs, int group,
                           int t0, int t1) {
	typedef RingGeom<LG> G;
	const int n = G::N, R0 = G::R0, B = G::B, NPAD = f.NPAD(), M = G::M;
	const float *raw = (const float *)buf;
#if DSP_GPU
	C2<float> va[16], vb[16];
#else
	static thread_local C2<float> va_all[kRingGroupThreads][16], vb_all[kRingGroupThreads][16];
#endif
	// ---- outer pass: (k, n-k) pairs of both lines from the raw buffer -> pre-twiddle -> radix-16 -> registers
	//      general thread: unit u = butterflies u (-> va) and M - u (-> vb); special warp: butterflies 0 and M/2 of the
	//      sequences, lane sl = 2 seq + which, up to two per lane (-> va, then vb)
	for (int tid = t0; tid < t1; tid++) {
		StageRows src;
		src.m = a.lscale;
		if (tid < kRingGroup) {
			const int seq = tid / G::UNITS, u = tid - seq * G::UNITS;
			src.ra = raw + seq * 2 * n; src.rb = src.ra + n;
			if (u != 0) dct3_outer_unit<float>(RegBf{RING_REGS(va, tid), RING_REGS(vb, tid)}, f, u, src);
			else if (kRingSpecial == 0) {
				dct3_outer_unit<float>(RegBf{RING_REGS(va, tid), RING_REGS(va, tid)}, f, 0, src);
				dct3_outer_unit<float>(RegBf{RING_REGS(vb, tid), RING_REGS(vb, tid)}, f, M / 2, src);
			}
		} else {
			const int sl = tid - kRingGroup;
			if (sl < 2 * G::NSEQ) {
				src.ra = raw + (sl >> 1) * 2 * n; src.rb = src.ra + n;
				dct3_outer_unit<float>(RegBf{RING_REGS(va, tid), RING_REGS(va, tid)}, f, (sl & 1) ? M / 2 : 0, src);
			}
			if (sl + kRingSpecial < 2 * G::NSEQ) {
				const int s2 = sl + kRingSpecial;
				src.ra = raw + (s2 >> 1) * 2 * n; src.rb = src.ra + n;
				dct3_outer_unit<float>(RegBf{RING_REGS(vb, tid), RING_REGS(vb, tid)}, f, (s2 & 1) ? M / 2 : 0, src);
			}
		}
	}
	RING_SYNC(group);                                            // every raw sample is in a register: the buffer is free
	for (int tid = t0; tid < t1; tid++) {
		const C2<float> *xa = RING_REGS(va, tid), *xb = RING_REGS(vb, tid);
		if (tid < kRingGroup) {
			const int seq = tid / G::UNITS, u = tid - seq * G::UNITS;
			if (u != 0 || kRingSpecial == 0) {
				C2<float> *s = buf + seq * NPAD;
				C2<float> *pa = s + Pad<float>::of(u), *pb = s + Pad<float>::of(u == 0 ? M / 2 : M - u);
#pragma unroll
				for (int j = 0; j < 16; j++) { pa[f.PO(f.NMID(), j)] = xa[j]; pb[f.PO(f.NMID(), j)] = xb[j]; }
			}
		} else {
			const int sl = tid - kRingGroup;
			if (sl < 2 * G::NSEQ) {
				C2<float> *pa = buf + (sl >> 1) * NPAD + Pad<float>::of((sl & 1) ? M / 2 : 0);
#pragma unroll
				for (int j = 0; j < 16; j++) pa[f.PO(f.NMID(), j)] = xa[j];
			}
			if (sl + kRingSpecial < 2 * G::NSEQ) {
				const int s2 = sl + kRingSpecial;
				C2<float> *pb = buf + (s2 >> 1) * NPAD + Pad<float>::of((s2 & 1) ? M / 2 : 0);
#pragma unroll
				for (int j = 0; j < 16; j++) pb[f.PO(f.NMID(), j)] = xb[j];
			}
		}
	}
	RING_SYNC(group);
	for (int q = f.NMID() - 1; q >= 0; q--) {
		for (int tid = t0; tid < t1 && tid < kRingGroup; tid++) mid_pass<float, false>(buf, G::NSEQ, f, q, tid, kRingGroup);
		RING_SYNC(group);
	}
	// ---- last radix-R0 DIF pass on contiguous slots; sample v[r + m B] = (re, -im) goes to x = 2e (e < n/2) | 2(n-1-e)+1
	for (int tid = t0; tid < t1 && tid < kRingGroup; tid++) {
#pragma unroll
		for (int rd = 0; rd < G::ROUNDS; rd++) {
			const int g = tid + rd * kRingGroup, seq = g / B, r = g - seq * B;
			const C2<float> *p = buf + seq * NPAD + (int)f.s_sig[r];
			C2<float> vv[R0];
#pragma unroll
			for (int j = 0; j < R0; j++) vv[j] = p[Pad<float>::of(j)];
			Dft<float, R0>::run(vv);
			if (seq < npairs) {
				float *qa = a.out + (line0 + 2 * seq) * a.ls_out, *qb = qa + a.ls_out;
#pragma unroll
				for (int m = 0; m < R0 / 2; m++) {
					const int x = 2 * (r + m * B);
					qa[x] = vv[m].x * a.sscale; qb[x] = -vv[m].y * a.sscale;
				}
#pragma unroll
				for (int m = R0 / 2; m < R0; m++) {
					const int x = 2 * ((B - 1 - r) + (R0 - 1 - m) * B) + 1;
					qa[x] = vv[m].x * a.sscale; qb[x] = -vv[m].y * a.sscale;
				}
			}
		}
	}
	RING_PROXY_FENCE();
	RING_SYNC(group);
}

// ------------------------------------------------------------------------------------------------ CTA body
// shared memory: [tables][3 buffers of NSEQ * NPAD complex][3 mbarriers]
template <int LG> struct RingSmem {
	typedef RingFixed<LG> F;
	static constexpr int kBufElems = RingGeom<LG>::NSEQ * F::kNPAD;
	static constexpr size_t kTablesBytes = ((size_t)F::kTableElems * sizeof(C2<float>) + 127) / 128 * 128;
	static constexpr size_t kBufBytes = ((size_t)kBufElems * sizeof(C2<float>) + 127) / 128 * 128;
	static constexpr int kBufStride = (int)(kBufBytes / sizeof(C2<float>));     // buffers start 128-byte aligned (bulk copies need 16)
	static constexpr size_t kTotal = kTablesBytes + kRingBufs * kBufBytes + 64;
};

template <int LG>
DSP_DEV void ring_fill_tables(const RingArgs &a, C2<float> *tab, RingFixed<LG> &f, int t0, int t1, int nthr) {
	typedef RingFixed<LG> F;
	const int half = F::kHalf, R0 = 1 << F::kL0, n = 1 << LG;
	const C2<float> *tw = (const C2<float> *)a.tw, *om = (const C2<float> *)a.om;
	C2<float> *s_out = tab, *s_mid = tab + 4 * half, *s_om = s_mid + 4 * R0;
	uint16_t *s_sig = (uint16_t *)(s_om + half);
	for (int tid = t0; tid < t1; tid++) {
		for (int i = tid; i < n / R0; i += nthr) s_sig[i] = DSP_LDG(a.sig + i);
		for (int i = tid; i < half; i += nthr) {
#pragma unroll
			for (int p = 0; p < 4; p++) s_out[p * half + i] = ldg_c2(tw + (i << p));
			s_om[i] = ldg_c2(om + i);
		}
		// middle pass q = 0: Lprev = R0, L = 16 R0: W_L^{i 2^p} = W_n^{(i 2^p) n / L}
		for (int i = tid; i < R0; i += nthr) {
#pragma unroll
			for (int p = 0; p < 4; p++) s_mid[p * R0 + i] = ldg_c2(tw + (((i << p) * (n / (16 * R0))) & (n - 1)));
		}
	}
	f.tw = a.tw; f.om = a.om; f.sig = a.sig;
	f.s_out = s_out; f.s_mid = s_mid; f.s_om = s_om; f.s_sig = s_sig;
}

// lines of CTA-iteration `it`: global iteration gi = cta + it * ncta, lines [gi * 2 NSEQ, ...)
template <int LG> DSP_DEV long long ring_line0(int cta, int ncta, int it) { return ((long long)cta + (long long)it * ncta) * (2 * RingGeom<LG>::NSEQ); }
template <int LG> DSP_DEV int ring_pairs(const RingArgs &a, long long line0) {
	const long long left = ((long long)a.nlines - line0) / 2;
	return left >= RingGeom<LG>::NSEQ ? RingGeom<LG>::NSEQ : (left > 0 ? (int)left : 0);
}

#if DSP_GPU
template <int LG>
DSP_DEV void ring_issue(const RingArgs &a, float *dst, long long line0, int npairs, uint64_t *bar) {
	const uint32_t lbytes = (uint32_t)(sizeof(float) << LG);
	mbar_expect_tx(bar, 2u * (uint32_t)npairs * lbytes);
	const char *src = (const char *)(a.in + line0 * a.ls_in);
	if (a.ls_in == (1 << LG)) bulk_g2s(dst, src, 2u * (uint32_t)npairs * lbytes, bar);        // the lines are one run
	else
		for (int l = 0; l < 2 * npairs; l++) bulk_g2s((char *)dst + (size_t)l * lbytes, src + (size_t)l * a.ls_in * sizeof(float), lbytes, bar);
}

template <int LG, bool FWD>
DSP_DEV void ring_cta(const RingArgs &a, unsigned char *smem, int cta, int ncta, int tid) {
	typedef RingSmem<LG> S;
	typedef RingGeom<LG> G;
	C2<float> *tab = (C2<float> *)smem;
	C2<float> *bufs = (C2<float> *)(smem + S::kTablesBytes);
	uint64_t *full = (uint64_t *)(smem + S::kTablesBytes + kRingBufs * S::kBufBytes);
	RingFixed<LG> f;
	ring_fill_tables<LG>(a, tab, f, tid, tid + 1, kRingGroups * kRingGroupThreads);
	const long long total_iters = ((long long)a.nlines / 2 + G::NSEQ - 1) / G::NSEQ;
	const int iters = (int)((total_iters - cta + ncta - 1) / ncta);              // iterations of this CTA
	if (tid == 0) {
		for (int b = 0; b < kRingBufs; b++) mbar_init(full + b, 1);
		mbar_fence_init();
	}
	__syncthreads();
	if (tid == 0) {
		for (int it = 0; it < kRingBufs && it < iters; it++) {
			const long long l0 = ring_line0<LG>(cta, ncta, it);
			ring_issue<LG>(a, (float *)(bufs + (size_t)it * S::kBufStride), l0, ring_pairs<LG>(a, l0), full + it);
		}
	}
	const int group = tid / kRingGroupThreads, gt = tid - group * kRingGroupThreads;
	for (int it = group; it < iters; it += kRingGroups) {
		const int b = it % kRingBufs;
		C2<float> *buf = bufs + (size_t)b * S::kBufStride;
		const long long l0 = ring_line0<LG>(cta, ncta, it);
		const int np = ring_pairs<LG>(a, l0);
		mbar_wait(full + b, (uint32_t)((it / kRingBufs) & 1));
		if (FWD) ring_fwd_iter<LG>(a, f, buf, l0, np, group, gt, gt + 1);
		else ring_inv_iter<LG>(a, f, buf, l0, np, group, gt, gt + 1);
		// the iteration ended on a group barrier: the buffer is free.  Refill it for iteration it + 3.
		if (gt == 0 && it + kRingBufs < iters) {
			fence_proxy_async();
			const long long l3 = ring_line0<LG>(cta, ncta, it + kRingBufs);
			ring_issue<LG>(a, (float *)buf, l3, ring_pairs<LG>(a, l3), full + b);
		}
	}
}

template <int LG, bool FWD>
__global__ void __launch_bounds__(kRingGroups *kRingGroupThreads, 1) k_row_ring(const __grid_constant__ RingArgs a) {
	extern __shared__ __align__(128) unsigned char ring_smem[];
	ring_cta<LG, FWD>(a, ring_smem, (int)blockIdx.x, (int)gridDim.x, (int)threadIdx.x);
}
#else
// emulation: CTAs, iterations and threads one after another; the bulk copy is a memcpy
template <int LG, bool FWD>
static void ring_emulate(const RingArgs &a, int ncta) {
	typedef RingSmem<LG> S;
	typedef RingGeom<LG> G;
	std::vector<unsigned char> smem(S::kTotal + 128);
	C2<float> *tab = (C2<float> *)smem.data();
	C2<float> *buf = (C2<float> *)(smem.data() + S::kTablesBytes);
	RingFixed<LG> f;
	ring_fill_tables<LG>(a, tab, f, 0, kRingGroupThreads, kRingGroupThreads);
	const long long total_iters = ((long long)a.nlines / 2 + G::NSEQ - 1) / G::NSEQ;
	for (int cta = 0; cta < ncta; cta++) {
		const int iters = (int)((total_iters - cta + ncta - 1) / ncta);
		for (int it = 0; it < iters; it++) {
			const long long l0 = ring_line0<LG>(cta, ncta, it);
			const int np = ring_pairs<LG>(a, l0);
			for (int l = 0; l < 2 * np; l++) memcpy((float *)buf + (size_t)l * G::N, a.in + (l0 + l) * a.ls_in, sizeof(float) * G::N);
			if (FWD) ring_fwd_iter<LG>(a, f, buf, l0, np, it & 1, 0, kRingGroupThreads);
			else ring_inv_iter<LG>(a, f, buf, l0, np, it & 1, 0, kRingGroupThreads);
		}
	}
}
#endif

}  // namespace dsp
