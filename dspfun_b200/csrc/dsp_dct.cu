// dsp_dct.cu -- planner, table cache, kernel launches and the C ABI of libdspdct (include/dsp_dct.h).
//
// A plan is a list of passes, one per transformed axis.  The pass over the contiguous axis uses the row
// kernel (a line of n*d elements per sequence group), every other axis uses the column kernel (a tile of
// adjacent columns over the whole axis).  Forward plans run the contiguous axis first, inverse plans run it
// last, so a forward+inverse round trip meets in the middle on the same column tiles (L2 reuse).
#include "../../include/dsp_dct.h"
#include "dsp_kernels.h"
#include "dct_split.cuh"
#include "dct_ring.cuh"
#include "dct_colring.cuh"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#if !DSP_GPU
#include <thread>
#else
#include <nvtx3/nvToolsExt.h>          // header-only NVTX 3: ranges cost nothing unless a tool is attached
#endif

// The power-of-two fast path is built for float and double (the Makefile passes -DDSP_FAST_F64=1 and adds the
// kern_*_fast_f64.cu / kern_split_f64.cu units); with DSP_FAST_F64=0 double runs on the generic mixed-radix engine.
#ifndef DSP_FAST_F64
#define DSP_FAST_F64 0
#endif

namespace dsp {

static thread_local std::string g_err;
static std::mutex g_mu;
static std::atomic<unsigned long long> g_launches(0);

// DSP_DCT_NVTX=1 wraps every pass launch in an NVTX range ("dct pass <i> <row|col|split> n=<len>"), so that the passes
// of a plan show up by name on an Nsight Systems / Nsight Compute timeline (tracing)
static bool nvtx_on() { static int t = -1; if (t < 0) { const char *e = getenv("DSP_DCT_NVTX"); t = (e && *e && *e != '0') ? 1 : 0; } return t == 1; }
// DSP_DCT_TRACE=1 prints planner / launch steps to stderr
static bool trace_on() { static int t = -1; if (t < 0) { const char *e = getenv("DSP_DCT_TRACE"); t = (e && *e && *e != '0') ? 1 : 0; } return t == 1; }
#define DSP_TRACE(...) do { if (trace_on()) { fprintf(stderr, "[dsp_dct] " __VA_ARGS__); fprintf(stderr, "\n"); fflush(stderr); } } while (0)


// ------------------------------------------------------------------------------------------------ tables
static FastDiv mk_fd(uint32_t d) {
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

// radices for the DIF passes: odd primes first (largest sub-stride), powers of two last with 16s at the end,
// which keeps every pass's smem access pattern a power-of-two stride under the bank-skew padding.
static bool factorize(int n, std::vector<int> &fac) {
	fac.clear();
	static const int odd[] = {13, 11, 7};
	for (int r : odd)
		while (n % r == 0) { fac.push_back(r); n /= r; }
	// 3s and 5s pair up into the composite radices 15 and 9 (one shared-memory pass instead of two)
	int c3 = 0, c5 = 0;
	while (n % 5 == 0) { c5++; n /= 5; }
	while (n % 3 == 0) { c3++; n /= 3; }
	const bool merge = !getenv("DSP_DCT_NO_MERGE");
	while (merge && c5 > 0 && c3 > 0) { fac.push_back(15); c5--; c3--; }
	while (merge && c3 >= 2) { fac.push_back(9); c3 -= 2; }
	while (c5-- > 0) fac.push_back(5);
	while (c3-- > 0) fac.push_back(3);
	int e = 0;
	while (n % 2 == 0) { e++; n /= 2; }
	if (n != 1) return false;
	if (e % 4) fac.push_back(1 << (e % 4));
	for (int i = 0; i < e / 4; i++) fac.push_back(16);
	return (int)fac.size() <= DSP_MAX_FAC;
}

struct Tables {
	int n;
	char prec;
	std::vector<int> fac;
	int npad;
	void *tw, *om;
	uint16_t *pos2, *pos3;
	// power-of-two fast path (n = r0 * 16^(nmid+1)); sig == nullptr when not applicable
	int r0, nmid;
	uint16_t *sig;
	// dense fallback
	bool dense;
	int half;
	void *ctab;
};
static std::map<std::pair<int, std::pair<int, char>>, Tables *> g_tables;

template <class T> static int pad_of(int e) { return Pad<T>::of(e); }

template <class T> static Tables *build_tables(int n) {
	Tables *t = new Tables();
	t->n = n;
	t->prec = sizeof(T) == 4 ? 'f' : 'd';
	t->tw = t->om = nullptr;
	t->pos2 = t->pos3 = nullptr;
	t->dense = false; t->half = 0; t->ctab = nullptr; t->sig = nullptr;
	t->npad = pad_of<T>(n - 1) + 1;
	// sequence stride == 1 (mod 16 complex floats / 8 complex doubles): the column kernels write the same slot of
	// neighbouring sequences from neighbouring lanes, and an even multiple of 128 B would put them all in one bank
	{
		const int m = sizeof(T) == 4 ? 16 : 8;
		while (t->npad % m != 1) t->npad++;
	}
	if (!factorize(n, t->fac)) {
		// a prime factor > 13: direct evaluation of the definition (O(n^2) per line, exact same results contract)
		t->dense = true;
		t->fac.clear();
		t->half = t->npad;
		t->npad = 2 * t->half;
	}
	const long double pi = 3.141592653589793238462643383279502884L;
	std::vector<C2<T>> tw(n), om(n / 2 + 1);
	for (int k = 0; k < n; k++) {
		const long double a = 2 * pi * (long double)k / (long double)n;
		tw[k].x = (T)cosl(a);
		tw[k].y = (T)(-sinl(a));
	}
	for (int k = 0; k <= n / 2; k++) {
		const long double a = pi * (long double)k / (2 * (long double)n);
		om[k].x = (T)cosl(a);
		om[k].y = (T)sinl(a);
	}
	std::vector<uint16_t> pos2(n), pos3(n);
	for (int k = 0; k < n; k++) {
		int kk = k, P = 0, stride = n;
		for (int r : t->fac) { stride /= r; P += (kk % r) * stride; kk /= r; }
		pos2[k] = (uint16_t)(t->dense ? k : P);
	}
	for (int j = 0; j < n; j++) pos3[j] = t->dense ? (uint16_t)j : pos2[(j & 1) ? n - 1 - (j >> 1) : (j >> 1)];
	std::vector<T> ctab;
	if (t->dense) {
		ctab.resize(4 * (size_t)n);
		for (long m = 0; m < 4L * n; m++) ctab[m] = (T)cosl(pi * (long double)m / (2 * (long double)n));
	}
	// fast path: n = r0 * 16^(nmid+1), r0 in {1,2,4,8,16,32}
	t->r0 = 0; t->nmid = 0;
	std::vector<uint16_t> sig;
	if (!t->dense && n >= 16 && (n & (n - 1)) == 0 && !getenv("DSP_DCT_NO_FAST") && (sizeof(T) == 4 || DSP_FAST_F64)) {
		int a = 0;
		while ((1 << a) < n) a++;
		int l0 = (a - 4) % 4;                       // log2 r0 in {0,1,2,3} ...
		int k = (a - 4 - l0) / 4;                   // ... middle passes
		if (l0 == 0 && k > 0) { l0 = 4; k--; }      // prefer r0 = 16 over r0 = 1 with one more middle pass
		else if (l0 == 1 && k > 0) { l0 = 5; k--; } // and r0 = 32 over r0 = 2
		if (k <= 3) {
			t->r0 = 1 << l0; t->nmid = k;
			sig.resize(n);
			for (int e = 0; e < n; e++) {
				int x = e, stride = n, slot = 0;
				for (int p = 0; p <= k; p++) { stride /= 16; slot += (x % 16) * stride; x /= 16; }
				slot += x;
				sig[e] = (uint16_t)pad_of<T>(slot);
			}
		}
	}
	std::string err;
	DSP_TRACE("tables: host side done (nfac=%d npad=%d), allocating", (int)t->fac.size(), t->npad);
	bool ok = rt_malloc(&t->tw, sizeof(C2<T>) * n, err) && rt_malloc(&t->om, sizeof(C2<T>) * (n / 2 + 1), err) &&
	          rt_malloc((void **)&t->pos2, sizeof(uint16_t) * n, err) && rt_malloc((void **)&t->pos3, sizeof(uint16_t) * n, err) &&
	          rt_h2d(t->tw, tw.data(), sizeof(C2<T>) * n, 0, err) && rt_h2d(t->om, om.data(), sizeof(C2<T>) * (n / 2 + 1), 0, err) &&
	          rt_h2d(t->pos2, pos2.data(), sizeof(uint16_t) * n, 0, err) && rt_h2d(t->pos3, pos3.data(), sizeof(uint16_t) * n, 0, err) &&
	          rt_sync(0, err);
	if (ok && t->dense)
		ok = rt_malloc(&t->ctab, sizeof(T) * ctab.size(), err) && rt_h2d(t->ctab, ctab.data(), sizeof(T) * ctab.size(), 0, err) && rt_sync(0, err);
	if (ok && !sig.empty())
		ok = rt_malloc((void **)&t->sig, sizeof(uint16_t) * n, err) && rt_h2d(t->sig, sig.data(), sizeof(uint16_t) * n, 0, err) && rt_sync(0, err);
	if (!ok) {
		g_err = err;
		rt_free(t->tw); rt_free(t->om); rt_free(t->pos2); rt_free(t->pos3); rt_free(t->sig); rt_free(t->ctab);
		delete t;
		return nullptr;
	}
	return t;
}

static std::mutex g_tab_mu;                // the table cache is shared by plan calls and by execute_host's lazily built chunk plans
static Tables *get_tables(int n, char prec) {
	std::lock_guard<std::mutex> tab_lock(g_tab_mu);
	if (n > 65535) { g_err = "transform length " + std::to_string(n) + " too large"; return nullptr; }
	auto key = std::make_pair(rt_device(), std::make_pair(n, prec));
	auto it = g_tables.find(key);
	if (it != g_tables.end()) return it->second;
	DSP_TRACE("building tables n=%d prec=%c", n, prec);
	Tables *t = prec == 'f' ? build_tables<float>(n) : build_tables<double>(n);
	DSP_TRACE("tables %s", t ? "ok" : g_err.c_str());
	if (t) g_tables[key] = t;
	return t;
}

static void fill_fft(FftDesc &f, const Tables *t) {
	memset(&f, 0, sizeof(f));
	f.n = t->n;
	f.nfac = (int)t->fac.size();
	int L = t->n;
	for (int p = 0; p < f.nfac; p++) {
		f.fac[p] = t->fac[p];
		f.dM[p] = mk_fd((uint32_t)(L / f.fac[p]));
		f.dNb[p] = mk_fd((uint32_t)(t->n / f.fac[p]));
		L /= f.fac[p];
	}
	f.npad = t->npad;
	f.tw = t->tw; f.om = t->om; f.pos2 = t->pos2; f.pos3 = t->pos3;
	f.dHalf = mk_fd((uint32_t)(t->n / 2 + 1));
	f.dense = t->dense ? 1 : 0; f.half = t->half; f.ctab = t->ctab;
	f.dN = mk_fd((uint32_t)t->n);
}

template <class T> static void fill_fast_t(FastDesc &f, const Tables *t) {
	memset(&f, 0, sizeof(f));
	f.n = t->n; f.M = t->n / 16;
	f.r0 = t->r0; f.nmid = t->nmid;
	f.npad = t->npad;
	f.tw = t->tw; f.om = t->om; f.sig = t->sig;
	int lp = t->r0;
	for (int q = 0; q <= t->nmid; q++) {
		for (int j = 0; j < 16; j++) f.poff[q][j] = pad_of<T>(j * lp);
		lp *= 16;
	}
	f.dHalf = mk_fd((uint32_t)(f.M / 2 + 1));
}
static void fill_fast(FastDesc &f, const Tables *t) {
	if (t->prec == 'f') fill_fast_t<float>(f, t); else fill_fast_t<double>(f, t);
}

// ------------------------------------------------------------------------------------------------ plan
struct Level { long long cnt, is, os; int slot; };

struct PassPlan {
	bool row;
	int axis;
	bool fast;
	FastDesc ff;
	RowArgs ra;
	ColArgs ca;
	int grid, block;
	size_t smem;
	int pf_dist;                             // CTAs ahead whose input gets prefetched into L2 (0 = off)
	bool vec_in_layout, vec_out_layout;
	OpAny lop, sop;
	bool fused;
	// split column pass (dct_split.cuh): n = 16 M in two L2-resident sub-passes per panel of sp_P columns
	bool split;
	FastDesc ffM;
	int sp_P, sp_tc;
	size_t sp_smem, sp_smem_inv;
	bool sp_force_inv;                       // DSP_DCT_SPLIT_MIN given: use the DIF-style split inverse (experiments)
	// segmented output of the last pass (dsp_dct_set_output_segments): absolute device bases, one per segment
	int seg_n;
	void *seg_base[8];
	// outermost loop level of the pass (chunked, L2-resident schedule): consecutive passes that share it run chunk by
	// chunk, so the intermediate of a chunk is still in L2 when the next pass reads it
	int ch_slot;                             // coordinate slot of that level (-1: the pass has no outer level)
	long long ch_cnt, ch_is, ch_os, ch_inner;// its count and strides (elements), product of the levels below it
	// ring sub-pass kernels (dct_colring.cuh): tensor maps per (input, output, scratch) pointer triple
	struct RingMaps { const void *in; void *out, *scratch; int nplanes; ColRingArgs args; };
	std::vector<RingMaps> ring_maps;
	int rg_P;                                // ring sub-pass kernels: panel width (0 = not eligible)
	bool seg_saved;                          // strides below hold the plan's own values while an override is active
	long long seg_os[4], seg_ax_os;
};

struct MultiGpu;

}  // namespace dsp

using namespace dsp;

struct dsp_dct_plan_s {
	char prec;
	int es;                                  // element size
	int rank, n[3], kind[3];
	int d;                                   // interleave (1 = planar)
	int device;
	void *in, *out;                          // pointers captured at plan time
	size_t in_span, out_span;                // elements covered by the layouts
	bool out_has_gaps;
	std::vector<PassPlan> passes;
	// host-pointer staging
	void *d_in, *d_out;
	size_t d_in_bytes, d_out_bytes;
	// fusion state
	int fuse_kind;                           // 0 none, 1 spec, 2 ispec
	double *d_scalars;                       // acc[4] | scale_z[4] | dc_out[4]
	unsigned char *d_signmap;
	void *d_work;                            // T scratch the passes run in when the final store is 8-bit
	int *d_ring_done;                        // completion counters of the ring sub-pass kernels (2 x max panels)
	void *d_split;                           // panel scratch of the split column passes (nscratch parts: panels rotate)
	size_t split_bytes;
	int nscratch;
#if DSP_GPU
	cudaStream_t aux[4];                     // panels rotate over a few streams so that one panel's tail and launch
	cudaEvent_t ev_fork, ev_join[4];         // gaps are filled by the next panels' kernels
	bool aux_ok;
	int naux;
#endif
	bool need_acc;
	rt_stream last_stream;
	// creation parameters, kept so that the host-buffer entry can re-plan sub-batches (chunk pipeline, execute_host)
	int c_howmany, c_istride, c_idist, c_ostride, c_odist, c_nbatch;
	long long c_ibdist, c_obdist;
	int c_ie[3], c_oe[3];
	bool c_has_ie, c_has_oe;
	std::vector<dsp_dct_plan_s *> kids;      // one sub-plan per chunk of batch elements
	std::vector<long long> kid_b0;           // first batch element of each chunk
	bool kids_failed;
#if DSP_GPU
	cudaStream_t kid_st[4];
#endif
	// per-pass profiling
	bool profiling;
	double samples_per_launch;
#if DSP_GPU
	std::vector<cudaEvent_t> ev;             // 2 events per pass per recorded execute
#endif
	std::vector<int> ev_pass;
	std::vector<float> ev_frac;              // share of the pass each recorded launch covered (chunked schedule)
	dsp::MultiGpu *mg;                       // rank-3 host-buffer plans over several GPUs of this process (dsp_dct_plan_with_ngpus)
	bool mg_tried;
};

namespace dsp {

static bool set_outer(Outer &o, std::vector<Level> lv) {
	std::vector<Level> keep;
	for (auto &l : lv) if (l.cnt > 1) keep.push_back(l);
	if (keep.size() > 4) { g_err = "layout needs more than four outer loop levels"; return false; }
	for (int i = 0; i < 4; i++) {
		if (i < (int)keep.size()) { o.cnt[i] = (int)keep[i].cnt; o.is[i] = keep[i].is; o.os[i] = keep[i].os; o.slot[i] = keep[i].slot; }
		else { o.cnt[i] = 1; o.is[i] = 0; o.os[i] = 0; o.slot[i] = -1; }
	}
	o.d0 = mk_fd((uint32_t)o.cnt[0]);
	o.d01 = mk_fd((uint32_t)((long long)o.cnt[0] * o.cnt[1]));
	o.d012 = mk_fd((uint32_t)((long long)o.cnt[0] * o.cnt[1] * o.cnt[2]));
	o.base_slot = -1; o.base = 0;
	return true;
}

static void set_chunk_level(PassPlan &pp, const std::vector<Level> &lv) {
	pp.ch_slot = -1; pp.ch_cnt = 1; pp.ch_is = pp.ch_os = 0; pp.ch_inner = 1;
	long long inner = 1;
	for (auto &l : lv) {
		if (l.cnt <= 1) continue;
		pp.ch_inner = inner;
		pp.ch_slot = l.slot; pp.ch_cnt = l.cnt; pp.ch_is = l.is; pp.ch_os = l.os;
		inner *= l.cnt;
	}
}

static bool build_plan(dsp_dct_plan_s *P, int howmany, const int *inembed, int istride, int idist, const int *onembed,
                       int ostride, int odist, int nbatch, long long ibdist, long long obdist) {
	const int r = P->rank;
	const int VN = 16 / P->es;
	// ---- layout
	int d;
	if (istride == 1 && ostride == 1) d = 1;
	else if (istride == howmany && ostride == howmany && idist == 1 && odist == 1 && howmany > 1) d = howmany;
	else {
		g_err = "unsupported layout: need planar (stride 1) or channel-interleaved (stride == howmany, dist == 1) buffers";
		return false;
	}
	P->d = d;
	long long sI[3], sO[3];
	const int *ie = inembed ? inembed : P->n, *oe = onembed ? onembed : P->n;
	for (int i = 0; i < r; i++)
		if (ie[i] < P->n[i] || oe[i] < P->n[i]) {
			if (i > 0) { g_err = "embed smaller than the logical size"; return false; }
		}
	sI[r - 1] = d; sO[r - 1] = d;
	for (int i = r - 2; i >= 0; i--) { sI[i] = sI[i + 1] * ie[i + 1]; sO[i] = sO[i + 1] * oe[i + 1]; }
	std::vector<Level> batch;
	if (d == 1 && howmany > 1) batch.push_back(Level{howmany, idist, odist, 3});
	if (nbatch > 1) batch.push_back(Level{nbatch, ibdist, obdist, 4});
	// spans (elements) for host staging
	long long ispan = 1, ospan = 1, dense = (long long)d;
	for (int i = 0; i < r; i++) { ispan += (long long)(P->n[i] - 1) * sI[i]; ospan += (long long)(P->n[i] - 1) * sO[i]; dense *= P->n[i]; }
	ispan += d - 1; ospan += d - 1;
	for (auto &b : batch) { ispan += (b.cnt - 1) * b.is; ospan += (b.cnt - 1) * b.os; dense *= b.cnt; }
	P->in_span = (size_t)ispan; P->out_span = (size_t)ospan;
	P->out_has_gaps = ospan != dense;
	P->samples_per_launch = (double)dense;

	// ---- pass order
	std::vector<int> order;
	if (P->kind[0] == DSP_DCT_REDFT10) for (int a = r - 1; a >= 0; a--) order.push_back(a);
	else for (int a = 0; a < r; a++) order.push_back(a);

	const size_t cbytes = 2 * (size_t)P->es;
	for (size_t pi = 0; pi < order.size(); pi++) {
		const int ax = order[pi];
		const bool first = pi == 0;
		const long long *sIn = first ? sI : sO;
		Tables *t = get_tables(P->n[ax], P->prec);
		if (!t) return false;
		PassPlan pp;
		memset(&pp.ra, 0, sizeof(pp.ra)); memset(&pp.ca, 0, sizeof(pp.ca));
		memset(&pp.lop, 0, sizeof(pp.lop)); memset(&pp.sop, 0, sizeof(pp.sop));
		pp.fused = false;
		pp.axis = ax;
		// the contiguous-axis (row) kernel keeps all d interleaved sequences of a line on chip; a wide interleave
		// (e.g. a temporal transform over [D][H*W] with stride H*W) is a strided axis over d contiguous columns instead
		// ... and so is an interleaved line too long for all its channels to sit on chip together (e.g. 8192 RGB doubles)
		const bool wide = d > 4 || (d > 1 && (size_t)t->npad * 2 * (size_t)P->es * (size_t)d > kMaxSmem);
		pp.row = ax == r - 1 && !wide;
		pp.fast = t->sig != nullptr;
		pp.rg_P = 0;
		pp.split = false; pp.sp_P = 0; pp.sp_tc = 0; pp.sp_smem = 0; pp.sp_smem_inv = 0; pp.sp_force_inv = false;
		pp.seg_n = 0; pp.seg_saved = false;
		memset(&pp.ffM, 0, sizeof(pp.ffM));
		memset(&pp.ff, 0, sizeof(pp.ff));
		if (pp.fast) fill_fast(pp.ff, t);
		const size_t seqb = (size_t)t->npad * cbytes;
		std::vector<Level> lv;
		bool vin = true, vout = true;
		if (pp.row) {
			for (int a = r - 2; a >= 0; a--) lv.push_back(Level{P->n[a], sIn[a], sO[a], a + 3 - r});
			for (auto &b : batch) lv.push_back(Level{b.cnt, first ? b.is : b.os, b.os, b.slot});
			RowArgs &A = pp.ra;
			fill_fft(A.f, t);
			A.kind = P->kind[ax];
			A.d = d; A.dd = mk_fd((uint32_t)d);
			long long nl = 1;
			for (auto &l : lv) { nl *= l.cnt; if (l.cnt > 1 && (l.is % VN)) vin = false; if (l.cnt > 1 && (l.os % VN)) vout = false; }
			if (nl >= (1ll << 31)) { g_err = "too many lines"; return false; }
			A.nlines = (int)nl;
			if (!set_outer(A.o, lv)) return false;
			A.ax_slot = 2;
			{
				// do all outer levels collapse to a single line stride?
				std::vector<Level> act;
				for (auto &l : lv) if (l.cnt > 1) act.push_back(l);
				bool simple = true;
				for (size_t k = 1; k < act.size(); k++)
					if (act[k].is != act[k - 1].cnt * act[k - 1].is || act[k].os != act[k - 1].cnt * act[k - 1].os) simple = false;
				A.simple = simple ? 1 : 0;
				A.ls_in = act.empty() ? 0 : act[0].is;
				A.ls_out = act.empty() ? 0 : act[0].os;
			}
			const size_t pairb = seqb * (size_t)d;
			if (pairb > kMaxSmem) { g_err = "transform length " + std::to_string(P->n[ax]) + " does not fit on chip"; return false; }
			long long pairs = (long long)((48 * 1024) / pairb);
			if (pairs < 1) pairs = 1;
			const long long total_pairs = (nl + 1) / 2;
			// keep at least ~4 CTAs per SM worth of work when the problem allows it
			while (pairs > 1 && total_pairs / pairs < 4 * 148) pairs--;
			if (pairs > total_pairs) pairs = total_pairs;
			A.lines_per_cta = (int)(2 * pairs);
			pp.grid = (int)((nl + A.lines_per_cta - 1) / A.lines_per_cta);
			pp.smem = (size_t)pairs * pairb;
		} else {
			for (int a = r - 2; a >= 0; a--) if (a != ax) lv.push_back(Level{P->n[a], sIn[a], sO[a], a + 3 - r});
			for (auto &b : batch) lv.push_back(Level{b.cnt, first ? b.is : b.os, b.os, b.slot});
			ColArgs &A = pp.ca;
			fill_fft(A.f, t);
			A.kind = P->kind[ax];
			const bool lastax = ax == r - 1;                   // only with a wide interleave: the columns are the d channels
			A.ncols = lastax ? d : P->n[r - 1] * d;
			A.d = d; A.dd = mk_fd((uint32_t)d);

			A.ax_is = sIn[ax]; A.ax_os = sO[ax];
			if (A.ax_is % VN) vin = false;
			if (A.ax_os % VN) vout = false;
			long long no = 1;
			for (auto &l : lv) { no *= l.cnt; if (l.cnt > 1 && (l.is % VN)) vin = false; if (l.cnt > 1 && (l.os % VN)) vout = false; }
			if (!set_outer(A.o, lv)) return false;
			A.ax_slot = ax + 3 - r; A.col_slot = lastax ? 3 : 2;
			if (lastax) { A.d = 1; A.dd = mk_fd(1); }          // column index == channel index
			// columns per CTA: as wide as fits ~75 KB, i.e. three CTAs per SM (wider rows of the tile = longer
			// contiguous global segments; measured at n = 1080: 16 columns 3.48 ms vs 8 columns 3.93 ms per 256 frames)
			const int cand[] = {8 * VN, 4 * VN, 2 * VN, VN};
			int tc = 0;
			for (int c : cand) if ((size_t)(c / 2) * seqb <= 75 * 1024) { tc = c; break; }
			// short axes (the 8..64-point transforms of a block DCT): a 32-column tile would hold only a few hundred
			// samples per CTA; widen it to ~36 KB of sequences (measured at n = 8 over 256x1080x1920: 13.6 -> 2.2 ms per pass)
			if (P->n[ax] <= 64)
				for (int c = 128 * VN; c > tc; c /= 2)
					if ((size_t)(c / 2) * seqb <= 36 * 1024) { tc = c; break; }
			if (!tc) {
				tc = (VN >= 4 && 2 * seqb <= kMaxSmem) ? VN : 2;
				if ((size_t)(tc / 2) * seqb > kMaxSmem) { g_err = "transform length " + std::to_string(P->n[ax]) + " does not fit on chip"; return false; }
			}
			if (getenv("DSP_DCT_TC")) { const int want = atoi(getenv("DSP_DCT_TC")); int t2 = 2; while (t2 * 2 <= want && t2 < 64) t2 *= 2; if ((size_t)(t2 / 2) * seqb <= kMaxSmem) tc = t2; }  /* tuning aid: power of two */
			while (tc > VN && tc / 2 >= A.ncols) tc /= 2;
			// enough CTAs to fill the machine
			while (tc > 2 * VN && ((A.ncols + tc - 1) / tc) * no < 2 * 148) tc /= 2;
			A.tc = tc;
			A.ntiles = (A.ncols + tc - 1) / tc;
			A.dtiles = mk_fd((uint32_t)A.ntiles);
			const long long g = (long long)A.ntiles * no;
			if (g >= (1ll << 31)) { g_err = "too many column tiles"; return false; }
			pp.grid = (int)g;
			pp.smem = (size_t)((tc + 1) / 2) * seqb;
			// long power-of-two axes: two L2-resident sub-passes per column panel (dct_split.cuh)
			// Measured on B200 (profiles/): at n = 8192 the split takes 0.49 ms (forward) / 0.54 ms (inverse) per
			// 2 planes against 0.88 ms for the one-kernel column pass; at n = 1024 it loses (many small launches), so it
			// is on for n >= 4096.  DSP_DCT_SPLIT_MIN=<n> moves the threshold and selects the DIF-style inverse.
			int split_min = 4096;
			bool split_inv = false;
			if (getenv("DSP_DCT_SPLIT_MIN")) { split_min = atoi(getenv("DSP_DCT_SPLIT_MIN")); split_inv = true; }
			// double: the forward split only (measured at n = 8192: forward 0.99 -> 0.86 ms, while the DIF-style inverse --
			// the only element-type agnostic one -- loses to the one-kernel pass, 1.20 vs 0.94 ms)
			const bool f64_ok = DSP_FAST_F64 && !getenv("DSP_DCT_NO_SPLIT_F64") && (P->kind[ax] == DSP_DCT_REDFT10 || split_inv);
			const int nn = P->n[ax];
			if (pp.fast && (P->prec == 'f' || f64_ok) && nn >= split_min && nn >= 256 && vin && vout && !lastax &&
			    (split_inv || P->kind[ax] == DSP_DCT_REDFT10 || (A.ncols % 16) == 0)) {
				Tables *tM = get_tables(nn / 16, P->prec);
				if (tM && tM->sig) {
					pp.split = true;
					fill_fast(pp.ffM, tM);
					long long pw = (32ll << 20) / ((long long)nn * P->es);
					if (getenv("DSP_DCT_SPLIT_PANEL_MB")) pw = ((long long)atoi(getenv("DSP_DCT_SPLIT_PANEL_MB")) << 20) / ((long long)nn * P->es);
					pw = (pw / 64) * 64;
					if (pw < 64) pw = 64;
					const long long cap = ((A.ncols + 63) / 64) * 64;
					if (pw > cap) pw = cap;
					pp.sp_P = (int)pw;
					const size_t seqM = (size_t)tM->npad * cbytes;
					int stc = 32;
					while (stc > VN && (size_t)(stc / 2) * seqM > 72 * 1024) stc /= 2;
					pp.sp_tc = stc;
					pp.sp_smem = (size_t)(stc / 2) * seqM;
					pp.sp_force_inv = split_inv;
					pp.sp_smem_inv = 2 * (size_t)(16 / 2) * seqM;               // DIT-style inverse: sub-sequence pair of a 16-column tile
					size_t need = (size_t)nn * (size_t)pw * (size_t)P->es;
					// ring sub-pass kernels (dct_colring.cuh): one launch walks all panels; three scratch panels in rotation
					pp.rg_P = 0;
					if (P->prec == 'f' && colring_supports(nn) && (A.ncols % 32) == 0 && (A.ax_is % 4) == 0 && (A.ax_os % 4) == 0 && !getenv("DSP_DCT_NO_COLRING")) {
						const double mb = getenv("DSP_DCT_RING_PANEL_MB") ? atof(getenv("DSP_DCT_RING_PANEL_MB")) : 24.0;   // measured best of 10 / 16 / 24 / 32 MB (profiles/r02_colring.md); >= one item per CTA per segment
						long long rp = (long long)(mb * 1048576.0) / ((long long)nn * 4);
						rp = (rp / 32) * 32;
						if (rp < 32) rp = 32;
						if (rp >= A.ncols) rp = ((A.ncols / 2) / 32) * 32;                  // at least two panels per plane (the kernel's dependency order)
						pp.rg_P = (int)rp;
						const size_t ring_need = ((size_t)colring_scratch_panels() * (size_t)nn * (size_t)(rp > 0 ? rp : 0) * 4 + 1) / 2;   // (the plan allocates nscratch >= 2 parts)
						if (ring_need > need) need = ring_need;
					}
					if (need > P->split_bytes) P->split_bytes = need;
				}
			}
		}
		pp.vec_in_layout = vin; pp.vec_out_layout = vout;
		set_chunk_level(pp, lv);
		// one CTA per SM (big tile): run it with 512 threads; otherwise 256 and rely on several CTAs per SM
		pp.block = (pp.fast && !pp.row && pp.smem > 113 * 1024) ? 2 * kThreads : kThreads;      // (row kernels are built for 256)
		if (pp.fast && !pp.row && getenv("DSP_DCT_THREADS")) pp.block = atoi(getenv("DSP_DCT_THREADS")) >= 512 ? 512 : 256;
		{
			// one resident wave ahead: CTAs per SM by shared memory (<= 2 by registers for the fast kernels) x 148 SMs
			int per_sm = (int)((kMaxSmem) / (pp.smem ? pp.smem : 1));
			if (per_sm > 2) per_sm = 2;
			if (per_sm < 1) per_sm = 1;
			pp.pf_dist = pp.fast ? per_sm * 148 : 0;
			if (getenv("DSP_DCT_PF")) pp.pf_dist = atoi(getenv("DSP_DCT_PF"));
			if (pp.pf_dist >= pp.grid) pp.pf_dist = 0;
		}
		DSP_TRACE("pass %zu: %s%s%s axis=%d n=%d grid=%d smem=%zu vec=%d/%d r0=%d nmid=%d panel=%d tcA=%d", pi, pp.row ? "row" : "col", pp.fast ? "(fast)" : "", pp.split ? "(split)" : "", ax, P->n[ax], pp.grid, pp.smem, (int)vin, (int)vout, pp.ff.r0, pp.ff.nmid, pp.sp_P, pp.sp_tc);
		P->passes.push_back(pp);
	}
	int naux = 2;
	if (getenv("DSP_DCT_NAUX")) { naux = atoi(getenv("DSP_DCT_NAUX")); if (naux < 1) naux = 1; if (naux > 4) naux = 4; }
	P->nscratch = naux;
	if (P->split_bytes && !rt_malloc(&P->d_split, (size_t)naux * P->split_bytes, g_err)) return false;
	if (P->split_bytes && !rt_malloc((void **)&P->d_ring_done, sizeof(int) * 2 * (size_t)colring_max_panels(), g_err)) return false;
#if DSP_GPU
	P->naux = naux;
	if (P->split_bytes && !getenv("DSP_DCT_NO_AUX")) {
		P->aux_ok = cudaEventCreateWithFlags(&P->ev_fork, cudaEventDisableTiming) == cudaSuccess;
		for (int k = 0; k < naux; k++)
			P->aux_ok = P->aux_ok && cudaStreamCreateWithFlags(&P->aux[k], cudaStreamNonBlocking) == cudaSuccess &&
			            cudaEventCreateWithFlags(&P->ev_join[k], cudaEventDisableTiming) == cudaSuccess;
	}
#endif
	return true;
}

// One pass over indices [c0, c0 + nc) of its outermost loop level (the whole pass when nc == pp.ch_cnt).  `in` / `out`
// already point at index c0; the coordinates get c0 back through Outer::base.
static bool run_pass(dsp_dct_plan_s *P, PassPlan &pp, const void *in, void *out, rt_stream st, long long c0, long long nc) {
	const bool ain = ((uintptr_t)in % 16) == 0, aout = ((uintptr_t)out % 16) == 0;
	const bool f32 = P->prec == 'f';
	const bool part = nc < pp.ch_cnt;
	bool ok = false;
	int grid = pp.grid, pf_dist = pp.pf_dist;
	if (pp.row) {
		RowArgs a = pp.ra;
		a.in = in; a.out = out;
		if (part) {
			a.nlines = (int)(nc * pp.ch_inner);
			grid = (a.nlines + a.lines_per_cta - 1) / a.lines_per_cta;
			a.o.base_slot = pp.ch_slot; a.o.base = (int)c0;
			if (pf_dist >= grid) pf_dist = 0;
		}
		a.vec_in = pp.vec_in_layout && ain && !a.in_u8; a.vec_out = pp.vec_out_layout && aout && !a.out_u8;
		a.pf_dist = pf_dist;
		const int nn = pp.ff.n;
		// layout-specialised kernels: planar or RGB lines at one stride, whole line pairs in every CTA, 16-byte access
		const bool spec = pp.fast && f32 && !pp.fused && (a.d == 1 || a.d == 3) && a.simple && a.vec_in && a.vec_out &&
		                  nn >= 256 && nn <= 8192 && (a.lines_per_cta % 2) == 0 && (a.nlines % 2) == 0 &&
		                  !getenv("DSP_DCT_NO_FIXED") && !getenv("DSP_DCT_NO_PLANAR");
		// planar lines, n = 256 .. 8192: the persistent TMA-fed ring kernel (dct_ring.cuh); DSP_DCT_NO_RING=1 keeps the
		// one-shot kernel (A/B measurements)
		static const bool no_ring = getenv("DSP_DCT_NO_RING") != nullptr;
		if (spec && a.d == 1 && !no_ring && ring_supports(nn) && ((a.ls_in | a.ls_out) % 4) == 0) {
			RingArgs r;
			r.in = (const float *)in; r.out = (float *)out;
			r.ls_in = a.ls_in; r.ls_out = a.ls_out; r.nlines = a.nlines;
			r.lscale = (float)(pp.lop.kind == OP_SCALE ? pp.lop.p[0] : 1.0); r.sscale = (float)(pp.sop.kind == OP_SCALE ? pp.sop.p[0] : 1.0);
			r.tw = pp.ff.tw; r.om = pp.ff.om; r.sig = pp.ff.sig;
			DSP_TRACE("row pass: ring kernel n=%d lines=%d", nn, a.nlines);
			ok = launch_row_ring_f32(r, nn, a.kind == DSP_KIND_REDFT10, st, g_err);
		} else if (spec && a.d == 1) ok = launch_row_fast_f32p(a, pp.ff, false, pp.lop, pp.sop, grid, pp.block, pp.smem, st, g_err);
		else if (spec) ok = launch_row_fast_f32i3(a, pp.ff, false, pp.lop, pp.sop, grid, pp.block, pp.smem, st, g_err);
		else if (pp.fast && f32) ok = launch_row_fast_f32(a, pp.ff, pp.fused, pp.lop, pp.sop, grid, pp.block, pp.smem, st, g_err);
#if DSP_FAST_F64
		else if (pp.fast) ok = launch_row_fast_f64(a, pp.ff, pp.fused, pp.lop, pp.sop, grid, pp.block, pp.smem, st, g_err);
#endif
		else ok = f32 ? launch_row_generic_f32(a, pp.ff, pp.fused, pp.lop, pp.sop, grid, pp.block, pp.smem, st, g_err)
		              : launch_row_generic_f64(a, pp.ff, pp.fused, pp.lop, pp.sop, grid, pp.block, pp.smem, st, g_err);
	} else if (pp.split && ain && aout &&
	           (pp.ca.kind == DSP_KIND_REDFT10 || pp.sp_force_inv || (!pp.fused && (pp.ca.ncols % 16) == 0 && !getenv("DSP_DCT_NO_SPLIT_INV")))) {
		// per outer index (batch / frame) and per column panel: sub-pass A then B (forward) or B' then A' (inverse)
		const ColArgs &c = pp.ca;
		const bool fwd = c.kind == DSP_KIND_REDFT10;
		long long no = 1;
		for (int k = 0; k < 4; k++) no *= c.o.cnt[k];
		if (part) no = nc * pp.ch_inner;
		int live = 0, at = 0;
		for (int k = 0; k < 4; k++) if (c.o.cnt[k] > 1) { live++; at = k; }
		// the inverse on the ring is correct but measured SLOWER than the two-kernel DIT split (0.595 vs 0.497 ms per 2 planes
		// of 8192^2: its sub-pass A'' forms every pre-twiddle pair twice and spills): opt-in, for experiments
		static const bool inv_ring = getenv("DSP_DCT_COLRING_INV") != nullptr;
		if ((fwd || (inv_ring && (c.ncols % 32) == 0)) && f32 && !pp.fused && pp.rg_P > 0 && live <= 1 && P->d_ring_done) {
			// ring sub-pass kernels: ONE persistent launch walks every panel of every plane (dct_colring.cuh); the tensor maps
			// depend on the pointers only and are cached in the plan
			const int ppp = (c.ncols + pp.rg_P - 1) / pp.rg_P;
			int per = colring_max_panels() / ppp;
			if (per < 1) per = 1;
			ok = ppp <= colring_max_panels();
			if (!ok) g_err = "ring column pass: too many panels per plane";
			for (long long p0 = 0; p0 < no && ok; p0 += per) {
				const int np = (int)(no - p0 < per ? no - p0 : per);
				const char *bin = (const char *)in + p0 * c.o.is[at] * P->es;
				char *bout = (char *)out + p0 * c.o.os[at] * P->es;
				PassPlan::RingMaps *rm = nullptr;
				for (auto &e : pp.ring_maps) if (e.in == bin && e.out == bout && e.scratch == P->d_split && e.nplanes == np) rm = &e;
				if (!rm) {
					PassPlan::RingMaps e;
					memset(&e.args, 0, sizeof(e.args));
					e.in = bin; e.out = bout; e.scratch = P->d_split; e.nplanes = np;
					ok = colring_encode(e.args, c.f.n, !fwd, (const float *)bin, c.ax_is, c.o.is[at], (float *)bout, c.ax_os, c.o.os[at], np, c.ncols,
					                    (float *)P->d_split, pp.rg_P, g_err);
					e.args.twM = pp.ffM.tw; e.args.sigM = pp.ffM.sig;
					e.args.twN = pp.ff.tw; e.args.omN = pp.ff.om; e.args.sigN = pp.ff.sig;
					e.args.nplanes = np; e.args.ppp = ppp; e.args.P = pp.rg_P; e.args.ncols = c.ncols; e.args.reverse = fwd ? 0 : 1;
					e.args.done = P->d_ring_done;
					e.args.scratch = (float *)P->d_split;
					e.args.trace = nullptr;
					e.args.flags = getenv("DSP_DCT_RING_NODISCARD") ? 1 : 0;
					if (getenv("DSP_DCT_RING_TRACE")) {                       // debug: timeline of CTA 0 (leaked on purpose; read with cudaMemcpy by the tool)
						void *tp = nullptr;
						if (rt_malloc(&tp, 4 * 4096 * sizeof(long long), g_err)) { rt_zero(tp, 4 * 4096 * sizeof(long long), st, g_err); e.args.trace = (long long *)tp; fprintf(stderr, "[dsp_dct] ring trace buffer %p\n", tp); }
					}
					e.args.out = (float *)bout; e.args.ax_os = c.ax_os; e.args.plane_os = c.o.os[at];
					if (ok) { if (pp.ring_maps.size() >= 64) pp.ring_maps.clear(); pp.ring_maps.push_back(e); rm = &pp.ring_maps.back(); }
				}
				if (ok) {
					ColRingArgs ra = rm->args;
					ra.lscale = (float)(pp.lop.kind == OP_SCALE ? pp.lop.p[0] : 1.0); ra.sscale = (float)(pp.sop.kind == OP_SCALE ? pp.sop.p[0] : 1.0);
					DSP_TRACE("split pass: ring sub-pass kernel n=%d planes %d panels/plane %d width %d", c.f.n, np, ppp, pp.rg_P);
					ok = rt_zero(P->d_ring_done, sizeof(int) * 2 * (size_t)colring_max_panels(), st, g_err) && launch_col_ring_f32(ra, c.f.n, !fwd, st, g_err);
					if (ok) g_launches++;
				}
			}
			if (ok) g_launches--;                  // the caller adds one
		} else {
		const int M = c.f.n / 16;
		ok = true;
		rt_stream pst[4] = {st, st, st, st};
		long long panel_no = 0;
#if DSP_GPU
		if (P->aux_ok) {
			cudaEventRecord(P->ev_fork, st);
			for (int k = 0; k < P->naux; k++) { cudaStreamWaitEvent(P->aux[k], P->ev_fork, 0); pst[k] = P->aux[k]; }
		}
#endif
		for (long long oi = 0; oi < no && ok; oi++) {
			long long rem = oi, ioff = 0, ooff = 0;
			SplitArgs sa;
			memset(&sa, 0, sizeof(sa));
			for (int k = 0; k < 4; k++) {
				const long long idx = rem % c.o.cnt[k];
				rem /= c.o.cnt[k];
				ioff += idx * c.o.is[k]; ooff += idx * c.o.os[k];
				sa.cbase.set(c.o.slot[k], (int)idx + ((part && c.o.slot[k] == pp.ch_slot) ? (int)c0 : 0));
			}
			sa.n = c.f.n; sa.M = M; sa.kind = c.kind; sa.d = c.d; sa.dd = c.dd;
			sa.ax_is = c.ax_is; sa.ax_os = c.ax_os; sa.ax_ss = pp.sp_P;
			sa.ax_slot = c.ax_slot; sa.col_slot = c.col_slot;
			sa.in = (const char *)in + ioff * P->es; sa.out = (char *)out + ooff * P->es; sa.scratch = P->d_split;
			sa.tc = pp.sp_tc;
			sa.pf_warps = getenv("DSP_DCT_PFW") ? atoi(getenv("DSP_DCT_PFW")) : 148 * 16;
			const int npanels = (c.ncols + pp.sp_P - 1) / pp.sp_P;
			for (int pi = 0; pi < npanels && ok; pi++) {
				// the inverse walks the panels backwards: a forward+inverse round trip then starts on the columns the
				// forward pass touched last, which are the ones still resident in L2
				const int pn = fwd ? pi : npanels - 1 - pi;
				const int half = (int)(panel_no++ % P->nscratch);
				rt_stream st = pst[half];                       // shadows the caller's stream for this panel's launches
				sa.scratch = (char *)P->d_split + (size_t)half * P->split_bytes;
				sa.pcol0 = pn * pp.sp_P;
				sa.pcols = c.ncols - sa.pcol0 < pp.sp_P ? c.ncols - sa.pcol0 : pp.sp_P;
				sa.ntiles = (sa.pcols + sa.tc - 1) / sa.tc;
				sa.ngroups = ((sa.pcols + 1) / 2 + 31) / 32;
				const int gridA = sa.ntiles * 16, warpsB = sa.ngroups * (M / 2 + 1);
				// start pulling the next panel's columns into L2 while this panel computes
				static const bool pf = getenv("DSP_DCT_PREFETCH") != nullptr;   // measured: no gain, off by default
				if (pf && pi + 1 < npanels) {
					const int nx = fwd ? pn + 1 : pn - 1;
					const int nc0 = nx * pp.sp_P;
					const int ncols = c.ncols - nc0 < pp.sp_P ? c.ncols - nc0 : pp.sp_P;
					ok = launch_l2_prefetch((const char *)sa.in + (size_t)nc0 * P->es, sa.ax_is * P->es, c.f.n, ncols * P->es, st, g_err);
				}
				const bool dit_inv = !fwd && !pp.sp_force_inv;
				if (ok && dit_inv) {
					sa.tci = 16; sa.ntilesi = (sa.pcols + 15) / 16;
					ok = launch_split_inv_fft_f32(sa, pp.ffM, pp.ff, pp.lop, sa.ntilesi * 9, pp.sp_smem_inv, st, g_err) &&
					     launch_split_inv_outer_f32(sa, pp.ff, pp.sop, sa.ngroups * M, st, g_err);
				} else if (ok && fwd && f32) ok = launch_split_fft_f32(sa, pp.ffM, pp.fused, pp.lop, pp.sop, gridA, pp.sp_smem, st, g_err) &&
				              launch_split_outer_f32(sa, pp.ff, pp.fused, pp.lop, pp.sop, warpsB, st, g_err);
				else if (ok && f32) ok = launch_split_outer_f32(sa, pp.ff, pp.fused, pp.lop, pp.sop, warpsB, st, g_err) &&
				          launch_split_fft_f32(sa, pp.ffM, pp.fused, pp.lop, pp.sop, gridA, pp.sp_smem, st, g_err);
#if DSP_FAST_F64
				else if (ok && fwd) ok = launch_split_fft_f64(sa, pp.ffM, pp.fused, pp.lop, pp.sop, gridA, pp.sp_smem, st, g_err) &&
				              launch_split_outer_f64(sa, pp.ff, pp.fused, pp.lop, pp.sop, warpsB, st, g_err);
				else if (ok) ok = launch_split_outer_f64(sa, pp.ff, pp.fused, pp.lop, pp.sop, warpsB, st, g_err) &&
				          launch_split_fft_f64(sa, pp.ffM, pp.fused, pp.lop, pp.sop, gridA, pp.sp_smem, st, g_err);
#endif
				if (ok) g_launches++;          // (the second launch is counted by the caller)
			}
		}
		if (ok) g_launches--;                  // the caller adds one
#if DSP_GPU
		if (P->aux_ok)
			for (int k = 0; k < P->naux; k++) { cudaEventRecord(P->ev_join[k], P->aux[k]); cudaStreamWaitEvent(st, P->ev_join[k], 0); }
#endif
		}
	} else {
		ColArgs a = pp.ca;
		a.in = in; a.out = out;
		a.vec_in = pp.vec_in_layout && ain; a.vec_out = pp.vec_out_layout && aout;
		if (part) {
			grid = (int)((long long)a.ntiles * nc * pp.ch_inner);
			a.o.base_slot = pp.ch_slot; a.o.base = (int)c0;
			if (pf_dist >= grid) pf_dist = 0;
		}
		a.pf_dist = pf_dist;
		if (pp.seg_n > 0) {
			// segment bases become element offsets from this execute's output pointer; 16-byte alignment keeps the lean move
			for (int g = 0; g < pp.seg_n; g++) {
				const long long db = (const char *)pp.seg_base[g] - (const char *)out;
				if (db % 16) { g_err = "output segment bases must be 16-byte aligned relative to the output pointer"; return false; }
				a.seg_off[g] = db / P->es;
			}
			if (!a.vec_in || !a.vec_out) { g_err = "segmented output needs 16-byte aligned buffers"; return false; }
		}
		if (pp.fast && f32) ok = launch_col_fast_f32(a, pp.ff, pp.fused, pp.lop, pp.sop, grid, pp.block, pp.smem, st, g_err);
#if DSP_FAST_F64
		else if (pp.fast) ok = launch_col_fast_f64(a, pp.ff, pp.fused, pp.lop, pp.sop, grid, pp.block, pp.smem, st, g_err);
#endif
		else ok = f32 ? launch_col_generic_f32(a, pp.ff, pp.fused, pp.lop, pp.sop, grid, pp.block, pp.smem, st, g_err)
		              : launch_col_generic_f64(a, pp.ff, pp.fused, pp.lop, pp.sop, grid, pp.block, pp.smem, st, g_err);
	}
	if (ok) g_launches++;
	return ok;
}

// L2 budget of the chunked schedule (bytes of one chunk's intermediate).  B200: 126 MB of L2 in two halves; a chunk and
// the next chunk's input have to sit in it together.  DSP_DCT_L2_CHUNK_MB tunes it, 0 turns the schedule off.
static size_t l2_chunk_bytes() {
	const char *e = getenv("DSP_DCT_L2_CHUNK_MB");             // read per execute: the tests switch it at run time
	const double mb = e ? atof(e) : 0.0;               // measured (profiles/r02_chunk_sweep.md): serial chunks lose to one launch per pass
	return mb > 0 ? (size_t)(mb * 1048576.0) : 0;
}

static bool run_one(dsp_dct_plan_s *P, size_t i, void *d_in, void *d_out, rt_stream st, long long c0, long long nc) {
	PassPlan &pp = P->passes[i];
	// with an 8-bit final store the T-typed intermediate lives in the plan's work buffer
	void *mid = P->d_work ? P->d_work : d_out;
	const bool first = i == 0, last = i + 1 == P->passes.size();
	const char *in = (const char *)(first ? d_in : mid);
	char *out = (char *)(last ? d_out : mid);
	if (nc < pp.ch_cnt) {
		// 8-bit sides (motion's pels) are addressed in bytes
		const long long esi = (pp.row && pp.ra.in_u8) ? 1 : P->es, eso = (pp.row && pp.ra.out_u8) ? 1 : P->es;
		in += c0 * pp.ch_is * esi;
		out += c0 * pp.ch_os * eso;
	}
	DSP_TRACE("launch pass %zu in=%p out=%p chunk %lld+%lld of %lld", i, (const void *)in, (void *)out, c0, nc, pp.ch_cnt);
#if DSP_GPU
	cudaEvent_t e0 = nullptr, e1 = nullptr;
	const bool prof = P->profiling && P->ev.size() < 2 * 16384;  // bounded: unread timings stop accumulating
	if (prof) {
		cudaEventCreate(&e0); cudaEventCreate(&e1);
		cudaEventRecord(e0, st);
	}
#endif
#if DSP_GPU
	const bool nv = nvtx_on();
	if (nv) {
		char name[96];
		snprintf(name, sizeof name, "dct pass %zu %s n=%d", i, pp.row ? "row" : (pp.split ? "col-split" : "col"), pp.row ? pp.ra.f.n : pp.ca.f.n);
		nvtxRangePushA(name);
	}
#endif
	const bool launched = run_pass(P, pp, in, out, st, c0, nc);
#if DSP_GPU
	if (nv) nvtxRangePop();
#endif
	if (!launched) return false;
#if DSP_GPU
	if (prof) {
		cudaEventRecord(e1, st);
		P->ev.push_back(e0); P->ev.push_back(e1);
		P->ev_pass.push_back((int)i);
		P->ev_frac.push_back((float)((double)nc / (double)pp.ch_cnt));
	}
#endif
	return true;
}

static bool run_passes(dsp_dct_plan_s *P, void *d_in, void *d_out, rt_stream st) {
	if (P->need_acc && !rt_zero(P->d_scalars, sizeof(double) * 4, st, g_err)) return false;
	const size_t np = P->passes.size();
	const size_t budget = l2_chunk_bytes();
	size_t i = 0;
	while (i < np) {
		// a run = consecutive passes that loop over the same outermost level (batch elements, frames): chunk by chunk,
		// so that each chunk's intermediate is read back from L2 instead of HBM
		size_t j = i + 1;
		const PassPlan &p0 = P->passes[i];
		// (spec's data-dependent range is resolved between its two passes from sums over the whole first pass)
		const bool can = budget && p0.ch_slot >= 0 && p0.ch_cnt > 1 && P->fuse_kind != 1;
		while (can && j < np && P->passes[j].ch_slot == p0.ch_slot && P->passes[j].ch_cnt == p0.ch_cnt) j++;
		long long nc = p0.ch_cnt;
		if (j - i >= 2) {
			const double per = P->samples_per_launch / (double)p0.ch_cnt * (double)P->es;       // bytes per index of the level
			long long fit = (long long)((double)budget / per);
			if (fit >= 1 && fit < nc) {
				// even chunks, each a whole number of "waves" is not needed: the kernels' grids are per chunk
				const long long nchunks = (p0.ch_cnt + fit - 1) / fit;
				nc = (p0.ch_cnt + nchunks - 1) / nchunks;
			}
		}
		for (long long c0 = 0; c0 < p0.ch_cnt; c0 += nc) {
			const long long n = c0 + nc <= p0.ch_cnt ? nc : p0.ch_cnt - c0;
			for (size_t k = i; k < j; k++) {
				if (!run_one(P, k, d_in, d_out, st, c0, n)) return false;
				if (P->fuse_kind == 1 && k == 0) {
					if (!launch_spec_resolve(P->prec, P->passes.back().sop, P->d_scalars, P->d_scalars + 4, st, g_err)) return false;
					g_launches++;
				}
			}
		}
		i = j;
	}
	P->last_stream = st;
	return true;
}

static dsp_dct_plan make_plan(char prec, int rank, const int *n, int howmany, void *in, const int *inembed, int istride,
                              int idist, void *out, const int *onembed, int ostride, int odist, const int *kind,
                              int nbatch, long long ibdist, long long obdist) {
	g_err.clear();
	if (prec != 'f' && prec != 'd') { g_err = "precision must be 'f' or 'd' (long double has no GPU equivalent)"; return nullptr; }
	if (rank < 1 || rank > 3) { g_err = "rank must be 1..3"; return nullptr; }
	if (!n || !kind) { g_err = "null n/kind"; return nullptr; }
	if (howmany < 1 || nbatch < 1) { g_err = "howmany/nbatch must be >= 1"; return nullptr; }
	for (int i = 0; i < rank; i++) {
		if (n[i] < 1) { g_err = "transform sizes must be >= 1"; return nullptr; }
		if (kind[i] != DSP_DCT_REDFT10 && kind[i] != DSP_DCT_REDFT01) {
			g_err = "only REDFT10 (DCT-II) and REDFT01 (DCT-III) are supported";
			return nullptr;
		}
	}
	std::lock_guard<std::mutex> lock(g_mu);
	DSP_TRACE("plan: prec=%c rank=%d n=%d,%d,%d howmany=%d istride=%d idist=%d nbatch=%d", prec, rank, n[0], rank > 1 ? n[1] : 1, rank > 2 ? n[2] : 1, howmany, istride, idist, nbatch);
	if (!rt_init(g_err)) return nullptr;
	DSP_TRACE("runtime ok, device %d", rt_device());
	dsp_dct_plan_s *P = new dsp_dct_plan_s();
	P->prec = prec;
	P->es = prec == 'f' ? 4 : 8;
	P->rank = rank;
	for (int i = 0; i < 3; i++) { P->n[i] = i < rank ? n[i] : 1; P->kind[i] = i < rank ? kind[i] : 0; }
	P->device = rt_device();
	P->in = in; P->out = out;
	P->d_in = P->d_out = nullptr;
	P->d_in_bytes = P->d_out_bytes = 0;
	P->fuse_kind = 0;
	P->d_scalars = nullptr;
	P->d_signmap = nullptr;
	P->d_work = nullptr;
	P->d_split = nullptr;
	P->d_ring_done = nullptr;
	P->split_bytes = 0;
	P->nscratch = 1;
#if DSP_GPU
	P->aux_ok = false;
#endif
	P->need_acc = false;
	P->last_stream = 0;
	P->profiling = false;
	P->samples_per_launch = 0;
	P->c_howmany = howmany; P->c_istride = istride; P->c_idist = idist; P->c_ostride = ostride; P->c_odist = odist;
	P->c_nbatch = nbatch; P->c_ibdist = ibdist; P->c_obdist = obdist;
	P->c_has_ie = inembed != nullptr; P->c_has_oe = onembed != nullptr;
	for (int i = 0; i < 3; i++) { P->c_ie[i] = (inembed && i < rank) ? inembed[i] : 0; P->c_oe[i] = (onembed && i < rank) ? onembed[i] : 0; }
	P->kids_failed = false;
	P->mg = nullptr; P->mg_tried = false;
	if (!build_plan(P, howmany, inembed, istride, idist, onembed, ostride, odist, nbatch, ibdist, obdist)) {
		delete P;
		return nullptr;
	}
	return P;
}

static void destroy_plan(dsp_dct_plan_s *p);

static bool ensure_staging(dsp_dct_plan_s *P, bool inplace) {
	const size_t ib = P->in_span * (size_t)P->es, ob = P->out_span * (size_t)P->es;
	if (inplace) {
		const size_t need = ib > ob ? ib : ob;
		if (P->d_in_bytes < need) {
			rt_free(P->d_in);
			P->d_in = nullptr; P->d_in_bytes = 0;
			if (!rt_malloc(&P->d_in, need, g_err)) return false;
			P->d_in_bytes = need;
		}
		return true;
	}
	if (P->d_in_bytes < ib) {
		rt_free(P->d_in);
		P->d_in = nullptr; P->d_in_bytes = 0;
		if (!rt_malloc(&P->d_in, ib, g_err)) return false;
		P->d_in_bytes = ib;
	}
	if (P->d_out_bytes < ob) {
		rt_free(P->d_out);
		P->d_out = nullptr; P->d_out_bytes = 0;
		if (!rt_malloc(&P->d_out, ob, g_err)) return false;
		P->d_out_bytes = ob;
	}
	return true;
}

#if DSP_GPU
// Host-buffer execution of a batch: PCIe is the bound (one copy in, one copy out), and the two directions are
// independent engines.  The batch is cut into up to four chunks of whole batch elements, each with its own sub-plan
// and stream: chunk c+1 uploads while chunk c transforms and chunk c-1 downloads.
static bool ensure_kids(dsp_dct_plan_s *P) {
	if (!P->kids.empty()) return true;
	if (P->kids_failed) return false;
	const int nchunks = P->c_nbatch < 4 ? P->c_nbatch : 4;
	long long b0 = 0;
	for (int c = 0; c < nchunks; c++) {
		const int nb = P->c_nbatch / nchunks + (c < P->c_nbatch % nchunks ? 1 : 0);
		dsp_dct_plan_s *K = new dsp_dct_plan_s(*P);             // same geometry; owned resources reset below
		K->passes.clear();
		K->d_in = K->d_out = nullptr; K->d_in_bytes = K->d_out_bytes = 0;
		K->d_scalars = nullptr; K->d_signmap = nullptr; K->d_work = nullptr; K->d_split = nullptr; K->d_ring_done = nullptr;
		K->split_bytes = 0; K->nscratch = 1; K->aux_ok = false;
		K->kids.clear(); K->kid_b0.clear(); K->kids_failed = true;  // no recursion
		K->ev.clear(); K->ev_pass.clear(); K->ev_frac.clear(); K->profiling = false;
		K->c_nbatch = nb;
		bool ok = build_plan(K, P->c_howmany, P->c_has_ie ? P->c_ie : nullptr, P->c_istride, P->c_idist,
		                     P->c_has_oe ? P->c_oe : nullptr, P->c_ostride, P->c_odist, nb, P->c_ibdist, P->c_obdist);
		ok = ok && rt_ok(cudaStreamCreateWithFlags(&P->kid_st[c], cudaStreamNonBlocking), g_err, "stream create");
		if (!ok) {
			destroy_plan(K);
			for (size_t k = 0; k < P->kids.size(); k++) { destroy_plan(P->kids[k]); cudaStreamDestroy(P->kid_st[k]); }
			P->kids.clear(); P->kid_b0.clear();
			P->kids_failed = true;
			return false;
		}
		P->kids.push_back(K);
		P->kid_b0.push_back(b0);
		b0 += nb;
	}
	return true;
}

static bool execute_host_chunks(dsp_dct_plan_s *P, void *in, void *out, void *din, void *dout) {
	const size_t es = (size_t)P->es;
	for (size_t c = 0; c < P->kids.size(); c++) {
		dsp_dct_plan_s *K = P->kids[c];
		// the plain scale factors may have been set after the chunks were planned
		K->passes.front().lop = P->passes.front().lop;
		K->passes.back().sop = P->passes.back().sop;
		const size_t io = (size_t)(P->kid_b0[c] * P->c_ibdist) * es, oo = (size_t)(P->kid_b0[c] * P->c_obdist) * es;
		rt_stream st = P->kid_st[c];
		if (!rt_h2d((char *)din + io, (const char *)in + io, K->in_span * es, st, g_err)) return false;
		if (!run_passes(K, (char *)din + io, (char *)dout + oo, st)) return false;
		if (!rt_d2h((char *)out + oo, (const char *)dout + oo, K->out_span * es, st, g_err)) return false;
	}
	bool ok = true;
	for (size_t c = 0; c < P->kids.size(); c++) ok = rt_sync(P->kid_st[c], g_err) && ok;
	return ok;
}

static bool chunkable(const dsp_dct_plan_s *P) {
	if (P->c_nbatch < 2 || P->fuse_kind != 0 || P->out_has_gaps || P->profiling || P->need_acc) return false;
	if (getenv("DSP_DCT_NO_PIPELINE")) return false;
	const size_t min_mb = getenv("DSP_DCT_PIPELINE_MIN_MB") ? (size_t)atoi(getenv("DSP_DCT_PIPELINE_MIN_MB")) : 32;
	if (P->in_span * (size_t)P->es < (min_mb << 20)) return false;              // small jobs: latency, not bandwidth
	// batch elements must not interleave in memory
	const long long ein = (long long)P->in_span - (long long)(P->c_nbatch - 1) * P->c_ibdist;
	const long long eout = (long long)P->out_span - (long long)(P->c_nbatch - 1) * P->c_obdist;
	return P->c_ibdist >= ein && P->c_obdist >= eout && ein > 0 && eout > 0;
}
#endif

// ------------------------------------------------------------------------------------------------ several GPUs, one process
// fftw_plan_with_nthreads(n) (motion/motion.c:485-486, scan/scan.c:289-290) maps to dsp_dct_plan_with_ngpus(n): a rank-3
// float plan over one contiguous [D][H][W] volume in HOST memory (motion -b 0x0x0: the whole clip is one block) is then
// executed on G = min(n, visible GPUs) devices of this process, SURVEY 8e's slab decomposition behind the C ABI:
//   forward  host slab g (frames [g D/G, (g+1) D/G)) -> GPU g -> REDFT10 over (h, w) per frame, its last pass storing rows
//            [r H/G, (r+1) H/G) straight into GPU r's column buffer over NVLink (dsp_dct_set_output_segments: the exchange is
//            fused into the transform) -> REDFT10 over d on every GPU's [D][H W / G] columns -> strided copy-out to the host
//   inverse  the mirror image: column slices in, REDFT01 over d storing frames [r D/G, ..) into GPU r's slab, REDFT01 over
//            (h, w), slabs out.
// One exchange per transform (the host layout absorbs the other one).  Anything not eligible (double, embedded sub-boxes,
// fused stages, sizes that do not divide, no peer access) stays on one GPU.
static std::atomic<int> g_ngpus(1);

struct MultiGpu {
	int G, D, H, W, Dl;
	long long Pl;
	bool fwd;
	std::vector<dsp_dct_plan> p2, pt;        // per device: frames plan, temporal plan
	std::vector<float *> slab, cols;         // per device: [Dl][H][W], [D][Pl]
};

#if DSP_GPU
static void mg_free(MultiGpu *m) {
	if (!m) return;
	int cur = 0;
	cudaGetDevice(&cur);
	for (int g = 0; g < m->G; g++) {
		cudaSetDevice(g);
		if (g < (int)m->p2.size() && m->p2[g]) dsp_dct_destroy(m->p2[g]);
		if (g < (int)m->pt.size() && m->pt[g]) dsp_dct_destroy(m->pt[g]);
		if (g < (int)m->slab.size()) rt_free(m->slab[g]);
		if (g < (int)m->cols.size()) rt_free(m->cols[g]);
	}
	cudaSetDevice(cur);
	delete m;
}

static bool mg_eligible(const dsp_dct_plan_s *P, int &G) {
	G = g_ngpus.load();
	if (G < 2 || P->prec != 'f' || P->rank != 3 || P->d != 1 || P->c_howmany != 1 || P->c_nbatch != 1 || P->fuse_kind != 0) return false;
	if (P->c_istride != 1 || P->c_ostride != 1 || P->out_has_gaps) return false;
	for (int i = 0; i < 3; i++) {
		if (P->kind[i] != P->kind[0]) return false;
		if (P->c_has_ie && P->c_ie[i] != P->n[i]) return false;
		if (P->c_has_oe && P->c_oe[i] != P->n[i]) return false;
	}
	if (P->passes.front().lop.kind != 0 || P->passes.back().sop.kind != 0) return false;      // plain scales stay on one GPU
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess) return false;
	if (G > ndev) G = ndev;
	while (G > 1 && (P->n[0] % G || P->n[1] % G)) G--;
	if (G < 2) return false;
	for (int a = 0; a < G; a++)
		for (int b = 0; b < G; b++) {
			int can = 0;
			if (a != b && (cudaDeviceCanAccessPeer(&can, a, b) != cudaSuccess || !can)) return false;
		}
	return true;
}

static MultiGpu *mg_create(const dsp_dct_plan_s *P, int G) {
	MultiGpu *m = new MultiGpu();
	m->G = G; m->D = P->n[0]; m->H = P->n[1]; m->W = P->n[2]; m->Dl = m->D / G;
	m->Pl = (long long)m->H * m->W / G;
	m->fwd = P->kind[0] == DSP_DCT_REDFT10;
	m->p2.assign(G, nullptr); m->pt.assign(G, nullptr); m->slab.assign(G, nullptr); m->cols.assign(G, nullptr);
	const int hw[2] = {m->H, m->W}, kk[2] = {P->kind[0], P->kind[0]}, nd = m->D;
	const long long fhw = (long long)m->H * m->W;
	int cur = 0;
	cudaGetDevice(&cur);
	bool ok = true;
	for (int g = 0; g < G && ok; g++) {
		cudaSetDevice(g);
		for (int r = 0; r < G; r++)
			if (r != g) {
				const cudaError_t e = cudaDeviceEnablePeerAccess(r, 0);
				if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = false;
				(void)cudaGetLastError();
			}
		ok = ok && rt_malloc((void **)&m->slab[g], (size_t)m->Dl * fhw * 4, g_err) && rt_malloc((void **)&m->cols[g], (size_t)m->D * m->Pl * 4, g_err);
		if (!ok) break;
		m->p2[g] = dsp_dct_plan_many_batched('f', 2, hw, 1, nullptr, nullptr, 1, 0, nullptr, nullptr, 1, 0, kk, 0, m->Dl, (ptrdiff_t)fhw, (ptrdiff_t)fhw);
		m->pt[g] = dsp_dct_plan_many_batched('f', 1, &nd, (int)m->Pl, nullptr, nullptr, (int)m->Pl, 1, nullptr, nullptr, (int)m->Pl, 1, kk, 0, 1, 0, 0);
		ok = m->p2[g] && m->pt[g];
	}
	// the pass in front of the exchange stores into every device's receive buffer
	for (int g = 0; g < G && ok; g++) {
		cudaSetDevice(g);
		void *bases[8];
		if (m->fwd) {
			for (int r = 0; r < G; r++) bases[r] = m->cols[r] + (size_t)g * m->Dl * m->Pl;       // frame dl of GPU g -> cols[r][g Dl + dl][..]
			ok = dsp_dct_set_output_segments(m->p2[g], G, m->H / G, bases, m->Pl, 0) == 0;
		} else {
			for (int r = 0; r < G; r++) bases[r] = m->slab[r] + (size_t)g * m->Pl;             // frame d of GPU r's slab, columns [g Pl, ..)
			ok = dsp_dct_set_output_segments(m->pt[g], G, m->Dl, bases, 0, fhw) == 0;
		}
	}
	cudaSetDevice(cur);
	if (!ok) { const std::string why = g_err; mg_free(m); g_err = why; return nullptr; }
	return m;
}

// host [D][H][W] in -> out (may be the same buffer); every device's work is enqueued from this thread and runs concurrently
static bool mg_execute(MultiGpu *m, const float *in, float *out) {
	const int G = m->G;
	const long long fhw = (long long)m->H * m->W;
	int cur = 0;
	cudaGetDevice(&cur);
	bool ok = true;
	auto sync_all = [&]() { for (int g = 0; g < G; g++) { cudaSetDevice(g); ok = rt_ok(cudaDeviceSynchronize(), g_err, "multi-GPU synchronize") && ok; } };
	if (m->fwd) {
		for (int g = 0; g < G && ok; g++) {
			cudaSetDevice(g);
			ok = rt_h2d(m->slab[g], in + (size_t)g * m->Dl * fhw, (size_t)m->Dl * fhw * 4, 0, g_err) &&
			     dsp_dct_execute_dev(m->p2[g], m->slab[g], m->slab[g], nullptr) == 0;          // (its last pass stores into the peers' cols)
		}
		sync_all();                                                                          // the exchange is complete on every device
		for (int g = 0; g < G && ok; g++) {
			cudaSetDevice(g);
			ok = dsp_dct_execute_dev(m->pt[g], m->cols[g], m->cols[g], nullptr) == 0 &&
			     rt_ok(cudaMemcpy2DAsync(out + (size_t)g * m->Pl, (size_t)fhw * 4, m->cols[g], (size_t)m->Pl * 4, (size_t)m->Pl * 4, (size_t)m->D,
			                             cudaMemcpyDeviceToHost, 0), g_err, "multi-GPU copy-out");
		}
		sync_all();
	} else {
		for (int g = 0; g < G && ok; g++) {
			cudaSetDevice(g);
			ok = rt_ok(cudaMemcpy2DAsync(m->cols[g], (size_t)m->Pl * 4, in + (size_t)g * m->Pl, (size_t)fhw * 4, (size_t)m->Pl * 4, (size_t)m->D,
			                             cudaMemcpyHostToDevice, 0), g_err, "multi-GPU copy-in") &&
			     dsp_dct_execute_dev(m->pt[g], m->cols[g], m->cols[g], nullptr) == 0;          // (stores into the peers' slabs)
		}
		sync_all();
		for (int g = 0; g < G && ok; g++) {
			cudaSetDevice(g);
			ok = dsp_dct_execute_dev(m->p2[g], m->slab[g], m->slab[g], nullptr) == 0 &&
			     rt_d2h(out + (size_t)g * m->Dl * fhw, m->slab[g], (size_t)m->Dl * fhw * 4, 0, g_err);
		}
		sync_all();
	}
	cudaSetDevice(cur);
	return ok;
}
#endif

static bool execute_host(dsp_dct_plan_s *P, void *in, void *out) {
#if DSP_GPU
	if (!P->mg_tried) {
		P->mg_tried = true;
		int G = 1;
		if (mg_eligible(P, G)) {
			P->mg = mg_create(P, G);
			if (!P->mg) { DSP_TRACE("multi-GPU setup failed (%s): single GPU", g_err.c_str()); g_err.clear(); }
			else DSP_TRACE("multi-GPU plan: %d devices, slabs of %d frames, %lld columns each", P->mg->G, P->mg->Dl, P->mg->Pl);
		}
	}
	if (P->mg) return mg_execute(P->mg, (const float *)in, (float *)out);
#endif
	const bool inplace = in == out;
	if (!ensure_staging(P, inplace)) return false;
	void *din = P->d_in, *dout = inplace ? P->d_in : P->d_out;
#if DSP_GPU
	if (chunkable(P) && ensure_kids(P)) return execute_host_chunks(P, in, out, din, dout);
#endif
	rt_stream st = 0;
	if (!rt_h2d(din, in, P->in_span * (size_t)P->es, st, g_err)) return false;
	if (!inplace && P->out_has_gaps && !rt_h2d(dout, out, P->out_span * (size_t)P->es, st, g_err)) return false;
	if (!run_passes(P, din, dout, st)) return false;
	if (!rt_d2h(out, dout, P->out_span * (size_t)P->es, st, g_err)) return false;
	return rt_sync(st, g_err);
}

static bool ensure_scalars(dsp_dct_plan_s *P) {
	if (P->d_scalars) return true;
	if (!rt_malloc((void **)&P->d_scalars, sizeof(double) * 16, g_err)) return false;
	return rt_zero(P->d_scalars, sizeof(double) * 16, 0, g_err) && rt_sync(0, g_err);
}

static void destroy_plan(dsp_dct_plan_s *p) {
#if DSP_GPU
	for (size_t k = 0; k < p->kids.size(); k++) { cudaStreamSynchronize(p->kid_st[k]); destroy_plan(p->kids[k]); cudaStreamDestroy(p->kid_st[k]); }
	if (p->aux_ok) for (int k = 0; k < p->naux; k++) cudaStreamSynchronize(p->aux[k]);
#endif
#if DSP_GPU
	mg_free(p->mg);
#endif
	rt_free(p->d_in);
	rt_free(p->d_out);
	rt_free(p->d_scalars);
	rt_free(p->d_signmap);
	rt_free(p->d_work);
	rt_free(p->d_split);
	rt_free(p->d_ring_done);
#if DSP_GPU
	if (p->aux_ok) {
		cudaEventDestroy(p->ev_fork);
		for (int k = 0; k < p->naux; k++) { cudaStreamDestroy(p->aux[k]); cudaEventDestroy(p->ev_join[k]); }
	}
	for (cudaEvent_t e : p->ev) cudaEventDestroy(e);
#endif
	delete p;
}

}  // namespace dsp

// ================================================================================================ C ABI
extern "C" {

dsp_dct_plan dsp_dct_plan_many(char prec, int rank, const int *n, int howmany, void *in, const int *inembed, int istride,
                               int idist, void *out, const int *onembed, int ostride, int odist, const int *kind,
                               unsigned flags) {
	(void)flags;
	return make_plan(prec, rank, n, howmany, in, inembed, istride, idist, out, onembed, ostride, odist, kind, 1, 0, 0);
}

dsp_dct_plan dsp_dct_plan_many_batched(char prec, int rank, const int *n, int howmany, void *in, const int *inembed,
                                       int istride, int idist, void *out, const int *onembed, int ostride, int odist,
                                       const int *kind, unsigned flags, int nbatch, ptrdiff_t ibdist, ptrdiff_t obdist) {
	(void)flags;
	return make_plan(prec, rank, n, howmany, in, inembed, istride, idist, out, onembed, ostride, odist, kind, nbatch,
	                 (long long)ibdist, (long long)obdist);
}

dsp_dct_plan dsp_dct_plan_2d(char prec, int n0, int n1, void *in, void *out, int kind0, int kind1, unsigned flags) {
	const int n[2] = {n0, n1}, kind[2] = {kind0, kind1};
	(void)flags;
	return make_plan(prec, 2, n, 1, in, nullptr, 1, 0, out, nullptr, 1, 0, kind, 1, 0, 0);
}

int dsp_dct_execute_host(dsp_dct_plan p, void *in, void *out) {
	g_err.clear();
	if (!p || !in || !out) { g_err = "null plan or buffer"; return 1; }
	return execute_host(p, in, out) ? 0 : 1;
}

int dsp_dct_execute_dev(dsp_dct_plan p, void *d_in, void *d_out, void *stream) {
	g_err.clear();
	if (!p || !d_in || !d_out) { g_err = "null plan or buffer"; return 1; }
	return run_passes(p, d_in, d_out, (rt_stream)stream) ? 0 : 1;
}

void dsp_dct_execute(dsp_dct_plan p) {
	g_err.clear();
	if (!p) { g_err = "null plan"; return; }
	if (!p->in || !p->out) {
		g_err = "plan was created without buffers: use dsp_dct_execute_host / dsp_dct_execute_dev";
		fprintf(stderr, "dsp_dct_execute: %s\n", g_err.c_str());
		return;
	}
	bool ok;
	if (rt_is_device_ptr(p->in)) ok = run_passes(p, p->in, p->out, 0) && rt_sync(0, g_err);
	else ok = execute_host(p, p->in, p->out);
	if (!ok) fprintf(stderr, "dsp_dct_execute: %s\n", g_err.c_str());
}

void dsp_dct_destroy(dsp_dct_plan p) {
	if (!p) return;
	dsp::destroy_plan(p);
}

void dsp_dct_plan_with_ngpus(int n) { g_ngpus.store(n < 1 ? 1 : (n > 8 ? 8 : n)); }
int dsp_dct_plan_ngpus(dsp_dct_plan p) {
#if DSP_GPU
	return (p && p->mg) ? p->mg->G : 1;
#else
	(void)p;
	return 1;
#endif
}

void *dsp_dct_alloc(size_t bytes) {
	g_err.clear();
	std::string e;
	if (!rt_init(e)) { g_err = e; return nullptr; }
	void *p = rt_host_alloc(bytes);
	if (!p) g_err = "pinned host allocation failed";
	return p;
}

void dsp_dct_free(void *p) { rt_host_free(p); }

void dsp_dct_cleanup(void) {
	std::lock_guard<std::mutex> lock(g_mu);
	std::lock_guard<std::mutex> tab_lock(g_tab_mu);
	for (auto &kv : g_tables) {
		Tables *t = kv.second;
		rt_free(t->tw); rt_free(t->om); rt_free(t->pos2); rt_free(t->pos3); rt_free(t->sig); rt_free(t->ctab);
		delete t;
	}
	g_tables.clear();
	block_mm_cleanup();
}

const char *dsp_dct_last_error(void) { return g_err.c_str(); }

unsigned long long dsp_dct_launch_count(void) { return g_launches.load(); }

int dsp_dct_is_emulation(void) { return DSP_GPU ? 0 : 1; }

int dsp_dct_profile(dsp_dct_plan p, int enable) {
	if (!p) return 1;
	p->profiling = enable != 0;
	return 0;
}

int dsp_dct_num_passes(dsp_dct_plan p) { return p ? (int)p->passes.size() : 0; }

int dsp_dct_pass_stat_get(dsp_dct_plan p, int i, dsp_dct_pass_stat *out) {
	g_err.clear();
	if (!p || !out || i < 0 || i >= (int)p->passes.size()) { g_err = "pass index out of range"; return 1; }
	const PassPlan &pp = p->passes[(size_t)i];
	memset(out, 0, sizeof(*out));
	out->is_row = pp.row ? 1 : 0;
	out->axis = pp.axis;
	out->n = p->n[pp.axis];
	out->grid = pp.grid;
	out->block = pp.block;
	out->smem_bytes = pp.smem;
	out->samples = p->samples_per_launch;
	out->split_panels = pp.split ? (pp.ca.ncols + pp.sp_P - 1) / pp.sp_P : 0;
#if DSP_GPU
	std::vector<cudaEvent_t> keep;
	std::vector<int> keep_pass;
	std::vector<float> keep_frac;
	double execs = 0;                                           // a chunked pass is several launches per execute
	for (size_t k = 0; k < p->ev_pass.size(); k++) {
		cudaEvent_t e0 = p->ev[2 * k], e1 = p->ev[2 * k + 1];
		if (p->ev_pass[k] != i) { keep.push_back(e0); keep.push_back(e1); keep_pass.push_back(p->ev_pass[k]); keep_frac.push_back(p->ev_frac[k]); continue; }
		float ms = 0;
		if (cudaEventSynchronize(e1) == cudaSuccess && cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) {
			out->ms_total += ms;
			execs += p->ev_frac[k];
			out->kernel_launches++;
		}
		cudaEventDestroy(e0); cudaEventDestroy(e1);
	}
	out->launches = (int)(execs + 0.5);
	p->ev.swap(keep);
	p->ev_pass.swap(keep_pass);
	p->ev_frac.swap(keep_frac);
#endif
	return 0;
}

int dsp_dct_fuse_scale(dsp_dct_plan p, double load_scale, double store_scale) {
	g_err.clear();
	if (!p) { g_err = "null plan"; return 1; }
	if (p->fuse_kind) { g_err = "plan already carries a fused stage"; return 1; }
	PassPlan &f = p->passes.front(), &l = p->passes.back();
	// a plain multiply rides in the lean kernels (no OpAny dispatch needed)
	if (load_scale != 1.0) { f.lop.kind = OP_SCALE; f.lop.p[0] = load_scale; }
	if (store_scale != 1.0) { l.sop.kind = OP_SCALE; l.sop.p[0] = store_scale; }
	return 0;
}

int dsp_dct_set_output_segments(dsp_dct_plan p, int nseg, int seg_rows, void *const *bases, long long outer_stride,
                                long long row_stride) {
	g_err.clear();
	if (!p) { g_err = "null plan"; return 1; }
	PassPlan &l = p->passes.back();
	if (l.seg_saved) {                                      // every call starts from the plan's own strides
		for (int k = 0; k < 4; k++) l.ca.o.os[k] = l.seg_os[k];
		l.ca.ax_os = l.seg_ax_os;
		l.seg_saved = false;
	}
	if (nseg == 0) { l.seg_n = 0; l.ca.seg_rows = 0; return 0; }
	if (!bases || nseg < 1 || nseg > 8 || seg_rows < 1) { g_err = "segments: 1..8 segments of >= 1 rows"; return 1; }
	if (l.row || l.split || p->prec != 'f') { g_err = "segmented output needs a float plan whose last pass is a one-kernel strided-axis pass"; return 1; }
	ColArgs &c = l.ca;
	if (c.f.n != nseg * seg_rows) { g_err = "segments must tile the last axis exactly"; return 1; }
	// the segmented store moves whole tiles: narrow the tile until it divides the columns (short axes get wide tiles,
	// e.g. 256 columns at n = 32, which need not divide a slab's h*w / G columns)
	if ((c.tc % 4) == 0 && (c.ncols % 4) == 0 && (c.ncols % c.tc) != 0) {
		const size_t seqb = l.smem / (size_t)((c.tc + 1) / 2);
		long long no = 1;
		for (int k = 0; k < 4; k++) no *= c.o.cnt[k];
		while (c.tc > 4 && (c.ncols % c.tc) != 0) c.tc /= 2;
		c.ntiles = c.ncols / c.tc;
		c.dtiles = mk_fd((uint32_t)c.ntiles);
		l.grid = (int)((long long)c.ntiles * no);
		l.smem = (size_t)(c.tc / 2) * seqb;
		if (l.pf_dist >= l.grid) l.pf_dist = 0;
	}
	const int gpr = c.tc / 4;
	if (c.f.dense) { g_err = "segmented output: the last axis has a prime factor above 13 (dense transform), which has no segmented store"; return 1; }
	if ((c.tc % 4) || (gpr & (gpr - 1)) || (l.block % gpr) || (c.ncols % c.tc) || !l.vec_in_layout || !l.vec_out_layout || c.f.dense || l.fused) {
		g_err = "segmented output needs full, 16-byte aligned column tiles and no fused stage";
		return 1;
	}
	for (int k = 0; k < 4; k++) l.seg_os[k] = c.o.os[k];
	l.seg_ax_os = c.ax_os;
	if (outer_stride) {
		int live = 0, at = -1;
		for (int k = 0; k < 4; k++) if (c.o.cnt[k] > 1) { live++; at = k; }
		if (live > 1) { g_err = "outer stride override needs a single outer level"; return 1; }
		if (at >= 0) { c.o.os[at] = outer_stride; l.seg_saved = true; }
	}
	if (row_stride) {
		if (row_stride % 4) { g_err = "row stride override must keep 16-byte alignment"; return 1; }
		c.ax_os = row_stride;
		l.seg_saved = true;
	}
	c.seg_rows = seg_rows;
	c.dseg = mk_fd((uint32_t)seg_rows);
	l.seg_n = nseg;
	for (int g = 0; g < nseg; g++) l.seg_base[g] = bases[g];
	return 0;
}

int dsp_block_quant(char prec, void *d_coeffs, int D, int H, int W, int bd, int bh, int bw, double quantizer,
                    unsigned long long *d_count, void *stream) {
	g_err.clear();
	if ((prec != 'f' && prec != 'd') || !d_coeffs || D < 1 || H < 1 || W < 1 || bd < 1 || bh < 1 || bw < 1) { g_err = "block quant: bad arguments"; return 1; }
	if (!rt_init(g_err)) return 1;
	const bool ok = launch_block_quant(prec, d_coeffs, (long long)D * H * W, H, W, bd, bh, bw, quantizer, d_count, (rt_stream)stream, g_err);
	if (ok) g_launches++;
	return ok ? 0 : 1;
}

int dsp_block_dquant(void *d_coeffs, int D, int H, int W, int bd, int bh, int bw, double quantizer, unsigned long long *d_count, void *stream) {
	g_err.clear();
	if (!d_coeffs || D < 1 || H < 1 || W < 1 || bh < 1 || bw < 1 || !block_dquant_supports(bd) || D % bd) { g_err = "block d-axis + quantiser: bad arguments (depth 2, 4, 8 or 16 dividing D)"; return 1; }
	if (!rt_init(g_err)) return 1;
	const bool ok = launch_block_dquant((float *)d_coeffs, D, H, W, bd, bh, bw, quantizer, d_count, (rt_stream)stream, g_err);
	if (ok) g_launches++;
	return ok ? 0 : 1;
}

int dsp_block_store_u8(char prec, const void *d_coeffs, unsigned char *d_pels, long long n, double scale, void *stream) {
	g_err.clear();
	if ((prec != 'f' && prec != 'd') || !d_coeffs || !d_pels || n < 1) { g_err = "block store: bad arguments"; return 1; }
	if (!rt_init(g_err)) return 1;
	const bool ok = launch_block_store_u8(prec, d_coeffs, d_pels, n, scale, (rt_stream)stream, g_err);
	if (ok) g_launches++;
	return ok ? 0 : 1;
}

struct dsp_motion_tiled_s {
	int D, H, W, bd, bh, bw;
	double quant;
	bool gemm;
	std::vector<dsp_dct_plan> fwd, inv;      // per-axis plans in execution order
	float *work;
	unsigned long long *d_count;
};

void dsp_motion_tiled_destroy(dsp_motion_tiled t) {
	if (!t) return;
	for (dsp_dct_plan p : t->fwd) dsp_dct_destroy(p);
	for (dsp_dct_plan p : t->inv) dsp_dct_destroy(p);
	rt_free(t->work);
	rt_free(t->d_count);
	delete t;
}

dsp_motion_tiled dsp_motion_tiled_create(int D, int H, int W, int bd, int bh, int bw, double quant) {
	g_err.clear();
	if (D < 1 || H < 1 || W < 1 || bd < 1 || bh < 1 || bw < 1 || D % bd || H % bh || W % bw) {
		g_err = "tiled motion: the volume must be a whole number of blocks (the reference pads the last block with zeros)";
		return nullptr;
	}
	if (!rt_init(g_err)) return nullptr;
	dsp_motion_tiled t = new dsp_motion_tiled_s();
	t->D = D; t->H = H; t->W = W; t->bd = bd; t->bh = bh; t->bw = bw; t->quant = quant;
	t->gemm = bh == bw && block_mm_supports(bw);
	t->work = nullptr; t->d_count = nullptr;
	const long long hw = (long long)H * W;
	bool ok = true;
	for (int dir = 0; dir < 2 && ok; dir++) {
		const int kind = dir == 0 ? DSP_DCT_REDFT10 : DSP_DCT_REDFT01;
		std::vector<dsp_dct_plan> &v = dir == 0 ? t->fwd : t->inv;
		if (!t->gemm) {
			// w: contiguous segments of bw; h: bh rows at stride W, W adjacent columns, one batch element per band of bh rows
			v.push_back(dsp_dct_plan_many_batched('f', 1, &bw, (int)((W / bw) * (long long)H * D), nullptr, nullptr, 1, bw, nullptr, nullptr, 1, bw, &kind, 0, 1, 0, 0));
			v.push_back(dsp_dct_plan_many_batched('f', 1, &bh, W, nullptr, nullptr, W, 1, nullptr, nullptr, W, 1, &kind, 0, (H / bh) * D, (ptrdiff_t)bh * W, (ptrdiff_t)bh * W));
		}
		if ((bd > 1 && !(t->gemm && block_dquant_supports(bd))) || !t->gemm)    // d: bd frames at stride H W, H W adjacent columns, one batch element per slab of bd frames
			v.push_back(dsp_dct_plan_many_batched('f', 1, &bd, (int)hw, nullptr, nullptr, (int)hw, 1, nullptr, nullptr, (int)hw, 1, &kind, 0, D / bd, (ptrdiff_t)bd * hw, (ptrdiff_t)bd * hw));
		for (dsp_dct_plan p : v) ok = ok && p != nullptr;
	}
	if (ok) std::reverse(t->inv.begin(), t->inv.end());
	const std::string keep = g_err;
	if (ok) {
		ok = rt_malloc((void **)&t->work, (size_t)D * hw * sizeof(float), g_err) && rt_malloc((void **)&t->d_count, sizeof(unsigned long long), g_err);
	}
	if (!ok) {
		const std::string why = keep.empty() ? g_err : keep;
		dsp_motion_tiled_destroy(t);
		g_err = why.empty() ? "tiled motion: setup failed" : why;
		return nullptr;
	}
	return t;
}

int dsp_motion_tiled_process_dev(dsp_motion_tiled t, const unsigned char *d_in, unsigned char *d_out, unsigned long long *coeffs_coded, void *stream) {
	g_err.clear();
	if (!t || !d_in || !d_out) { g_err = "tiled motion: null argument"; return 1; }
	rt_stream st = (rt_stream)stream;
	const long long n = (long long)t->D * t->H * t->W;
	const double vol = (double)t->bd * t->bh * t->bw;
	if (!launch_block_load_u8('f', d_in, t->work, n, st, g_err)) return 1;                      // motion.c:618-624
	g_launches++;
	if (t->gemm) {
		// depth-1 blocks have no d plan: FFTW's REDFT10 of length 1 doubles its sample
		if (dsp_block_dct2d('f', t->work, t->work, t->D, t->H, t->W, t->bw, DSP_DCT_REDFT10, t->bd == 1 ? 2.0 : 1.0, stream)) return 1;
	}
	for (dsp_dct_plan p : t->fwd)
		if (dsp_dct_execute_dev(p, t->work, t->work, stream)) return 1;                              // :641 for every block
	const double q = t->quant != 0.0 ? (double)(float)(t->quant * 8.0 * std::sqrt(vol)) : 0.0;    // :570
	if (coeffs_coded && !rt_zero(t->d_count, sizeof(unsigned long long), st, g_err)) return 1;
	if (t->gemm && block_dquant_supports(t->bd)) {          // d forward + coefficient stage + d inverse in one pass (no d plans were made)
		if (dsp_block_dquant(t->work, t->D, t->H, t->W, t->bd, t->bh, t->bw, q, coeffs_coded ? t->d_count : nullptr, stream)) return 1;
	} else if (dsp_block_quant('f', t->work, t->D, t->H, t->W, t->bd, t->bh, t->bw, q, coeffs_coded ? t->d_count : nullptr, stream)) return 1;   // :644-647, :740-751
	for (dsp_dct_plan p : t->inv)
		if (dsp_dct_execute_dev(p, t->work, t->work, stream)) return 1;                              // :753
	if (t->gemm && dsp_block_dct2d('f', t->work, t->work, t->D, t->H, t->W, t->bw, DSP_DCT_REDFT01, 1.0, stream)) return 1;
	const double norm = 1.0 / std::sqrt(vol * 8.0);
	// (a fused 8-bit store in the GEMM kernel's epilogue was tried: the double-precision clamp / lround on the four store
	// warps made them the bottleneck, 7.6 -> 9.0 ms for the 256 x 1080 x 1920 volume; the separate sweep stays)
	if (dsp_block_store_u8('f', t->work, d_out, n, norm * norm, stream)) return 1;               // :757-776 (sf = 1)
	if (coeffs_coded) {
		unsigned long long c = 0;
		if (!rt_d2h(&c, t->d_count, sizeof c, st, g_err) || !rt_sync(st, g_err)) return 1;
		if (t->quant != 0.0) *coeffs_coded += c;
	}
	return 0;
}

int dsp_block_dct2d_debug(const void *d_in, void *d_out, long long nplanes, int H, int W, int B, int kind, double scale, void *stream, float *d_debug) {
	g_err.clear();
	if (!d_in || !d_out || nplanes < 1 || H < 1 || W < 1) { g_err = "block DCT: bad arguments"; return 1; }
	if (!rt_init(g_err)) return 1;
	const bool ok = launch_block_mm_f32((const float *)d_in, (float *)d_out, nplanes, H, W, B, kind, scale, (rt_stream)stream, g_err, d_debug);
	if (ok) g_launches++;
	return ok ? 0 : 1;
}

int dsp_block_dct2d(char prec, const void *d_in, void *d_out, long long nplanes, int H, int W, int B, int kind, double scale, void *stream) {
	if (prec != 'f') { g_err = "block DCT by GEMM: float only (use the per-axis plans for double)"; return 1; }
	return dsp_block_dct2d_debug(d_in, d_out, nplanes, H, W, B, kind, scale, stream, nullptr);
}

int dsp_dct_fuse_spec(dsp_dct_plan p, const dsp_spec_params *sp) {
	g_err.clear();
	if (!p || !sp) { g_err = "null plan or params"; return 1; }
	if (p->rank != 2 || p->kind[0] != DSP_DCT_REDFT10 || p->kind[1] != DSP_DCT_REDFT10) { g_err = "spec fusion needs a rank-2 REDFT10 plan"; return 1; }
	if (p->d > 4) { g_err = "spec fusion supports at most 4 channels"; return 1; }
	if (p->fuse_kind) { g_err = "plan already carries a fused stage"; return 1; }
	if (!ensure_scalars(p)) return 1;
	const int h = p->n[0], w = p->n[1];
	PassPlan &rowp = p->passes.front(), &colp = p->passes.back();
	OpAny op;
	memset(&op, 0, sizeof(op));
	op.kind = OP_SPEC;
	op.scaletype = sp->scaletype; op.signtype = sp->signtype; op.rangetype = sp->rangetype;
	op.d = p->d; op.w = w; op.h = h;
	op.p[0] = sp->gain; op.p[1] = 2.0 * (double)w * (double)h;
	op.q[0] = 1.0 / op.p[1]; op.q[1] = 254.0 / 255.0;          // reciprocals for the per-coefficient chain
	op.aux_c = p->d_scalars + 4;
	op.aux = p->d_scalars + 8;
	colp.sop = op; colp.fused = true;
	if (sp->rangetype != DSP_SPEC_RANGE_ONE) {
		memset(&rowp.sop, 0, sizeof(OpAny));
		rowp.sop.kind = OP_ACCUM_DC;
		rowp.sop.aux = p->d_scalars;
		rowp.fused = true;
		p->need_acc = true;
	}
	p->fuse_kind = 1;
	return 0;
}

int dsp_dct_spec_dc(dsp_dct_plan p, double *dc, int d) {
	g_err.clear();
	if (!p || p->fuse_kind != 1 || !dc || d < 1 || d > 4) { g_err = "plan has no spec stage"; return 1; }
	if (!rt_sync(p->last_stream, g_err)) return 1;
	if (!rt_d2h(dc, p->d_scalars + 8, sizeof(double) * (size_t)d, 0, g_err)) return 1;
	return rt_sync(0, g_err) ? 0 : 1;
}

int dsp_dct_fuse_ispec(dsp_dct_plan p, const dsp_ispec_params *ip) {
	g_err.clear();
	if (!p || !ip) { g_err = "null plan or params"; return 1; }
	if (p->rank != 2 || p->kind[0] != DSP_DCT_REDFT01 || p->kind[1] != DSP_DCT_REDFT01) { g_err = "ispec fusion needs a rank-2 REDFT01 plan"; return 1; }
	if (p->d > 4) { g_err = "ispec fusion supports at most 4 channels"; return 1; }
	if (p->fuse_kind) { g_err = "plan already carries a fused stage"; return 1; }
	const int h = p->n[0], w = p->n[1];
	OpAny op;
	memset(&op, 0, sizeof(op));
	op.kind = OP_ISPEC;
	op.scaletype = ip->scaletype; op.signtype = ip->signtype;
	op.d = p->d; op.w = w; op.h = h;
	op.p[0] = ip->gain; op.p[2] = 1.0 / ip->gain; op.p[3] = 255.0 / 254.0;   // reciprocals for the per-coefficient chain
	op.flag = ip->preserve_dc;
	for (int z = 0; z < 4; z++) {
		// spec/ispec.c:138: max[z] = log1p(max[z]) stored back into coeff precision
		double m = ip->max[z];
		if (p->prec == 'f') m = (double)(float)m;
		if (ip->scaletype == DSP_SPEC_SCALE_LOG) { m = log1p(m); if (p->prec == 'f') m = (double)(float)m; }
		op.q[z] = m;
		op.dc[z] = ip->dc[z];
	}
	if (ip->signmap) {
		const size_t bytes = (size_t)h * w * p->d;
		if (rt_is_device_ptr(ip->signmap)) op.aux_c = ip->signmap;
		else {
			if (!rt_malloc((void **)&p->d_signmap, bytes, g_err)) return 1;
			if (!rt_h2d(p->d_signmap, ip->signmap, bytes, 0, g_err) || !rt_sync(0, g_err)) return 1;
			op.aux_c = p->d_signmap;
		}
	}
	PassPlan &f = p->passes.front();
	f.lop = op; f.fused = true;
	p->fuse_kind = 2;
	return 0;
}

// ------------------------------------------------------------------------------------------------ scan session
struct dsp_scan_s {
	char prec;
	int h, w, d;
	size_t bytes;
	void *d_coeffs, *d_image, *d_sum;
	int32_t *d_index;
	dsp_dct_plan inverse;
};

static void scan_free(dsp_scan_s *s) {
	if (!s) return;
	if (s->inverse) dsp_dct_destroy(s->inverse);
	rt_free(s->d_coeffs); rt_free(s->d_image); rt_free(s->d_sum); rt_free(s->d_index);
	delete s;
}

dsp_scan dsp_scan_create(char prec, int h, int w, int d, const void *pixels, const int32_t *index_map) {
	g_err.clear();
	if ((prec != 'f' && prec != 'd') || h < 1 || w < 1 || d < 1 || d > 4 || !pixels || !index_map) { g_err = "bad scan arguments"; return nullptr; }
	dsp_scan_s *s = new dsp_scan_s();
	memset(s, 0, sizeof(*s));
	s->prec = prec; s->h = h; s->w = w; s->d = d;
	const size_t es = prec == 'f' ? 4 : 8, l = (size_t)h * w * d;
	s->bytes = l * es;
	const int n[2] = {h, w}, k10[2] = {DSP_DCT_REDFT10, DSP_DCT_REDFT10}, k01[2] = {DSP_DCT_REDFT01, DSP_DCT_REDFT01};
	const int hm = d, st = d > 1 ? d : 1, di = d > 1 ? 1 : 0;
	bool ok = rt_init(g_err) && rt_malloc(&s->d_coeffs, s->bytes, g_err) && rt_malloc(&s->d_image, s->bytes, g_err) &&
	          rt_malloc(&s->d_sum, s->bytes, g_err) && rt_malloc((void **)&s->d_index, sizeof(int32_t) * (size_t)h * w, g_err) &&
	          rt_h2d(s->d_coeffs, pixels, s->bytes, 0, g_err) && rt_h2d(s->d_index, index_map, sizeof(int32_t) * (size_t)h * w, 0, g_err);
	dsp_dct_plan fwd = nullptr;
	if (ok) {
		// scan.c:292-298: forward transform, coefficients normalised to the non-uniform range -1..1
		fwd = dsp_dct_plan_many(prec, 2, n, hm, s->d_coeffs, nullptr, st, di, s->d_coeffs, nullptr, st, di, k10, 0);
		ok = fwd && dsp_dct_fuse_scale(fwd, 1.0, 1.0 / (4.0 * (double)w * (double)h)) == 0 &&
		     dsp_dct_execute_dev(fwd, s->d_coeffs, s->d_coeffs, nullptr) == 0;
	}
	if (fwd) { rt_sync(0, g_err); dsp_dct_destroy(fwd); }
	if (ok) {
		// scan.c:381-383: the sum starts as the DC value of every channel
		std::vector<unsigned char> dc(es * (size_t)d), fill(s->bytes);
		ok = rt_d2h(dc.data(), s->d_coeffs, dc.size(), 0, g_err) && rt_sync(0, g_err);
		for (size_t i = 0; ok && i < (size_t)h * w; i++) memcpy(fill.data() + i * dc.size(), dc.data(), dc.size());
		ok = ok && rt_h2d(s->d_sum, fill.data(), s->bytes, 0, g_err) && rt_sync(0, g_err);
	}
	if (ok) {
		// scan.c:359: out-of-place inverse plan reconstruction -> image
		s->inverse = dsp_dct_plan_many(prec, 2, n, hm, s->d_coeffs, nullptr, st, di, s->d_image, nullptr, st, di, k01, 0);
		ok = s->inverse != nullptr;
	}
	if (ok) {
		PassPlan &f = s->inverse->passes.front(), &l2 = s->inverse->passes.back();
		memset(&f.lop, 0, sizeof(OpAny));
		f.lop.kind = OP_SCAN_MASK; f.lop.w = w; f.lop.h = h; f.lop.d = d; f.lop.aux_c = s->d_index; f.lop.lo = 0; f.lop.hi = 0;
		f.fused = true;
		memset(&l2.sop, 0, sizeof(OpAny));
		l2.sop.kind = OP_SCAN_ACCUM; l2.sop.w = w; l2.sop.h = h; l2.sop.d = d; l2.sop.aux = s->d_sum;
		l2.fused = true;
		s->inverse->fuse_kind = 3;
	}
	if (!ok) { scan_free(s); return nullptr; }
	return s;
}

int dsp_scan_frame(dsp_scan s, int lo, int hi, void *frame) {
	g_err.clear();
	if (!s) { g_err = "null scan"; return 1; }
	PassPlan &f = s->inverse->passes.front();
	f.lop.lo = lo; f.lop.hi = hi;
	if (dsp_dct_execute_dev(s->inverse, s->d_coeffs, s->d_image, nullptr) != 0) return 1;
	if (frame && !rt_d2h(frame, s->d_image, s->bytes, 0, g_err)) return 1;
	return rt_sync(0, g_err) ? 0 : 1;
}

int dsp_scan_coeffs(dsp_scan s, void *coeffs) {
	g_err.clear();
	if (!s || !coeffs) { g_err = "null scan"; return 1; }
	return (rt_d2h(coeffs, s->d_coeffs, s->bytes, 0, g_err) && rt_sync(0, g_err)) ? 0 : 1;
}

int dsp_scan_sum(dsp_scan s, void *sum) {
	g_err.clear();
	if (!s || !sum) { g_err = "null scan"; return 1; }
	return (rt_d2h(sum, s->d_sum, s->bytes, 0, g_err) && rt_sync(0, g_err)) ? 0 : 1;
}

void dsp_scan_destroy(dsp_scan s) { scan_free(s); }

// ------------------------------------------------------------------------------------------------ motion stages
// The three fused stages of motion's block body as plan-level calls, so that the slab-sharded volume (dist3d) can put
// them on its own plans: pel load on the first forward pass, coefficient stages on the first inverse pass, pel store
// on the last inverse pass.
int dsp_dct_fuse_pel_load(dsp_dct_plan p, int float_pixels) {
	g_err.clear();
	if (!p) { g_err = "null plan"; return 1; }
	PassPlan &f0 = p->passes.front();
	if (!f0.row) { g_err = "pel load needs a plan whose first pass runs along the contiguous axis"; return 1; }
	if (float_pixels) {
		if (p->prec != 'f') { g_err = "float pels need the float build (COEFF_PRECISION=F)"; return 1; }
		f0.lop.kind = OP_SCALE; f0.lop.p[0] = 255.0;                                        // motion.c:622
	} else f0.ra.in_u8 = 1;                                                                  // motion.c:624
	return 0;
}

static void motion_constants(const dsp_motion_params *mp, double &scalefactor, double &norm) {
	const double sw = mp->scaled[2], sh = mp->scaled[1], sd = mp->scaled[0];
	const double bw = mp->block[2], bh = mp->block[1], bd = mp->block[0];
	scalefactor = (sw * sh * sd) / (bw * bh * bd);                                           // motion.c:566
	norm = 1.0 / sqrt(sw * sh * sd * 8.0);                                                   // motion.c:567
}

// the coefficient stage of motion's block loop (motion.c:617, 644-751) as one pointwise map
static bool motion_coeff_op(const dsp_motion_params *mp, char prec, unsigned long long *d_counter, int flat_w, long long flat_base, OpAny &op) {
	double scalefactor, norm;
	motion_constants(mp, scalefactor, norm);
	const double sw = mp->scaled[2], sh = mp->scaled[1], sd = mp->scaled[0];
	memset(&op, 0, sizeof(op));
	op.kind = OP_MOTION_COEFF;
	for (int i = 0; i < 3; i++) {
		const int active = mp->block[i] < mp->scaled[i] ? mp->block[i] : mp->scaled[i];      // motion.c:494-496
		if (mp->bp_begin[i] < 0 || mp->bp_end[i] > active || mp->bp_begin[i] > mp->bp_end[i]) { g_err = "band-pass box outside the active box"; return false; }
		op.a3[i] = active; op.b3[i] = mp->bp_begin[i]; op.e3[i] = mp->bp_end[i];
	}
	op.m[0] = mp->damp; op.m[1] = mp->boost;
	op.m[2] = mp->threshold_min * 255.0 / norm / norm; op.m[3] = mp->threshold_max * 255.0 / norm / norm;   // motion.c:571-572
	if (prec == 'f') { op.m[2] = (double)(float)op.m[2]; op.m[3] = (double)(float)op.m[3]; }
	op.m[4] = mp->quant * 8.0 * sqrt(sw * sh * sd);                                          // motion.c:570
	if (prec == 'f') op.m[4] = (double)(float)op.m[4];
	op.m[5] = 127.5 / (norm * norm * scalefactor);                                           // motion.c:736
	op.flag = mp->preserve_dc;
	op.aux = d_counter;
	op.skipn = mp->ispec != 0; op.skipd = mp->spec != 0;
	op.fast = (mp->threshold_max == 0.0 && mp->preserve_dc == 0 && mp->quant == 0.0 && !op.skipd) ? 1 : 0;
	for (int k = 0; k < 4; k++) {
		const double s2 = 1.41421356237309504880168872420969808;
		double prod = 1.0;
		for (int q = 0; q < k; q++) prod *= s2;                                              // the reference multiplies the three factors
		op.nf[k] = (2 * s2) / prod;                                                          // motion.c:647
		op.rnf[k] = 1.0 / op.nf[k];
	}
	if (getenv("DSP_DCT_NO_FAST_OPS")) op.fast = 0;
	// flat_w > 0: the plan is the temporal pass of a slab-sharded volume over a [D][hw-slice] array -- the axis index is
	// z and (y, x) come from the flattened column index: hw = flat_base + column, y = hw / flat_w, x = hw % flat_w
	op.w = flat_w; op.lo = (int)flat_base;
	return true;
}

int dsp_dct_fuse_motion_coeff(dsp_dct_plan p, const dsp_motion_params *mp, unsigned long long *d_counter, int flat_w,
                              long long flat_base) {
	g_err.clear();
	if (!p || !mp) { g_err = "null plan or params"; return 1; }
	if (p->fuse_kind && p->fuse_kind != 4) { g_err = "plan already carries a fused stage"; return 1; }
	OpAny op;
	if (!motion_coeff_op(mp, p->prec, d_counter, flat_w, flat_base, op)) return 1;
	if (flat_w > 0 && (p->rank != 1 || p->passes.size() != 1 || p->passes[0].row)) {
		g_err = "flat coefficient coordinates need the rank-1 strided-axis plan of a [D][h*w slice] array";
		return 1;
	}
	if (p->kind[0] == DSP_DCT_REDFT10) {
		// on a forward plan the stages ride in the last pass's store (same pointwise map, applied one step earlier)
		PassPlan &l = p->passes.back();
		l.sop = op; l.fused = true;
	} else {
		PassPlan &i0 = p->passes.front();
		i0.lop = op; i0.fused = true;
	}
	p->fuse_kind = 4;
	return 0;
}

int dsp_motion_coeff_stage(char prec, const dsp_motion_params *mp, void *d_coeffs, unsigned long long *d_counter, void *stream) {
	g_err.clear();
	if ((prec != 'f' && prec != 'd') || !mp || !d_coeffs) { g_err = "coefficient stage: bad arguments"; return 1; }
	if (!rt_init(g_err)) return 1;
	OpAny op;
	if (!motion_coeff_op(mp, prec, d_counter, 0, 0, op)) return 1;
	int mb[3];
	for (int i = 0; i < 3; i++) mb[i] = mp->block[i] > mp->scaled[i] ? mp->block[i] : mp->scaled[i];      // minbuf, motion.c:491-493
	const bool ok = launch_motion_coeff(prec, op, d_coeffs, mb[0], mb[1], mb[2], (rt_stream)stream, g_err);
	if (ok) g_launches++;
	return ok ? 0 : 1;
}

int dsp_motion_coeff_stage_flat(char prec, const dsp_motion_params *mp, void *d_coeffs, int D, long long ncols, int flat_w, long long flat_base,
                                unsigned long long *d_counter, void *stream) {
	g_err.clear();
	if ((prec != 'f' && prec != 'd') || !mp || !d_coeffs || D < 1 || ncols < 1 || ncols > 0x7fffffffLL || flat_w < 1 || flat_base < 0) { g_err = "coefficient stage: bad arguments"; return 1; }
	if (!rt_init(g_err)) return 1;
	OpAny op;
	if (!motion_coeff_op(mp, prec, d_counter, flat_w, flat_base, op)) return 1;
	const bool ok = launch_motion_coeff(prec, op, d_coeffs, D, 0, (int)ncols, (rt_stream)stream, g_err);
	if (ok) g_launches++;
	return ok ? 0 : 1;
}

int dsp_dct_fuse_pel_store(dsp_dct_plan p, const dsp_motion_params *mp) {
	g_err.clear();
	if (!p || !mp) { g_err = "null plan or params"; return 1; }
	if (p->fuse_kind && p->fuse_kind != 4) { g_err = "plan already carries a fused stage"; return 1; }
	PassPlan &il = p->passes.back();
	if (!il.row) { g_err = "pel store needs a plan whose last pass runs along the contiguous axis"; return 1; }
	if (mp->float_pixels && p->prec != 'f') { g_err = "float pels need the float build (COEFF_PRECISION=F)"; return 1; }
	double scalefactor, norm;
	motion_constants(mp, scalefactor, norm);
	OpAny op;
	memset(&op, 0, sizeof(op));
	op.kind = OP_MOTION_STORE;
	op.m[6] = scalefactor * norm * norm;                                                     // motion.c:757,767
	op.flag2 = mp->float_pixels;
	il.sop = op; il.fused = true;
	if (!mp->float_pixels) {
		il.ra.out_u8 = 1;
		// the passes before the 8-bit store keep their T-typed intermediate in a work buffer of the output's extent
		if (p->passes.size() > 1 && !p->d_work && !rt_malloc(&p->d_work, p->out_span * (size_t)p->es, g_err)) return 1;
	}
	p->fuse_kind = 4;
	return 0;
}

// ------------------------------------------------------------------------------------------------ motion session
struct dsp_motion_s {
	char prec;
	dsp_motion_params mp;
	int minbuf[3];
	size_t pels, pel_bytes, coeff_bytes;
	dsp_dct_plan fwd, inv;
	void *d_coeffs;                 // T [minbuf]: zero outside the block box for the life of the session
	void *d_in, *d_out;             // staging for the host entry
	unsigned long long *d_counter;
	MotionSpecArgs sp_in, sp_out;   // --ispec / --spec stages
};

static void motion_free(dsp_motion_s *m) {
	if (!m) return;
	if (m->fwd) dsp_dct_destroy(m->fwd);
	if (m->inv) dsp_dct_destroy(m->inv);
	rt_free(m->d_coeffs); rt_free(m->d_in); rt_free(m->d_out); rt_free(m->d_counter);
	delete m;
}

dsp_motion dsp_motion_create(char prec, const dsp_motion_params *mp) {
	g_err.clear();
	if ((prec != 'f' && prec != 'd') || !mp) { g_err = "bad motion arguments"; return nullptr; }
	for (int i = 0; i < 3; i++)
		if (mp->block[i] < 1 || mp->scaled[i] < 1) { g_err = "motion block / scaled sizes must be >= 1"; return nullptr; }
	dsp_motion_s *m = new dsp_motion_s();
	memset(m, 0, sizeof(*m));
	m->prec = prec; m->mp = *mp;
	int active[3];
	for (int i = 0; i < 3; i++) {
		m->minbuf[i] = mp->block[i] > mp->scaled[i] ? mp->block[i] : mp->scaled[i];       // motion.c:491-493
		active[i] = mp->block[i] < mp->scaled[i] ? mp->block[i] : mp->scaled[i];          // motion.c:494-496
		if (mp->bp_begin[i] < 0 || mp->bp_end[i] > active[i] || mp->bp_begin[i] > mp->bp_end[i]) { g_err = "band-pass box outside the active box"; delete m; return nullptr; }
	}
	const size_t es = prec == 'f' ? 4 : 8;
	m->pels = (size_t)m->minbuf[0] * m->minbuf[1] * m->minbuf[2];
	m->pel_bytes = m->pels * (mp->float_pixels ? 4 : 1);
	m->coeff_bytes = m->pels * es;
	if (mp->float_pixels && prec != 'f') { g_err = "float pels need the float build (COEFF_PRECISION=F)"; delete m; return nullptr; }
	const int k10[3] = {DSP_DCT_REDFT10, DSP_DCT_REDFT10, DSP_DCT_REDFT10}, k01[3] = {DSP_DCT_REDFT01, DSP_DCT_REDFT01, DSP_DCT_REDFT01};
	bool ok = rt_init(g_err) && rt_malloc(&m->d_coeffs, m->coeff_bytes, g_err) && rt_zero(m->d_coeffs, m->coeff_bytes, 0, g_err) &&
	          rt_malloc((void **)&m->d_counter, sizeof(unsigned long long), g_err) && rt_zero(m->d_counter, sizeof(unsigned long long), 0, g_err) &&
	          rt_sync(0, g_err);
	if (ok && (mp->spec < 0 || mp->spec > 4 || mp->ispec < 0 || mp->ispec > 4 || mp->ispec == DSP_MOTION_SPEC_ABS)) { g_err = "bad motion spectrogram type"; ok = false; }
	if (ok) {
		// motion.c:535-538 / :549-552: both plans address their box inside the same minbuf-sized buffer
		// (a plan that --ispec / --spec replaces is still built: its coefficient stage descriptor is taken from it)
		m->fwd = dsp_dct_plan_many(prec, 3, mp->block, 1, nullptr, m->minbuf, 1, 0, nullptr, m->minbuf, 1, 0, k10, 0);
		m->inv = m->fwd ? dsp_dct_plan_many(prec, 3, mp->scaled, 1, nullptr, m->minbuf, 1, 0, nullptr, m->minbuf, 1, 0, k01, 0) : nullptr;
		ok = m->fwd && m->inv;
	}
	if (ok)
		ok = dsp_dct_fuse_pel_load(m->fwd, mp->float_pixels) == 0 &&
		     dsp_dct_fuse_motion_coeff(m->inv, mp, m->d_counter, 0, 0) == 0 && dsp_dct_fuse_pel_store(m->inv, mp) == 0;
	if (ok && (mp->spec || mp->ispec)) {
		double sf, norm;
		motion_constants(mp, sf, norm);
		MotionSpecArgs a;
		memset(&a, 0, sizeof(a));
		a.md = m->minbuf[0]; a.mh = m->minbuf[1]; a.mw = m->minbuf[2];
		a.float_pixels = mp->float_pixels;
		a.norm = norm; a.sf = sf;
		a.c = 127.5 / log1p((double)mp->scaled[2] * mp->scaled[1] * mp->scaled[0] * norm * 255 * 8);      // motion.c:568-569
		a.coeff = m->inv->passes.front().lop;
		m->sp_in = a; m->sp_in.type = mp->ispec; m->sp_in.bd = mp->block[0]; m->sp_in.bh = mp->block[1]; m->sp_in.bw = mp->block[2];
		m->sp_out = a; m->sp_out.type = mp->spec; m->sp_out.bd = mp->scaled[0]; m->sp_out.bh = mp->scaled[1]; m->sp_out.bw = mp->scaled[2];
	}
	if (!ok) { motion_free(m); return nullptr; }
	return m;
}

int dsp_motion_block_dev(dsp_motion m, const void *d_pels_in, void *d_pels_out, void *stream) {
	g_err.clear();
	if (!m || !d_pels_in || !d_pels_out) { g_err = "null motion session or buffer"; return 1; }
	// forward: pels -> coefficients (block box of the zero-initialised buffer); inverse: coefficients -> pels
	if (m->mp.ispec) {
		if (!launch_motion_ispec(m->prec, m->sp_in, d_pels_in, m->d_coeffs, (rt_stream)stream, g_err)) return 1;
		g_launches++;
	} else if (dsp_dct_execute_dev(m->fwd, (void *)d_pels_in, m->d_coeffs, stream) != 0) return 1;
	if (m->mp.spec) {
		if (!launch_motion_spec(m->prec, m->sp_out, m->d_coeffs, d_pels_out, (rt_stream)stream, g_err)) return 1;
		g_launches++;
		return 0;
	}
	return dsp_dct_execute_dev(m->inv, m->d_coeffs, d_pels_out, stream);
}

int dsp_motion_block(dsp_motion m, const void *pels_in, void *pels_out, unsigned long long *coeffs_coded) {
	g_err.clear();
	if (!m || !pels_in || !pels_out) { g_err = "null motion session or buffer"; return 1; }
	if (!m->d_in && !(rt_malloc(&m->d_in, m->pel_bytes, g_err) && rt_malloc(&m->d_out, m->pel_bytes, g_err))) return 1;
	if (!rt_h2d(m->d_in, pels_in, m->pel_bytes, 0, g_err)) return 1;
	// the reference writes the result over the staging block, leaving pels outside the scaled box as they were
	if (!rt_d2d(m->d_out, m->d_in, m->pel_bytes, 0, g_err)) return 1;
	if (!rt_zero(m->d_counter, sizeof(unsigned long long), 0, g_err)) return 1;
	if (dsp_motion_block_dev(m, m->d_in, m->d_out, nullptr) != 0) return 1;
	if (!rt_d2h(pels_out, m->d_out, m->pel_bytes, 0, g_err)) return 1;
	unsigned long long cnt = 0;
	if (!rt_d2h(&cnt, m->d_counter, sizeof(cnt), 0, g_err) || !rt_sync(0, g_err)) return 1;
	if (coeffs_coded) *coeffs_coded += cnt;
	return 0;
}

void dsp_motion_destroy(dsp_motion m) { motion_free(m); }

// ------------------------------------------------------------------------------------------------ zoom session
struct dsp_zoom_s {
	char prec;
	int h, w;
	size_t es;
	void *d_coeffs;                        // [h][w][3] REDFT10 x REDFT10 of the pixels
	void *d_xb, *d_yb, *d_tmp, *d_out, *d_pad, *d_cpl;
	size_t xb_bytes, yb_bytes, tmp_bytes, out_bytes, pad_bytes, cpl_bytes;
	int last_path;
};

static bool zoom_reserve(void **p, size_t *have, size_t need) {
	if (*have >= need) return true;
	rt_free(*p);
	*p = nullptr; *have = 0;
	if (!rt_malloc(p, need, g_err)) return false;
	*have = need;
	return true;
}

static void zoom_free(dsp_zoom_s *z) {
	if (!z) return;
	rt_free(z->d_coeffs); rt_free(z->d_xb); rt_free(z->d_yb); rt_free(z->d_tmp); rt_free(z->d_out); rt_free(z->d_pad); rt_free(z->d_cpl);
	delete z;
}

// zoom.c:37-41: scales below one sample clamp to 1/len; ncomponents = min(len, round(len * scale))
static void zoom_axis(int len, double &num, double &den, int &ncomp) {
	if (len * num / den < 1) { num = 1; den = len; }
	const double r = round(len * num / den);
	ncomp = (int)(r < len ? r : len);
	if (ncomp < 1) ncomp = 1;
}

dsp_zoom dsp_zoom_create(char prec, int h, int w, const void *pixels) {
	g_err.clear();
	if ((prec != 'f' && prec != 'd') || h < 1 || w < 1 || !pixels) { g_err = "bad zoom arguments"; return nullptr; }
	dsp_zoom_s *z = new dsp_zoom_s();
	memset(z, 0, sizeof(*z));
	z->prec = prec; z->h = h; z->w = w; z->es = prec == 'f' ? 4 : 8;
	const size_t bytes = (size_t)h * w * 3 * z->es;
	const int n[2] = {h, w}, k10[2] = {DSP_DCT_REDFT10, DSP_DCT_REDFT10};
	bool ok = rt_init(g_err) && rt_malloc(&z->d_coeffs, bytes, g_err) && rt_h2d(z->d_coeffs, pixels, bytes, 0, g_err);
	dsp_dct_plan fwd = ok ? dsp_dct_plan_many(prec, 2, n, 3, z->d_coeffs, nullptr, 3, 1, z->d_coeffs, nullptr, 3, 1, k10, 0) : nullptr;   // zoom.c:263
	ok = ok && fwd && dsp_dct_execute_dev(fwd, z->d_coeffs, z->d_coeffs, nullptr) == 0 && rt_sync(0, g_err);
	if (fwd) dsp_dct_destroy(fwd);
	if (!ok) { zoom_free(z); return nullptr; }
	return z;
}

int dsp_zoom_view_size(dsp_zoom z, const dsp_zoom_params *zp, int *vw, int *vh) {
	if (!z || !zp || !vw || !vh) return 1;
	double xn = zp->xscale_num, xd = zp->xscale_den, yn = zp->yscale_num, yd = zp->yscale_den;
	int cw, ch;
	zoom_axis(z->w, xn, xd, cw); zoom_axis(z->h, yn, yd, ch);
	*vw = zp->vw ? zp->vw : (int)(z->w * xn / xd);                                       // zoom.c:286-289
	*vh = zp->vh ? zp->vh : (int)(z->h * yn / yd);
	return 0;
}

int dsp_zoom_last_path(dsp_zoom z) { return z ? z->last_path : -1; }

int dsp_zoom_frame(dsp_zoom z, const dsp_zoom_params *zp, void *out) {
	g_err.clear();
	if (!z || !zp || !out) { g_err = "null zoom session or buffer"; return 1; }
	double xn = zp->xscale_num, xd = zp->xscale_den, yn = zp->yscale_num, yd = zp->yscale_den;
	int cw, ch, vw, vh;
	zoom_axis(z->w, xn, xd, cw); zoom_axis(z->h, yn, yd, ch);
	dsp_zoom_view_size(z, zp, &vw, &vh);
	if (vw < 1 || vh < 1) { g_err = "empty zoom view"; return 1; }
	const int W = z->w, H = z->h;
	const size_t es = z->es;
	const double inv_wh = 1.0 / ((double)W * (double)H);                                 // zoom.c:373
	if (!zoom_reserve(&z->d_out, &z->out_bytes, (size_t)vh * vw * 3 * es)) return 1;

	// ---- fast path: native basis, integer scaled size, whole image, no offset == zero-padded / cropped REDFT01
	const double sw = W * xn / xd, sh = H * yn / yd;
	const bool native_fft = zp->basis == 2 && zp->vx == 0 && zp->vy == 0 && sw == floor(sw) && sh == floor(sh) &&
	                        vw == (int)sw && vh == (int)sh;
	if (native_fft) {
		// out[j][i] = (C00/4 + ...)/(WH) = REDFT01^2 (zero-padded or cropped C) / (4 W H)
		const int pw = vw > W ? vw : W, ph = vh > H ? vh : H;                            // padded box holding both sub-boxes
		const size_t pbytes = (size_t)ph * pw * 3 * es;
		if (!zoom_reserve(&z->d_pad, &z->pad_bytes, pbytes)) return 1;
		if (!rt_zero(z->d_pad, pbytes, 0, g_err)) return 1;
		// copy the coefficient sub-box that survives (crop) into the padded box, row by row
		const int cwid = W < vw ? W : vw, chei = H < vh ? H : vh;
		for (int y = 0; y < chei; y++)
			if (!rt_d2d((char *)z->d_pad + (size_t)y * pw * 3 * es, (char *)z->d_coeffs + (size_t)y * W * 3 * es, (size_t)cwid * 3 * es, 0, g_err)) return 1;
		const int n[2] = {vh, vw}, emb[2] = {ph, pw}, k01[2] = {DSP_DCT_REDFT01, DSP_DCT_REDFT01};
		dsp_dct_plan inv = dsp_dct_plan_many(z->prec, 2, n, 3, z->d_pad, emb, 3, 1, z->d_out, n, 3, 1, k01, 0);
		if (!inv) return 1;
		bool ok = dsp_dct_fuse_scale(inv, 1.0, inv_wh / 4.0) == 0 && dsp_dct_execute_dev(inv, z->d_pad, z->d_out, nullptr) == 0 &&
		          rt_d2h(out, z->d_out, (size_t)vh * vw * 3 * es, 0, g_err) && rt_sync(0, g_err);
		dsp_dct_destroy(inv);
		z->last_path = 1;
		return ok ? 0 : 1;
	}

	// ---- fast path: interpolated basis with an integer scaled size == four phase-shifted REDFT01 x REDFT01 (kern_zoom.cu)
	const bool interp_fft = zp->basis == 0 && sw == floor(sw) && sh == floor(sh) && sw >= 1 && sh >= 1 && vw <= (int)sw &&
	                        vh <= (int)sh && cw <= (int)sw && ch <= (int)sh && !getenv("DSP_ZOOM_NO_SHIFT");
	if (interp_fft) {
		const int Nw = (int)sw, Nh = (int)sh;
		const double dx = zp->vx + (xn / xd - 1.0) / 2.0, dy = zp->vy + (yn / yd - 1.0) / 2.0;
		const size_t pbytes = 4 * (size_t)Nh * Nw * 3 * es;
		if (!zoom_reserve(&z->d_pad, &z->pad_bytes, pbytes)) return 1;
		if (!rt_zero(z->d_pad, pbytes, 0, g_err)) return 1;
		if (!launch_zoom_shift_build(z->prec, z->d_coeffs, z->d_pad, W, Nh, Nw, ch, cw, dx, dy, 0, g_err)) return 1;
		const int n[2] = {Nh, Nw}, k01[2] = {DSP_DCT_REDFT01, DSP_DCT_REDFT01};
		dsp_dct_plan inv = dsp_dct_plan_many_batched(z->prec, 2, n, 3, z->d_pad, nullptr, 3, 1, z->d_pad, nullptr, 3, 1, k01, 0, 4,
		                                             (ptrdiff_t)Nh * Nw * 3, (ptrdiff_t)Nh * Nw * 3);
		if (!inv) return 1;
		bool ok = dsp_dct_execute_dev(inv, z->d_pad, z->d_pad, nullptr) == 0 &&
		          launch_zoom_shift_combine(z->prec, z->d_pad, z->d_out, Nh, Nw, vh, vw, inv_wh / 4.0, 0, g_err) &&
		          rt_d2h(out, z->d_out, (size_t)vh * vw * 3 * es, 0, g_err) && rt_sync(0, g_err);
		dsp_dct_destroy(inv);
		g_launches += 2;
		z->last_path = 2;
		return ok ? 0 : 1;
	}

	// ---- general path on the tensor cores (float): the two contractions per channel as 3 x TF32 GEMMs (kern_gemm_tc.cu):
	//   tmpT[i][row] = sum_u xb[i][u] C[row][u][c]          (zoom.c:363-367, transposed so that it is K-major for the next product)
	//   out[j][i][c] = sum_v yb[j][v] tmpT[i][v] / (W H)     (zoom.c:368-374)
	if (z->prec == 'f' && gemm_tc_available()) {
		const int cwp = (cw + 3) & ~3, chp = (ch + 3) & ~3;                  // leading dimensions: multiples of 16 bytes for the TMA
		const size_t nxb = (size_t)vw * cwp, nyb = (size_t)vh * chp, ntm = (size_t)vw * chp, ncp = (size_t)ch * cwp;
		if (!zoom_reserve(&z->d_xb, &z->xb_bytes, 2 * nxb * 4) || !zoom_reserve(&z->d_yb, &z->yb_bytes, 2 * nyb * 4) ||
		    !zoom_reserve(&z->d_tmp, &z->tmp_bytes, 2 * ntm * 4) || !zoom_reserve(&z->d_cpl, &z->cpl_bytes, 2 * ncp * 4))
			return 1;
		float *xb = (float *)z->d_xb, *yb = (float *)z->d_yb, *tm = (float *)z->d_tmp, *cp = (float *)z->d_cpl;
		if (!launch_zoom_basis('f', xb, vw, cwp, zp->basis, xn, xd, zp->vx, W, 0, g_err) || !launch_tf32_residual(xb, xb + nxb, (long long)nxb, 0, g_err) ||
		    !launch_zoom_basis('f', yb, vh, chp, zp->basis, yn, yd, zp->vy, H, 0, g_err) || !launch_tf32_residual(yb, yb + nyb, (long long)nyb, 0, g_err))
			return 1;
		g_launches += 4;
		for (int c = 0; c < 3; c++) {
			if (!launch_planarize3((const float *)z->d_coeffs, ch, W, cw, c, cp, cp + ncp, cwp, 0, g_err) ||
			    !launch_gemm_tf32x3(vw, ch, cw, xb, xb + nxb, cwp, cp, cp + ncp, cwp, tm, tm + ntm, chp, 1, 1.0, 0, g_err) ||
			    !launch_gemm_tf32x3(vh, vw, ch, yb, yb + nyb, chp, tm, tm + ntm, chp, (float *)z->d_out + c, nullptr, (long long)vw * 3, 3, inv_wh, 0, g_err))
				return 1;
			g_launches += 3;
		}
		z->last_path = 3;
		return (rt_d2h(out, z->d_out, (size_t)vh * vw * 3 * es, 0, g_err) && rt_sync(0, g_err)) ? 0 : 1;
	}

	// ---- general path: scaled bases (zoom.c:347-358) and the separable synthesis (zoom.c:361-375)
	if (!zoom_reserve(&z->d_xb, &z->xb_bytes, (size_t)vw * cw * es) || !zoom_reserve(&z->d_yb, &z->yb_bytes, (size_t)vh * ch * es) ||
	    !zoom_reserve(&z->d_tmp, &z->tmp_bytes, (size_t)ch * vw * es))
		return 1;
	if (!launch_zoom_basis(z->prec, z->d_xb, vw, cw, zp->basis, xn, xd, zp->vx, W, 0, g_err)) return 1;
	if (!launch_zoom_basis(z->prec, z->d_yb, vh, ch, zp->basis, yn, yd, zp->vy, H, 0, g_err)) return 1;
	g_launches += 2;
	for (int c = 0; c < 3; c++) {
		// tmp[row][i] = sum_u C[row][u][c] * xb[i][u]      (zoom.c:363-367; xb[i][0] = 1/2)
		if (!launch_zoom_gemm(z->prec, ch, vw, cw, (char *)z->d_coeffs + c * es, (long long)W * 3, 3, z->d_xb, 1, cw, z->d_tmp, vw, 1, 1.0, 0, g_err)) return 1;
		// out[j][i][c] = sum_v yb[j][v] * tmp[v][i] / (W H)   (zoom.c:368-374; yb[j][0] = 1/2)
		if (!launch_zoom_gemm(z->prec, vh, vw, ch, z->d_yb, ch, 1, z->d_tmp, vw, 1, (char *)z->d_out + c * es, (long long)vw * 3, 3, inv_wh, 0, g_err)) return 1;
		g_launches += 2;
	}
	z->last_path = 0;
	return (rt_d2h(out, z->d_out, (size_t)vh * vw * 3 * es, 0, g_err) && rt_sync(0, g_err)) ? 0 : 1;
}

void dsp_zoom_destroy(dsp_zoom z) { zoom_free(z); }

}  // extern "C"
