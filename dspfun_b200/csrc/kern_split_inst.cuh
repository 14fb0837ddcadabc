// kern_split_inst.cuh -- the sub-passes of the split column pass (dct_split.cuh), one element type per translation
// unit.  The including .cu defines KERN_T (KERN_T|double) and KERN_SUFFIX (f32|f64).  The element type is a template
// parameter of every kernel so that the two units instantiate distinctly named kernels.
#include "dsp_kernels.h"
#define KS_CAT_(a, b) a##b
#define KS_CAT(a, b) KS_CAT_(a, b)
#define KS_NAME(base) KS_CAT(base, KERN_SUFFIX)
#include "dct_split.cuh"
#include <vector>
#include <cstdint>
#include <cstdlib>

#ifndef DSP_SPLIT_MINB
#define DSP_SPLIT_MINB 3      // M = 256 sub-pass A kernels: 3 CTAs per SM (35 KB tiles, <= 85 registers).  Measured: n = 4096
                              // column pass 0.345 -> 0.314 ms; at M = 512 (512 CTAs per panel = 1.15 waves of 444) it loses.
#endif

namespace dsp {

// LGM: log2 of the sub-FFT length M = n/16 when fixed at compile time (FastFixed: every smem offset and loop bound of
// sub-pass A folds), 0 = runtime length.  The thread count is the constant kThreads for the same reason.
template <int LGM> struct SubDesc {
	typedef FastFixed<(LGM ? LGM : 8)> type;
	DSP_DEVM static type make(const FastDesc &f) { type r; r.tw = f.tw; r.om = f.om; r.sig = f.sig; return r; }
};
template <int LGM, bool FWD, class L, class S>
DSP_DEV void split_fft_body(const SplitArgs &a, const FastDesc &fM, const L &l, const S &s, int cta, int t0, int t1, int nthr, C2<KERN_T> *smem) {
	if (LGM) cta_split_fft<KERN_T, FWD, L, S>(a, SubDesc<LGM>::make(fM), l, s, cta, t0, t1, nthr, smem);
	else cta_split_fft<KERN_T, FWD, L, S>(a, fM, l, s, cta, t0, t1, nthr, smem);
}
template <int LGM, class L>
DSP_DEV void split_inv_fft_body(const SplitArgs &a, const FastDesc &fM, const FastDesc &fN, const L &l, int cta, int t0, int t1, int nthr, C2<KERN_T> *smem) {
	if (LGM) cta_split_inv_fft<KERN_T, L>(a, SubDesc<LGM>::make(fM), fN, l, cta, t0, t1, nthr, smem);
	else cta_split_inv_fft<KERN_T, L>(a, fM, fN, l, cta, t0, t1, nthr, smem);
}

#if DSP_GPU
template <class TT, int LGM, bool FWD, class L, class S>
__global__ void __launch_bounds__(kThreads, KERN_IS_F32 ? ((LGM == 8) ? DSP_SPLIT_MINB : 2) : 1)
k_split_fft(const __grid_constant__ SplitArgs a, const __grid_constant__ FastDesc fM, const __grid_constant__ L l,
            const __grid_constant__ S s) {
	extern __shared__ __align__(16) unsigned char smem[];
	split_fft_body<LGM, FWD, L, S>(a, fM, l, s, (int)blockIdx.x, (int)threadIdx.x, (int)threadIdx.x + 1, kThreads, (C2<KERN_T> *)smem);
}
template <class TT, bool FWD, class L, class S, bool LEAN>
__global__ void __launch_bounds__(kThreads, KERN_IS_F32 ? 2 : 1)
k_split_outer(const __grid_constant__ SplitArgs a, const __grid_constant__ FastDesc fN, const __grid_constant__ L l,
              const __grid_constant__ S s) {
	split_outer_thread<KERN_T, FWD, L, S, LEAN>(a, fN, l, s, (int)blockIdx.x * (kThreads / 32) + (int)threadIdx.x / 32, (int)threadIdx.x % 32);
}
#endif

#if DSP_GPU
template <class TT, int LGM, class L>
__global__ void __launch_bounds__(kThreads, KERN_IS_F32 ? ((LGM == 8) ? DSP_SPLIT_MINB : 2) : 1)
k_split_inv_fft(const __grid_constant__ SplitArgs a, const __grid_constant__ FastDesc fM, const __grid_constant__ FastDesc fN,
                const __grid_constant__ L l) {
	extern __shared__ __align__(16) unsigned char smem[];
	split_inv_fft_body<LGM, L>(a, fM, fN, l, (int)blockIdx.x, (int)threadIdx.x, (int)threadIdx.x + 1, kThreads, (C2<KERN_T> *)smem);
}
template <class TT, class S, bool LEAN>
__global__ void __launch_bounds__(kThreads, KERN_IS_F32 ? 3 : 1)
k_split_inv_outer(const __grid_constant__ SplitArgs a, const __grid_constant__ FastDesc fN, const __grid_constant__ S s) {
	split_inv_outer_thread<KERN_T, S, LEAN>(a, fN, s, (int)blockIdx.x * (kThreads / 32) + (int)threadIdx.x / 32, (int)threadIdx.x % 32);
}
#endif

// image side of sub-pass B: whole column pairs, 8-byte aligned rows, < 2^31 elements (GlobalCols<LEAN>)
static bool outer_lean(const SplitArgs &a, const void *img, long long rs) {
	return KERN_IS_F32 && (a.pcol0 % 2) == 0 && (a.pcols % 2) == 0 && (rs % 2) == 0 && ((uintptr_t)img % 8) == 0 &&
	       (long long)a.n * rs < (1ll << 31) && !getenv("DSP_DCT_NO_FIXED");
}

// DIT-style inverse (lean only: full 16-column tiles, plain scale ops)
template <int LGM>
static bool split_inv_fft_t(const SplitArgs &a, const FastDesc &fM, const FastDesc &fN, const OpMul<KERN_T> &lm, int grid, size_t smem,
                            rt_stream st, std::string &err) {
#if DSP_GPU
	static unsigned long long attr_dev = 0;      // one bit per device: the attribute is per (function, device)
	const int dev = rt_device() & 63;
	if (smem > 48 * 1024 && !((attr_dev >> dev) & 1ull)) {
		if (!rt_ok(cudaFuncSetAttribute(k_split_inv_fft<KERN_T, LGM, OpMul<KERN_T>>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem), err, "smem attribute")) return false;
		attr_dev |= 1ull << dev;
	}
	k_split_inv_fft<KERN_T, LGM, OpMul<KERN_T>><<<grid, kThreads, smem, st>>>(a, fM, fN, lm);
	return rt_ok(cudaGetLastError(), err, "split inverse fft launch");
#else
	(void)st; (void)err;
	std::vector<unsigned char> buf(smem + 64);
	for (int cta = 0; cta < grid; cta++) split_inv_fft_body<LGM, OpMul<KERN_T>>(a, fM, fN, lm, cta, 0, kThreads, kThreads, (C2<KERN_T> *)buf.data());
	return true;
#endif
}

bool KS_NAME(launch_split_inv_fft_)(const SplitArgs &a, const FastDesc &fM, const FastDesc &fN, const OpAny &lop, int grid, size_t smem,
                              rt_stream st, std::string &err) {
	const OpMul<KERN_T> lm = {(KERN_T)(lop.kind == OP_SCALE ? lop.p[0] : 1.0)};
#if KERN_IS_F32                       /* FastFixed pads like float */
	if (!getenv("DSP_DCT_NO_FIXED")) {
		if (fM.n == 256) return split_inv_fft_t<8>(a, fM, fN, lm, grid, smem, st, err);
		if (fM.n == 512) return split_inv_fft_t<9>(a, fM, fN, lm, grid, smem, st, err);
		if (fM.n == 1024) return split_inv_fft_t<10>(a, fM, fN, lm, grid, smem, st, err);
	}
#endif
	return split_inv_fft_t<0>(a, fM, fN, lm, grid, smem, st, err);
}

bool KS_NAME(launch_split_inv_outer_)(const SplitArgs &a, const FastDesc &fN, const OpAny &sop, int nwarps, rt_stream st, std::string &err) {
	const OpMul<KERN_T> sm = {(KERN_T)(sop.kind == OP_SCALE ? sop.p[0] : 1.0)};
	const bool lean = outer_lean(a, a.out, a.ax_os);
#if DSP_GPU
	const int wpb = kThreads / 32;
	if (lean) k_split_inv_outer<KERN_T, OpMul<KERN_T>, true><<<(nwarps + wpb - 1) / wpb, kThreads, 0, st>>>(a, fN, sm);
	else k_split_inv_outer<KERN_T, OpMul<KERN_T>, false><<<(nwarps + wpb - 1) / wpb, kThreads, 0, st>>>(a, fN, sm);
	return rt_ok(cudaGetLastError(), err, "split inverse outer launch");
#else
	(void)st; (void)err;
	for (int w = 0; w < nwarps; w++)
		for (int lane = 0; lane < 32; lane++) {
			if (lean) split_inv_outer_thread<KERN_T, OpMul<KERN_T>, true>(a, fN, sm, w, lane);
			else split_inv_outer_thread<KERN_T, OpMul<KERN_T>, false>(a, fN, sm, w, lane);
		}
	return true;
#endif
}

template <int LGM, bool FWD, class L, class S>
static bool split_fft_t(const SplitArgs &a, const FastDesc &fM, const L &l, const S &s, int grid, size_t smem, rt_stream st, std::string &err) {
#if DSP_GPU
	static unsigned long long attr_dev = 0;      // one bit per device: the attribute is per (function, device)
	const int dev = rt_device() & 63;
	if (smem > 48 * 1024 && !((attr_dev >> dev) & 1ull)) {
		if (!rt_ok(cudaFuncSetAttribute(k_split_fft<KERN_T, LGM, FWD, L, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem), err, "smem attribute")) return false;
		attr_dev |= 1ull << dev;
	}
	k_split_fft<KERN_T, LGM, FWD, L, S><<<grid, kThreads, smem, st>>>(a, fM, l, s);
	return rt_ok(cudaGetLastError(), err, "split fft launch");
#else
	(void)st; (void)err;
	std::vector<unsigned char> buf(smem + 64);
	for (int cta = 0; cta < grid; cta++) split_fft_body<LGM, FWD, L, S>(a, fM, l, s, cta, 0, kThreads, kThreads, (C2<KERN_T> *)buf.data());
	return true;
#endif
}

template <bool FWD, class L, class S, bool LEAN>
static bool split_outer_t(const SplitArgs &a, const FastDesc &fN, const L &l, const S &s, int nwarps, rt_stream st, std::string &err) {
#if DSP_GPU
	const int wpb = kThreads / 32;
	k_split_outer<KERN_T, FWD, L, S, LEAN><<<(nwarps + wpb - 1) / wpb, kThreads, 0, st>>>(a, fN, l, s);
	return rt_ok(cudaGetLastError(), err, "split outer launch");
#else
	(void)st; (void)err;
	for (int w = 0; w < nwarps; w++)
		for (int lane = 0; lane < 32; lane++) split_outer_thread<KERN_T, FWD, L, S, LEAN>(a, fN, l, s, w, lane);
	return true;
#endif
}

bool KS_NAME(launch_split_fft_)(const SplitArgs &a, const FastDesc &fM, bool fused, const OpAny &lop, const OpAny &sop, int grid, size_t smem,
                          rt_stream st, std::string &err) {
	const bool fwd = a.kind == DSP_KIND_REDFT10;
	const OpMul<KERN_T> lm = {(KERN_T)(lop.kind == OP_SCALE ? lop.p[0] : 1.0)}, sm = {(KERN_T)(sop.kind == OP_SCALE ? sop.p[0] : 1.0)};
#if KERN_IS_F32
	if (fwd && !fused && !getenv("DSP_DCT_NO_FIXED")) {
		if (fM.n == 256) return split_fft_t<8, true, OpMul<KERN_T>, OpMul<KERN_T>>(a, fM, lm, sm, grid, smem, st, err);
		if (fM.n == 512) return split_fft_t<9, true, OpMul<KERN_T>, OpMul<KERN_T>>(a, fM, lm, sm, grid, smem, st, err);
		if (fM.n == 1024) return split_fft_t<10, true, OpMul<KERN_T>, OpMul<KERN_T>>(a, fM, lm, sm, grid, smem, st, err);
	}
#endif
	if (fwd) return fused ? split_fft_t<0, true, OpAny, OpAny>(a, fM, lop, sop, grid, smem, st, err)
	                      : split_fft_t<0, true, OpMul<KERN_T>, OpMul<KERN_T>>(a, fM, lm, sm, grid, smem, st, err);
	return fused ? split_fft_t<0, false, OpAny, OpAny>(a, fM, lop, sop, grid, smem, st, err)
	             : split_fft_t<0, false, OpMul<KERN_T>, OpMul<KERN_T>>(a, fM, lm, sm, grid, smem, st, err);
}

bool KS_NAME(launch_split_outer_)(const SplitArgs &a, const FastDesc &fN, bool fused, const OpAny &lop, const OpAny &sop, int nwarps,
                            rt_stream st, std::string &err) {
	const bool fwd = a.kind == DSP_KIND_REDFT10;
	const OpMul<KERN_T> lm = {(KERN_T)(lop.kind == OP_SCALE ? lop.p[0] : 1.0)}, sm = {(KERN_T)(sop.kind == OP_SCALE ? sop.p[0] : 1.0)};
	if (fwd && !fused && outer_lean(a, a.out, a.ax_os))
		return split_outer_t<true, OpMul<KERN_T>, OpMul<KERN_T>, true>(a, fN, lm, sm, nwarps, st, err);
	if (fwd) return fused ? split_outer_t<true, OpAny, OpAny, false>(a, fN, lop, sop, nwarps, st, err)
	                      : split_outer_t<true, OpMul<KERN_T>, OpMul<KERN_T>, false>(a, fN, lm, sm, nwarps, st, err);
	return fused ? split_outer_t<false, OpAny, OpAny, false>(a, fN, lop, sop, nwarps, st, err)
	             : split_outer_t<false, OpMul<KERN_T>, OpMul<KERN_T>, false>(a, fN, lm, sm, nwarps, st, err);
}

}  // namespace dsp
