// kern_split_f32.cu -- the two sub-passes of the split column pass (dct_split.cuh), float.
#include "dsp_kernels.h"
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
DSP_DEV void split_fft_body(const SplitArgs &a, const FastDesc &fM, const L &l, const S &s, int cta, int t0, int t1, int nthr, C2<float> *smem) {
	if (LGM) cta_split_fft<float, FWD, L, S>(a, SubDesc<LGM>::make(fM), l, s, cta, t0, t1, nthr, smem);
	else cta_split_fft<float, FWD, L, S>(a, fM, l, s, cta, t0, t1, nthr, smem);
}
template <int LGM, class L>
DSP_DEV void split_inv_fft_body(const SplitArgs &a, const FastDesc &fM, const FastDesc &fN, const L &l, int cta, int t0, int t1, int nthr, C2<float> *smem) {
	if (LGM) cta_split_inv_fft<float, L>(a, SubDesc<LGM>::make(fM), fN, l, cta, t0, t1, nthr, smem);
	else cta_split_inv_fft<float, L>(a, fM, fN, l, cta, t0, t1, nthr, smem);
}

#if DSP_GPU
template <int LGM, bool FWD, class L, class S>
__global__ void __launch_bounds__(kThreads, (LGM == 8) ? DSP_SPLIT_MINB : 2)
k_split_fft(const __grid_constant__ SplitArgs a, const __grid_constant__ FastDesc fM, const __grid_constant__ L l,
            const __grid_constant__ S s) {
	extern __shared__ __align__(16) unsigned char smem[];
	split_fft_body<LGM, FWD, L, S>(a, fM, l, s, (int)blockIdx.x, (int)threadIdx.x, (int)threadIdx.x + 1, kThreads, (C2<float> *)smem);
}
template <bool FWD, class L, class S, bool LEAN>
__global__ void __launch_bounds__(kThreads, 2)
k_split_outer(const __grid_constant__ SplitArgs a, const __grid_constant__ FastDesc fN, const __grid_constant__ L l,
              const __grid_constant__ S s) {
	split_outer_thread<float, FWD, L, S, LEAN>(a, fN, l, s, (int)blockIdx.x * (kThreads / 32) + (int)threadIdx.x / 32, (int)threadIdx.x % 32);
}
#endif

#if DSP_GPU
template <int LGM, class L>
__global__ void __launch_bounds__(kThreads, (LGM == 8) ? DSP_SPLIT_MINB : 2)
k_split_inv_fft(const __grid_constant__ SplitArgs a, const __grid_constant__ FastDesc fM, const __grid_constant__ FastDesc fN,
                const __grid_constant__ L l) {
	extern __shared__ __align__(16) unsigned char smem[];
	split_inv_fft_body<LGM, L>(a, fM, fN, l, (int)blockIdx.x, (int)threadIdx.x, (int)threadIdx.x + 1, kThreads, (C2<float> *)smem);
}
template <class S, bool LEAN>
__global__ void __launch_bounds__(kThreads, 3)
k_split_inv_outer(const __grid_constant__ SplitArgs a, const __grid_constant__ FastDesc fN, const __grid_constant__ S s) {
	split_inv_outer_thread<float, S, LEAN>(a, fN, s, (int)blockIdx.x * (kThreads / 32) + (int)threadIdx.x / 32, (int)threadIdx.x % 32);
}
#endif

// image side of sub-pass B: whole column pairs, 8-byte aligned rows, < 2^31 elements (GlobalCols<LEAN>)
static bool outer_lean(const SplitArgs &a, const void *img, long long rs) {
	return (a.pcol0 % 2) == 0 && (a.pcols % 2) == 0 && (rs % 2) == 0 && ((uintptr_t)img % 8) == 0 &&
	       (long long)a.n * rs < (1ll << 31) && !getenv("DSP_DCT_NO_FIXED");
}

// DIT-style inverse (lean only: full 16-column tiles, plain scale ops)
template <int LGM>
static bool split_inv_fft_t(const SplitArgs &a, const FastDesc &fM, const FastDesc &fN, const OpMul<float> &lm, int grid, size_t smem,
                            rt_stream st, std::string &err) {
#if DSP_GPU
	static size_t attr_set = 0;
	if (smem > 48 * 1024 && smem > attr_set) {
		if (!rt_ok(cudaFuncSetAttribute(k_split_inv_fft<LGM, OpMul<float>>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem), err, "smem attribute")) return false;
		attr_set = kMaxSmem;
	}
	k_split_inv_fft<LGM, OpMul<float>><<<grid, kThreads, smem, st>>>(a, fM, fN, lm);
	return rt_ok(cudaGetLastError(), err, "split inverse fft launch");
#else
	(void)st; (void)err;
	std::vector<unsigned char> buf(smem + 64);
	for (int cta = 0; cta < grid; cta++) split_inv_fft_body<LGM, OpMul<float>>(a, fM, fN, lm, cta, 0, kThreads, kThreads, (C2<float> *)buf.data());
	return true;
#endif
}

bool launch_split_inv_fft_f32(const SplitArgs &a, const FastDesc &fM, const FastDesc &fN, const OpAny &lop, int grid, size_t smem,
                              rt_stream st, std::string &err) {
	const OpMul<float> lm = {(float)(lop.kind == OP_SCALE ? lop.p[0] : 1.0)};
	if (!getenv("DSP_DCT_NO_FIXED")) {
		if (fM.n == 256) return split_inv_fft_t<8>(a, fM, fN, lm, grid, smem, st, err);
		if (fM.n == 512) return split_inv_fft_t<9>(a, fM, fN, lm, grid, smem, st, err);
		if (fM.n == 1024) return split_inv_fft_t<10>(a, fM, fN, lm, grid, smem, st, err);
	}
	return split_inv_fft_t<0>(a, fM, fN, lm, grid, smem, st, err);
}

bool launch_split_inv_outer_f32(const SplitArgs &a, const FastDesc &fN, const OpAny &sop, int nwarps, rt_stream st, std::string &err) {
	const OpMul<float> sm = {(float)(sop.kind == OP_SCALE ? sop.p[0] : 1.0)};
	const bool lean = outer_lean(a, a.out, a.ax_os);
#if DSP_GPU
	const int wpb = kThreads / 32;
	if (lean) k_split_inv_outer<OpMul<float>, true><<<(nwarps + wpb - 1) / wpb, kThreads, 0, st>>>(a, fN, sm);
	else k_split_inv_outer<OpMul<float>, false><<<(nwarps + wpb - 1) / wpb, kThreads, 0, st>>>(a, fN, sm);
	return rt_ok(cudaGetLastError(), err, "split inverse outer launch");
#else
	(void)st; (void)err;
	for (int w = 0; w < nwarps; w++)
		for (int lane = 0; lane < 32; lane++) {
			if (lean) split_inv_outer_thread<float, OpMul<float>, true>(a, fN, sm, w, lane);
			else split_inv_outer_thread<float, OpMul<float>, false>(a, fN, sm, w, lane);
		}
	return true;
#endif
}

template <int LGM, bool FWD, class L, class S>
static bool split_fft_t(const SplitArgs &a, const FastDesc &fM, const L &l, const S &s, int grid, size_t smem, rt_stream st, std::string &err) {
#if DSP_GPU
	static size_t attr_set = 0;
	if (smem > 48 * 1024 && smem > attr_set) {
		if (!rt_ok(cudaFuncSetAttribute(k_split_fft<LGM, FWD, L, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem), err, "smem attribute")) return false;
		attr_set = kMaxSmem;
	}
	k_split_fft<LGM, FWD, L, S><<<grid, kThreads, smem, st>>>(a, fM, l, s);
	return rt_ok(cudaGetLastError(), err, "split fft launch");
#else
	(void)st; (void)err;
	std::vector<unsigned char> buf(smem + 64);
	for (int cta = 0; cta < grid; cta++) split_fft_body<LGM, FWD, L, S>(a, fM, l, s, cta, 0, kThreads, kThreads, (C2<float> *)buf.data());
	return true;
#endif
}

template <bool FWD, class L, class S, bool LEAN>
static bool split_outer_t(const SplitArgs &a, const FastDesc &fN, const L &l, const S &s, int nwarps, rt_stream st, std::string &err) {
#if DSP_GPU
	const int wpb = kThreads / 32;
	k_split_outer<FWD, L, S, LEAN><<<(nwarps + wpb - 1) / wpb, kThreads, 0, st>>>(a, fN, l, s);
	return rt_ok(cudaGetLastError(), err, "split outer launch");
#else
	(void)st; (void)err;
	for (int w = 0; w < nwarps; w++)
		for (int lane = 0; lane < 32; lane++) split_outer_thread<float, FWD, L, S, LEAN>(a, fN, l, s, w, lane);
	return true;
#endif
}

bool launch_split_fft_f32(const SplitArgs &a, const FastDesc &fM, bool fused, const OpAny &lop, const OpAny &sop, int grid, size_t smem,
                          rt_stream st, std::string &err) {
	const bool fwd = a.kind == DSP_KIND_REDFT10;
	const OpMul<float> lm = {(float)(lop.kind == OP_SCALE ? lop.p[0] : 1.0)}, sm = {(float)(sop.kind == OP_SCALE ? sop.p[0] : 1.0)};
	if (fwd && !fused && !getenv("DSP_DCT_NO_FIXED")) {
		if (fM.n == 256) return split_fft_t<8, true, OpMul<float>, OpMul<float>>(a, fM, lm, sm, grid, smem, st, err);
		if (fM.n == 512) return split_fft_t<9, true, OpMul<float>, OpMul<float>>(a, fM, lm, sm, grid, smem, st, err);
		if (fM.n == 1024) return split_fft_t<10, true, OpMul<float>, OpMul<float>>(a, fM, lm, sm, grid, smem, st, err);
	}
	if (fwd) return fused ? split_fft_t<0, true, OpAny, OpAny>(a, fM, lop, sop, grid, smem, st, err)
	                      : split_fft_t<0, true, OpMul<float>, OpMul<float>>(a, fM, lm, sm, grid, smem, st, err);
	return fused ? split_fft_t<0, false, OpAny, OpAny>(a, fM, lop, sop, grid, smem, st, err)
	             : split_fft_t<0, false, OpMul<float>, OpMul<float>>(a, fM, lm, sm, grid, smem, st, err);
}

bool launch_split_outer_f32(const SplitArgs &a, const FastDesc &fN, bool fused, const OpAny &lop, const OpAny &sop, int nwarps,
                            rt_stream st, std::string &err) {
	const bool fwd = a.kind == DSP_KIND_REDFT10;
	const OpMul<float> lm = {(float)(lop.kind == OP_SCALE ? lop.p[0] : 1.0)}, sm = {(float)(sop.kind == OP_SCALE ? sop.p[0] : 1.0)};
	if (fwd && !fused && outer_lean(a, a.out, a.ax_os))
		return split_outer_t<true, OpMul<float>, OpMul<float>, true>(a, fN, lm, sm, nwarps, st, err);
	if (fwd) return fused ? split_outer_t<true, OpAny, OpAny, false>(a, fN, lop, sop, nwarps, st, err)
	                      : split_outer_t<true, OpMul<float>, OpMul<float>, false>(a, fN, lm, sm, nwarps, st, err);
	return fused ? split_outer_t<false, OpAny, OpAny, false>(a, fN, lop, sop, nwarps, st, err)
	             : split_outer_t<false, OpMul<float>, OpMul<float>, false>(a, fN, lm, sm, nwarps, st, err);
}

}  // namespace dsp
