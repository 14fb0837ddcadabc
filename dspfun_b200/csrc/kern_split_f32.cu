// kern_split_f32.cu -- the two sub-passes of the split column pass (dct_split.cuh), float.
#include "dsp_kernels.h"
#include "dct_split.cuh"
#include <vector>

namespace dsp {

#if DSP_GPU
template <bool FWD, class L, class S>
__global__ void __launch_bounds__(kThreads, 2)
k_split_fft(const __grid_constant__ SplitArgs a, const __grid_constant__ FastDesc fM, const __grid_constant__ L l,
            const __grid_constant__ S s) {
	extern __shared__ __align__(16) unsigned char smem[];
	cta_split_fft<float, FWD, L, S>(a, fM, l, s, (int)blockIdx.x, (int)threadIdx.x, (int)threadIdx.x + 1, (int)blockDim.x, (C2<float> *)smem);
}
template <bool FWD, class L, class S>
__global__ void __launch_bounds__(kThreads, 2)
k_split_outer(const __grid_constant__ SplitArgs a, const __grid_constant__ FastDesc fN, const __grid_constant__ L l,
              const __grid_constant__ S s) {
	split_outer_thread<float, FWD, L, S>(a, fN, l, s, (int)blockIdx.x * (kThreads / 32) + (int)threadIdx.x / 32, (int)threadIdx.x % 32);
}
#endif

#if DSP_GPU
template <class L>
__global__ void __launch_bounds__(kThreads, 2)
k_split_inv_fft(const __grid_constant__ SplitArgs a, const __grid_constant__ FastDesc fM, const __grid_constant__ FastDesc fN,
                const __grid_constant__ L l) {
	extern __shared__ __align__(16) unsigned char smem[];
	cta_split_inv_fft<float, L>(a, fM, fN, l, (int)blockIdx.x, (int)threadIdx.x, (int)threadIdx.x + 1, (int)blockDim.x, (C2<float> *)smem);
}
template <class S>
__global__ void __launch_bounds__(kThreads, 3)
k_split_inv_outer(const __grid_constant__ SplitArgs a, const __grid_constant__ FastDesc fN, const __grid_constant__ S s) {
	split_inv_outer_thread<float, S>(a, fN, s, (int)blockIdx.x * (kThreads / 32) + (int)threadIdx.x / 32, (int)threadIdx.x % 32);
}
#endif

// DIT-style inverse (lean only: full 16-column tiles, plain scale ops)
bool launch_split_inv_fft_f32(const SplitArgs &a, const FastDesc &fM, const FastDesc &fN, const OpAny &lop, int grid, size_t smem,
                              rt_stream st, std::string &err) {
	const OpMul<float> lm = {(float)(lop.kind == OP_SCALE ? lop.p[0] : 1.0)};
#if DSP_GPU
	static size_t attr_set = 0;
	if (smem > 48 * 1024 && smem > attr_set) {
		if (!rt_ok(cudaFuncSetAttribute(k_split_inv_fft<OpMul<float>>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem), err, "smem attribute")) return false;
		attr_set = kMaxSmem;
	}
	k_split_inv_fft<OpMul<float>><<<grid, kThreads, smem, st>>>(a, fM, fN, lm);
	return rt_ok(cudaGetLastError(), err, "split inverse fft launch");
#else
	(void)st; (void)err;
	std::vector<unsigned char> buf(smem + 64);
	for (int cta = 0; cta < grid; cta++) cta_split_inv_fft<float, OpMul<float>>(a, fM, fN, lm, cta, 0, kThreads, kThreads, (C2<float> *)buf.data());
	return true;
#endif
}

bool launch_split_inv_outer_f32(const SplitArgs &a, const FastDesc &fN, const OpAny &sop, int nwarps, rt_stream st, std::string &err) {
	const OpMul<float> sm = {(float)(sop.kind == OP_SCALE ? sop.p[0] : 1.0)};
#if DSP_GPU
	const int wpb = kThreads / 32;
	k_split_inv_outer<OpMul<float>><<<(nwarps + wpb - 1) / wpb, kThreads, 0, st>>>(a, fN, sm);
	return rt_ok(cudaGetLastError(), err, "split inverse outer launch");
#else
	(void)st; (void)err;
	for (int w = 0; w < nwarps; w++)
		for (int lane = 0; lane < 32; lane++) split_inv_outer_thread<float, OpMul<float>>(a, fN, sm, w, lane);
	return true;
#endif
}

template <bool FWD, class L, class S>
static bool split_fft_t(const SplitArgs &a, const FastDesc &fM, const L &l, const S &s, int grid, size_t smem, rt_stream st, std::string &err) {
#if DSP_GPU
	static size_t attr_set = 0;
	if (smem > 48 * 1024 && smem > attr_set) {
		if (!rt_ok(cudaFuncSetAttribute(k_split_fft<FWD, L, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem), err, "smem attribute")) return false;
		attr_set = kMaxSmem;
	}
	k_split_fft<FWD, L, S><<<grid, kThreads, smem, st>>>(a, fM, l, s);
	return rt_ok(cudaGetLastError(), err, "split fft launch");
#else
	(void)st; (void)err;
	std::vector<unsigned char> buf(smem + 64);
	for (int cta = 0; cta < grid; cta++) cta_split_fft<float, FWD, L, S>(a, fM, l, s, cta, 0, kThreads, kThreads, (C2<float> *)buf.data());
	return true;
#endif
}

template <bool FWD, class L, class S>
static bool split_outer_t(const SplitArgs &a, const FastDesc &fN, const L &l, const S &s, int nwarps, rt_stream st, std::string &err) {
#if DSP_GPU
	const int wpb = kThreads / 32;
	k_split_outer<FWD, L, S><<<(nwarps + wpb - 1) / wpb, kThreads, 0, st>>>(a, fN, l, s);
	return rt_ok(cudaGetLastError(), err, "split outer launch");
#else
	(void)st; (void)err;
	for (int w = 0; w < nwarps; w++)
		for (int lane = 0; lane < 32; lane++) split_outer_thread<float, FWD, L, S>(a, fN, l, s, w, lane);
	return true;
#endif
}

bool launch_split_fft_f32(const SplitArgs &a, const FastDesc &fM, bool fused, const OpAny &lop, const OpAny &sop, int grid, size_t smem,
                          rt_stream st, std::string &err) {
	const bool fwd = a.kind == DSP_KIND_REDFT10;
	const OpMul<float> lm = {(float)(lop.kind == OP_SCALE ? lop.p[0] : 1.0)}, sm = {(float)(sop.kind == OP_SCALE ? sop.p[0] : 1.0)};
	if (fwd) return fused ? split_fft_t<true, OpAny, OpAny>(a, fM, lop, sop, grid, smem, st, err)
	                      : split_fft_t<true, OpMul<float>, OpMul<float>>(a, fM, lm, sm, grid, smem, st, err);
	return fused ? split_fft_t<false, OpAny, OpAny>(a, fM, lop, sop, grid, smem, st, err)
	             : split_fft_t<false, OpMul<float>, OpMul<float>>(a, fM, lm, sm, grid, smem, st, err);
}

bool launch_split_outer_f32(const SplitArgs &a, const FastDesc &fN, bool fused, const OpAny &lop, const OpAny &sop, int nwarps,
                            rt_stream st, std::string &err) {
	const bool fwd = a.kind == DSP_KIND_REDFT10;
	const OpMul<float> lm = {(float)(lop.kind == OP_SCALE ? lop.p[0] : 1.0)}, sm = {(float)(sop.kind == OP_SCALE ? sop.p[0] : 1.0)};
	if (fwd) return fused ? split_outer_t<true, OpAny, OpAny>(a, fN, lop, sop, nwarps, st, err)
	                      : split_outer_t<true, OpMul<float>, OpMul<float>>(a, fN, lm, sm, nwarps, st, err);
	return fused ? split_outer_t<false, OpAny, OpAny>(a, fN, lop, sop, nwarps, st, err)
	             : split_outer_t<false, OpMul<float>, OpMul<float>>(a, fN, lm, sm, nwarps, st, err);
}

}  // namespace dsp
