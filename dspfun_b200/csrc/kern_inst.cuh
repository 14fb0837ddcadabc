// kern_inst.cuh -- one pass-kernel family per translation unit.  The including .cu defines
//   KERN_T (float|double), KERN_SUFFIX (f32|f64), KERN_ROW (1|0), KERN_FAST (1|0)
#include "dsp_kernels.h"
#include <vector>
#include <cstdlib>

#ifndef KERN_GENERIC_MINB
// generic (mixed-radix) float kernels are issue-latency bound (ncu r02: ~50 % issue utilisation at 16 warps per SM, top stall
// "wait"), so the row kernels are built for 4 CTAs of 256 threads per SM (64 registers, ~100-300 B of spills): 1920-point
// rows of the 256x1080x1920 volume 3.25 -> 2.37 ms forward, 3.67 -> 2.58 ms inverse (3 CTAs: 2.63 / 2.92, 5: 2.34 / 2.54).
// Column kernels stay at 2: their 16-column tiles take 74 KB of shared memory, and neither a register cap (2.21 -> 2.40 ms
// at n = 1080) nor 8-column tiles with 4 CTAs (2.39 ms) paid.
#define KERN_GENERIC_MINB (KERN_ROW ? 4 : 2)
#endif
#ifndef KERN_ROW_MINB
#define KERN_ROW_MINB 2
#endif
#ifndef KERN_DD
#define KERN_DD 1             // interleave the KERN_PLANAR translation unit is specialised for (1 planar, 3 RGB)
#endif

namespace dsp {

#define KERN_CAT_(a, b, c, d) a##b##c##d
#define KERN_CAT(a, b, c, d) KERN_CAT_(a, b, c, d)
#if KERN_ROW
#define KERN_ARGS RowArgs
#define KERN_RC _row_
#else
#define KERN_ARGS ColArgs
#define KERN_RC _col_
#endif
#if KERN_FAST
#define KERN_GF fast_
#else
#define KERN_GF generic_
#endif
#define KERN_LAUNCH KERN_CAT(launch, KERN_RC, KERN_GF, KERN_SUFFIX)

// NB: the element type and the row/fast flags are template parameters so that every translation unit
// instantiates distinctly named kernels (same-signature templates in different TUs would be merged by the linker).
// FAST: bit 0 = fast path, bit 1 = forward (fast kernels are specialised on the transform kind),
// bit 2 = layout-specialised row kernels (their own TUs, KERN_PLANAR), bits 4..6 = their interleave DD (KERN_DD),
// bits 8.. = log2 n when the length is fixed at compile time (FastFixed), 0 = runtime length
template <class TT, int ROW, int FAST, class L, class S>
DSP_DEV void cta_body(const KERN_ARGS &a, const FastDesc &f, const L &l, const S &s, int cta, int t0, int t1, int nthr,
                      C2<KERN_T> *smem) {
#if KERN_FAST
	if ((FAST >> 8) != 0) {
		FastFixed<((FAST >> 8) != 0) ? (FAST >> 8) : 8> ff;
		ff.tw = f.tw; ff.om = f.om; ff.sig = f.sig;
#if KERN_ROW
		typedef FastFixed<((FAST >> 8) != 0) ? (FAST >> 8) : 8> FF;
		if (FAST & 4) cta_row_fast<KERN_T, (FAST & 2) != 0, L, S, FF, ((FAST >> 4) & 7)>(a, ff, l, s, cta, t0, t1, nthr, smem);
		else cta_row_fast<KERN_T, (FAST & 2) != 0, L, S>(a, ff, l, s, cta, t0, t1, nthr, smem);
#else
		cta_col_fast<KERN_T, (FAST & 2) != 0, L, S>(a, ff, l, s, cta, t0, t1, nthr, smem);
#endif
		return;
	}
#endif
#if KERN_ROW && KERN_FAST
	cta_row_fast<KERN_T, (FAST & 2) != 0, L, S>(a, f, l, s, cta, t0, t1, nthr, smem);
#elif KERN_ROW
	(void)f;
	cta_row_pass<KERN_T, L, S>(a, l, s, cta, t0, t1, nthr, smem);
#elif KERN_FAST
	cta_col_fast<KERN_T, (FAST & 2) != 0, L, S>(a, f, l, s, cta, t0, t1, nthr, smem);
#else
	(void)f;
	cta_col_pass<KERN_T, L, S>(a, l, s, cta, t0, t1, nthr, smem);
#endif
}

#if DSP_GPU
// f32: two CTAs per SM (<= 128 registers); f64: one (the paired outer pass keeps 64 complex doubles live)
template <class TT, int ROW, int FAST, class L, class S>
__global__ void __launch_bounds__((KERN_FAST && !KERN_ROW) ? 2 * kThreads : kThreads, sizeof(KERN_T) != 4 ? 1 : (KERN_FAST ? (KERN_ROW ? KERN_ROW_MINB : 1) : KERN_GENERIC_MINB))
k_pass(const __grid_constant__ KERN_ARGS a, const __grid_constant__ FastDesc f, const __grid_constant__ L l,
       const __grid_constant__ S s) {
	extern __shared__ __align__(16) unsigned char smem[];
	cta_body<TT, ROW, FAST, L, S>(a, f, l, s, (int)blockIdx.x, (int)threadIdx.x, (int)threadIdx.x + 1, (int)blockDim.x, (C2<KERN_T> *)smem);
}
#endif

template <class TT, int ROW, int FAST, class L, class S>
static bool launch_t(const KERN_ARGS &a, const FastDesc &f, const L &l, const S &s, int grid, int block, size_t smem,
                     rt_stream st, std::string &err) {
#if DSP_GPU
	static unsigned long long attr_dev = 0;      // one bit per device: the attribute is per (function, device)
	const int dev = rt_device() & 63;
	if (smem > 48 * 1024 && !((attr_dev >> dev) & 1ull)) {
		if (!rt_ok(cudaFuncSetAttribute(k_pass<TT, ROW, FAST, L, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem), err, "smem attribute"))
			return false;
		attr_dev |= 1ull << dev;
	}
	k_pass<TT, ROW, FAST, L, S><<<grid, block, smem, st>>>(a, f, l, s);
	return rt_ok(cudaGetLastError(), err, "pass kernel launch");
#else
	(void)st; (void)err;
	std::vector<unsigned char> buf(smem + 64);
	for (int cta = 0; cta < grid; cta++) cta_body<TT, ROW, FAST, L, S>(a, f, l, s, cta, 0, block, block, (C2<KERN_T> *)buf.data());
	return true;
#endif
}

bool KERN_LAUNCH(const KERN_ARGS &a, const FastDesc &f, bool fused, const OpAny &lop, const OpAny &sop, int grid, int block,
                 size_t smem, rt_stream st, std::string &err) {
	if (!KERN_FAST) block = kThreads;
	// lean kernels: the only pointwise stage is a multiply (1 unless dsp_dct_fuse_scale set it)
	const OpMul<KERN_T> lm = {(KERN_T)(lop.kind == OP_SCALE ? lop.p[0] : 1.0)}, sm = {(KERN_T)(sop.kind == OP_SCALE ? sop.p[0] : 1.0)};
#ifdef KERN_PLANAR
	// planar row specialisation: the planner has checked the layout (PassPlan::planar); fixed lengths only
	{
		const bool fw = a.kind == DSP_KIND_REDFT10;
		(void)fused;
#define DSP_PLANAR_CASE(LG)                                                                                                  \
	case (1 << LG):                                                                                                          \
		return fw ? launch_t<KERN_T, KERN_ROW, 7 | (KERN_DD << 4) | (LG << 8), OpMul<KERN_T>, OpMul<KERN_T>>(a, f, lm, sm, grid, block, smem, st, err) \
		          : launch_t<KERN_T, KERN_ROW, 5 | (KERN_DD << 4) | (LG << 8), OpMul<KERN_T>, OpMul<KERN_T>>(a, f, lm, sm, grid, block, smem, st, err);
		switch (f.n) {
			DSP_PLANAR_CASE(8) DSP_PLANAR_CASE(9) DSP_PLANAR_CASE(10) DSP_PLANAR_CASE(11) DSP_PLANAR_CASE(12) DSP_PLANAR_CASE(13)
		default: break;
		}
#undef DSP_PLANAR_CASE
		err = "no planar row kernel for this length";
		return false;
	}
#else
	// motion's hot stages on their own small functors (dct_ops.cuh), float only
	if constexpr (sizeof(KERN_T) == 4) if (fused && !getenv("DSP_DCT_NO_FAST_OPS")) {
		const bool lplain = lop.kind == OP_NONE || lop.kind == OP_SCALE, splain = sop.kind == OP_NONE || sop.kind == OP_SCALE;
		const bool fw = a.kind == DSP_KIND_REDFT10;
#if KERN_FAST
#define DSP_OPS_LAUNCH(L, S, l, s) (fw ? launch_t<KERN_T, KERN_ROW, 3, L, S>(a, f, l, s, grid, block, smem, st, err) : launch_t<KERN_T, KERN_ROW, 1, L, S>(a, f, l, s, grid, block, smem, st, err))
#else
#define DSP_OPS_LAUNCH(L, S, l, s) launch_t<KERN_T, KERN_ROW, 0, L, S>(a, f, l, s, grid, block, smem, st, err)
#endif
		(void)fw;
#if !KERN_ROW
		if (lop.kind == OP_MOTION_COEFF && lop.fast && splain) return DSP_OPS_LAUNCH(OpMotionCoeff, OpMul<KERN_T>, OpMotionCoeff::from(lop), sm);
		if (sop.kind == OP_MOTION_COEFF && sop.fast && lplain) return DSP_OPS_LAUNCH(OpMul<KERN_T>, OpMotionCoeff, lm, OpMotionCoeff::from(sop));
		if (sop.kind == OP_SPEC && lplain) return DSP_OPS_LAUNCH(OpMul<KERN_T>, OpSpecStore, lm, OpSpecStore::from(sop));
		if (lop.kind == OP_ISPEC && splain) return DSP_OPS_LAUNCH(OpIspecLoad, OpMul<KERN_T>, OpIspecLoad::from(lop), sm);
#else
		if (sop.kind == OP_MOTION_STORE && lplain) return DSP_OPS_LAUNCH(OpMul<KERN_T>, OpMotionStore, lm, OpMotionStore::from(sop));
		if (sop.kind == OP_ACCUM_DC && lplain) return DSP_OPS_LAUNCH(OpMul<KERN_T>, OpAccumDc, lm, OpAccumDc::from(sop));
#endif
#undef DSP_OPS_LAUNCH
		(void)lplain; (void)splain;
	}
#if KERN_FAST
	// lean kernels with the length fixed at compile time for the common sizes (float only: FastFixed pads like float)
	if (!fused && sizeof(KERN_T) == 4 && !getenv("DSP_DCT_NO_FIXED")) {
		const bool fw = a.kind == DSP_KIND_REDFT10;
#define DSP_FIXED_CASE(LG)                                                                                                   \
	case (1 << LG):                                                                                                          \
		return fw ? launch_t<KERN_T, KERN_ROW, 3 | (LG << 8), OpMul<KERN_T>, OpMul<KERN_T>>(a, f, lm, sm, grid, block, smem, st, err) \
		          : launch_t<KERN_T, KERN_ROW, 1 | (LG << 8), OpMul<KERN_T>, OpMul<KERN_T>>(a, f, lm, sm, grid, block, smem, st, err);
		switch (f.n) {
			DSP_FIXED_CASE(8) DSP_FIXED_CASE(9) DSP_FIXED_CASE(10) DSP_FIXED_CASE(11) DSP_FIXED_CASE(12) DSP_FIXED_CASE(13)
		default: break;
		}
#undef DSP_FIXED_CASE
	}
	if (a.kind == DSP_KIND_REDFT10) {
		if (fused) return launch_t<KERN_T, KERN_ROW, 3, OpAny, OpAny>(a, f, lop, sop, grid, block, smem, st, err);
		return launch_t<KERN_T, KERN_ROW, 3, OpMul<KERN_T>, OpMul<KERN_T>>(a, f, lm, sm, grid, block, smem, st, err);
	}
#endif
	if (fused) return launch_t<KERN_T, KERN_ROW, KERN_FAST, OpAny, OpAny>(a, f, lop, sop, grid, block, smem, st, err);
	return launch_t<KERN_T, KERN_ROW, KERN_FAST, OpMul<KERN_T>, OpMul<KERN_T>>(a, f, lm, sm, grid, block, smem, st, err);
#endif
}

}  // namespace dsp
