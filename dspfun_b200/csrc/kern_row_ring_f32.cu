// persistent TMA-fed row pass (dct_ring.cuh): planar float lines, n = 256 .. 8192, DCT-II and DCT-III
#include "dsp_kernels.h"
#include <vector>
#include "dct_ring.cuh"

namespace dsp {

template <int LG, bool FWD>
static bool ring_launch_t(const RingArgs &a, int sms, rt_stream st, std::string &err) {
	typedef RingGeom<LG> G;
	const long long total_iters = ((long long)a.nlines / 2 + G::NSEQ - 1) / G::NSEQ;
	const int grid = (int)(total_iters < sms ? total_iters : sms);
#if DSP_GPU
	const size_t smem = RingSmem<LG>::kTotal;
	static unsigned long long attr_dev = 0;      // one bit per device: the attribute is per (function, device)
	const int dev = rt_device() & 63;
	if (!((attr_dev >> dev) & 1ull)) {
		if (!rt_ok(cudaFuncSetAttribute(k_row_ring<LG, FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), err, "smem attribute")) return false;
		attr_dev |= 1ull << dev;
	}
	k_row_ring<LG, FWD><<<grid, kRingGroups * kRingGroupThreads, smem, st>>>(a);
	return rt_ok(cudaGetLastError(), err, "ring row kernel launch");
#else
	(void)st; (void)err;
	ring_emulate<LG, FWD>(a, grid);
	return true;
#endif
}

// true when the ring kernel serves this length; the caller has checked the layout (planar float lines at one stride,
// 16-byte aligned, an even number of them)
bool ring_supports(int n) { return n >= 256 && n <= 8192 && (n & (n - 1)) == 0; }

bool launch_row_ring_f32(const RingArgs &a, int n, bool fwd, rt_stream st, std::string &err) {
	int sms = 148;
#if DSP_GPU
	static int sm_count[64] = {0};
	const int dev = rt_device() & 63;
	if (!sm_count[dev]) {
		int v = 0;
		if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v < 1) v = 148;
		sm_count[dev] = v;
	}
	sms = sm_count[dev];
#else
	sms = 3;                                   // emulation: a few "SMs", so that CTAs take several iterations each
#endif
#define DSP_RING_CASE(LG)                                                                         \
	case (1 << LG): return fwd ? ring_launch_t<LG, true>(a, sms, st, err) : ring_launch_t<LG, false>(a, sms, st, err);
	switch (n) {
		DSP_RING_CASE(8) DSP_RING_CASE(9) DSP_RING_CASE(10) DSP_RING_CASE(11) DSP_RING_CASE(12) DSP_RING_CASE(13)
	default: break;
	}
#undef DSP_RING_CASE
	err = "no ring row kernel for this length";
	return false;
}

}  // namespace dsp
