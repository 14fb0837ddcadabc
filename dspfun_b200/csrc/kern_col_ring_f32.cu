// persistent TMA-fed sub-passes of the split column pass (dct_colring.cuh): float, n = 4096 | 8192
#include "dsp_kernels.h"
#include <vector>
#include "dct_colring.cuh"

namespace dsp {

static int ring_sm_count() {
#if DSP_GPU
	static int sm_count[64] = {0};
	const int dev = rt_device() & 63;
	if (!sm_count[dev]) {
		int v = 0;
		if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v < 1) v = 148;
		sm_count[dev] = v;
	}
	return sm_count[dev];
#else
	return 3;                                  // emulation: a few "SMs", so that CTAs take several iterations each
#endif
}

template <int LGM, bool INV>
static bool colring_launch_t(const ColRingArgs &a, rt_stream st, std::string &err) {
	const int Q = a.nplanes * a.ppp;
	int minseg = 1 << 30;
	for (int s = 0; s < 2 * Q; s++) { const int n = ColWork<LGM, INV>::seg_items(a, s, Q); if (n < minseg) minseg = n; }
	// The counters are free of cycles when every segment holds at least one item of every CTA and B(q) does not follow
	// A(q) directly (>= 2 panels): see colring_cta.
	if (Q < 2 || minseg < 1) { err = "ring column pass needs at least two panels"; return false; }
	const int sms = ring_sm_count();
	int grid = minseg;
	if (grid > sms) grid = sms;
#if DSP_GPU
	const size_t smem = ColRingSmem<LGM>::kTotal;
	static unsigned long long attr_dev = 0;      // one bit per device: the attribute is per (function, device)
	const int dev = rt_device() & 63;
	if (!((attr_dev >> dev) & 1ull)) {
		if (!rt_ok(cudaFuncSetAttribute(k_col_ring<LGM, INV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), err, "smem attribute")) return false;
		attr_dev |= 1ull << dev;
	}
	// every CTA must be resident (items wait on counters other CTAs bump): a cooperative launch makes the driver check it
	void *params[] = {(void *)&a};
	return rt_ok(cudaLaunchCooperativeKernel((const void *)k_col_ring<LGM, INV>, dim3(grid), dim3(kRingGroups * kRingGroupThreads), params, smem, st), err,
	             "ring column kernel launch");
#else
	(void)st; (void)err;
	colring_emulate<LGM, INV>(a, grid);
	return true;
#endif
}

bool colring_supports(int n) { return n == 4096 || n == 8192; }
int colring_max_panels() { return kColRingMaxPanels; }
int colring_scratch_panels() { return kColRingScratch; }

// tensor maps of one (input image, output image, scratch) triple; planes and panels enter through coordinates
bool colring_encode(ColRingArgs &a, int n, bool inverse, const float *in, long long ax_is, long long plane_is, float *out, long long ax_os,
                    long long plane_os, int nplanes, int ncols, float *scratch, int P, std::string &err) {
	const int M = n / 16;
	TmaView v;
	v.rank = 4;
	v.strides[0] = 4;
	if (inverse) {
		const unsigned SR = (unsigned)(M < 256 ? M : 256);
		// image as [planes][M][16][cols]: rows k = 16 k' + phase (sub-pass A'' load)
		v.base = (void *)in;
		v.dims[0] = (unsigned long long)ncols; v.dims[1] = 16; v.dims[2] = (unsigned long long)M; v.dims[3] = (unsigned long long)nplanes;
		v.strides[1] = (unsigned long long)ax_is * 4; v.strides[2] = 16ull * (unsigned long long)ax_is * 4;
		v.strides[3] = nplanes > 1 ? (unsigned long long)plane_is * 4 : v.strides[2] * v.dims[2];
		v.box[0] = 16; v.box[1] = 1; v.box[2] = SR; v.box[3] = 1;
		if (!tma_encode(&a.in_map, v, err)) return false;
		// scratch as [3][16][M][P]: A'' stores 16-column boxes, B'' loads {32 cols, 32 i, 16 j}
		v.base = scratch;
		v.dims[0] = (unsigned long long)P; v.dims[1] = (unsigned long long)M; v.dims[2] = 16; v.dims[3] = kColRingScratch;
		v.strides[1] = (unsigned long long)P * 4; v.strides[2] = (unsigned long long)M * (unsigned long long)P * 4; v.strides[3] = 16 * v.strides[2];
		v.box[0] = 16; v.box[1] = SR; v.box[2] = 1;
		if (!tma_encode(&a.sc_st_map, v, err)) return false;
		v.box[0] = 32; v.box[1] = 32; v.box[2] = 16;
		if (!tma_encode(&a.sc_ld_map, v, err)) return false;
		a.sc_ld1_map = a.sc_ld_map; a.out_map = a.sc_ld_map; a.out1_map = a.sc_ld_map;       // unused by the inverse
		return true;
	}
	// image as [planes][n/32][32][cols] (sub-pass A load)
	v.base = (void *)in;
	v.dims[0] = (unsigned long long)ncols; v.dims[1] = 32; v.dims[2] = (unsigned long long)(n / 32); v.dims[3] = (unsigned long long)nplanes;
	v.strides[1] = (unsigned long long)ax_is * 4; v.strides[2] = 32ull * (unsigned long long)ax_is * 4;
	v.strides[3] = nplanes > 1 ? (unsigned long long)plane_is * 4 : v.strides[2] * v.dims[2];
	v.box[0] = 32; v.box[1] = 1; v.box[2] = (unsigned)(M / 2); v.box[3] = 1;
	if (!tma_encode(&a.in_map, v, err)) return false;
	// scratch as [3][16][M][P]
	v.base = scratch;
	v.dims[0] = (unsigned long long)P; v.dims[1] = (unsigned long long)M; v.dims[2] = 16; v.dims[3] = kColRingScratch;
	v.strides[1] = (unsigned long long)P * 4; v.strides[2] = (unsigned long long)M * (unsigned long long)P * 4; v.strides[3] = 16 * v.strides[2];
	v.box[0] = 32; v.box[1] = (unsigned)(M < 256 ? M : 256); v.box[2] = 1;
	if (!tma_encode(&a.sc_st_map, v, err)) return false;
	v.box[1] = 16; v.box[2] = 16;
	if (!tma_encode(&a.sc_ld_map, v, err)) return false;
	v.box[1] = 1;
	if (!tma_encode(&a.sc_ld1_map, v, err)) return false;
	// image as [planes][16][M][cols] (sub-pass B store)
	v.base = out;
	v.dims[0] = (unsigned long long)ncols; v.dims[1] = (unsigned long long)M; v.dims[2] = 16; v.dims[3] = (unsigned long long)nplanes;
	v.strides[1] = (unsigned long long)ax_os * 4; v.strides[2] = (unsigned long long)M * (unsigned long long)ax_os * 4;
	v.strides[3] = nplanes > 1 ? (unsigned long long)plane_os * 4 : v.strides[2] * v.dims[2];
	v.box[1] = 16; v.box[2] = 16;
	if (!tma_encode(&a.out_map, v, err)) return false;
	v.box[1] = 1;
	return tma_encode(&a.out1_map, v, err);
}

bool launch_col_ring_f32(const ColRingArgs &a, int n, bool inverse, rt_stream st, std::string &err) {
	if (n == 8192) return inverse ? colring_launch_t<9, true>(a, st, err) : colring_launch_t<9, false>(a, st, err);
	if (n == 4096) return inverse ? colring_launch_t<8, true>(a, st, err) : colring_launch_t<8, false>(a, st, err);
	err = "no ring column kernel for this length";
	return false;
}

}  // namespace dsp
