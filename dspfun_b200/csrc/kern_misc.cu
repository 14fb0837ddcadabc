// kern_misc.cu -- small helper kernels
#include "dsp_kernels.h"

namespace dsp {

#if DSP_GPU
template <class T> __global__ void k_spec_resolve(OpAny op, const double *acc, double *scale_z) {
	if (threadIdx.x == 0 && blockIdx.x == 0) spec_resolve_range<T>(op, acc, scale_z);
}
#endif

#if DSP_GPU
// pulls a [nrows][row_bytes] region (row pitch `pitch` bytes) into L2, one prefetch per 128-byte line; the kernel
// only issues the prefetches, the transfers complete while the following kernels of the stream run
__global__ void k_l2_prefetch(const char *base, long long pitch, int nrows, int lines_per_row) {
	const long long total = (long long)nrows * lines_per_row;
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
		const long long r = i / lines_per_row;
		const int l = (int)(i - r * lines_per_row);
		const char *p = base + r * pitch + (long long)l * 128;
		asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
	}
}
#endif

bool launch_l2_prefetch(const void *base, long long pitch_bytes, int nrows, int row_bytes, rt_stream st, std::string &err) {
#if DSP_GPU
	if (nrows <= 0 || row_bytes <= 0) return true;
	const int lpr = (row_bytes + 127) / 128;
	const long long total = (long long)nrows * lpr;
	int grid = (int)((total + 255) / 256);
	if (grid > 148 * 8) grid = 148 * 8;
	k_l2_prefetch<<<grid, 256, 0, st>>>((const char *)base, pitch_bytes, nrows, lpr);
	return rt_ok(cudaGetLastError(), err, "L2 prefetch launch");
#else
	(void)base; (void)pitch_bytes; (void)nrows; (void)row_bytes; (void)st; (void)err;
	return true;
#endif
}

bool launch_spec_resolve(char prec, const OpAny &op, const double *acc, double *scale_z, rt_stream st, std::string &err) {
#if DSP_GPU
	if (prec == 'f') k_spec_resolve<float><<<1, 32, 0, st>>>(op, acc, scale_z);
	else k_spec_resolve<double><<<1, 32, 0, st>>>(op, acc, scale_z);
	return rt_ok(cudaGetLastError(), err, "spec resolve launch");
#else
	(void)st; (void)err;
	if (prec == 'f') spec_resolve_range<float>(op, acc, scale_z);
	else spec_resolve_range<double>(op, acc, scale_z);
	return true;
#endif
}

}  // namespace dsp
