// kern_misc.cu -- small helper kernels
#include "dsp_kernels.h"

namespace dsp {

#if DSP_GPU
template <class T> __global__ void k_spec_resolve(OpAny op, const double *acc, double *scale_z) {
	if (threadIdx.x == 0 && blockIdx.x == 0) spec_resolve_range<T>(op, acc, scale_z);
}
#endif

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
