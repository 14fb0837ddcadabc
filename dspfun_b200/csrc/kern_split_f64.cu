// generated shape: split column pass kernels, double (see kern_split_inst.cuh)
#define KERN_T double
#define KERN_SUFFIX f64
#define KERN_IS_F32 0
#include "kern_split_inst.cuh"
