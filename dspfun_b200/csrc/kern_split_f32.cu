// generated shape: split column pass kernels, float (see kern_split_inst.cuh)
#define KERN_T float
#define KERN_SUFFIX f32
#define KERN_IS_F32 1
#include "kern_split_inst.cuh"
