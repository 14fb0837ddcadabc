// generated shape: one kernel family per translation unit (see kern_inst.cuh)
#define KERN_T float
#define KERN_SUFFIX f32
#define KERN_ROW 0
#define KERN_FAST 1
#include "kern_inst.cuh"
