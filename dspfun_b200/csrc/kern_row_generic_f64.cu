// generated shape: one kernel family per translation unit (see kern_inst.cuh)
#define KERN_T double
#define KERN_SUFFIX f64
#define KERN_ROW 1
#define KERN_FAST 0
#include "kern_inst.cuh"
