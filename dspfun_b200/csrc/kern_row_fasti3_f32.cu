// generated shape: one kernel family per translation unit (see kern_inst.cuh) -- RGB-interleaved row specialisation
#define KERN_T float
#define KERN_SUFFIX f32i3
#define KERN_ROW 1
#define KERN_FAST 1
#define KERN_PLANAR 1
#define KERN_DD 3
#include "kern_inst.cuh"
