// generated shape: one kernel family per translation unit (see kern_inst.cuh) -- planar row specialisation
#define KERN_T float
#define KERN_SUFFIX f32p
#define KERN_ROW 1
#define KERN_FAST 1
#define KERN_PLANAR 1
#include "kern_inst.cuh"
