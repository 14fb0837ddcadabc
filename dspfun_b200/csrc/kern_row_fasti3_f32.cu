// generated shape: one kernel family per translation unit (see kern_inst.cuh) -- RGB-interleaved row specialisation
#define KERN_T float
#define KERN_SUFFIX f32i3
#define KERN_ROW 1
#define KERN_FAST 1
#define KERN_PLANAR 1
#define KERN_DD 3
// rows up to 1024 points: 3 CTAs per SM (80 registers; batch1024 rows 2.65 -> 2.86 TB/s forward, 2.29 -> 2.53 inverse; 4 CTAs
// spill and lose).  Longer rows are limited by shared memory, a register cap only costs them spills (plane4096x3: 74 -> 67).
// FAST >> 8 is log2 of the fixed length (kern_inst.cuh).
#define KERN_ROW_MINB ((((FAST) >> 8) >= 8 && ((FAST) >> 8) <= 10) ? 3 : 2)
#include "kern_inst.cuh"
