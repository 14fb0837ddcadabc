/* raw-file stand-in for FFmpeg (tools/ffstub): everything lives in ffstub.h */
#include "../ffstub.h"
