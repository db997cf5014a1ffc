// Erf-capable instantiations of the packed-FP32 stage kernel: stage_packed.cu compiled with ERF = true.
#define NTK_PACKED_ERF 1
#include "stage_packed.cu"
