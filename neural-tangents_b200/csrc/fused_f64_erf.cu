// Erf-capable instantiations of the stage kernels (fused_kernels.cuh) for double.
#include "instantiate.cuh"
namespace ntk {
NTK_FUSED_ERF_INSTANCES(, double)
}  // namespace ntk
