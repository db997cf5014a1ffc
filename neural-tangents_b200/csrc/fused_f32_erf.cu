// Erf-capable instantiations of the stage kernels (fused_kernels.cuh) for float.
#include "instantiate.cuh"
namespace ntk {
NTK_FUSED_ERF_INSTANCES(, float)
}  // namespace ntk
