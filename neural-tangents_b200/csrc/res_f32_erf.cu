// Erf-capable instantiations of the residual kernels (res_kernels.cuh) for float.
#include "instantiate.cuh"
namespace ntk {
NTK_RES_ERF_INSTANCES(, float)
}  // namespace ntk
