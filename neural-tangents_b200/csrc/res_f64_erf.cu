// Erf-capable instantiations of the residual kernels (res_kernels.cuh) for double.
#include "instantiate.cuh"
namespace ntk {
NTK_RES_ERF_INSTANCES(, double)
}  // namespace ntk
