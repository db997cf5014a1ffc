// Residual / diagonal-column kernels (res_kernels.cuh) for double.
#include "instantiate.cuh"
namespace ntk {
NTK_RES_ERF_INSTANCES(extern, double)
NTK_RES_INSTANCES(, double)
}  // namespace ntk
