// Residual / diagonal-column kernels (res_kernels.cuh) for float.
#include "instantiate.cuh"
namespace ntk {
NTK_RES_ERF_INSTANCES(extern, float)
NTK_RES_INSTANCES(, float)
}  // namespace ntk
