// Embedded-size instantiations of the stage kernels with the general activation family (fused_kernels.cuh, EMB = true,
// ERF = 2) for float: MNIST-sized / non-square inputs and VALID stacks with Erf / Gelu / Sin / Cos / Rbf layers.
#include "instantiate.cuh"
namespace ntk {
NTK_FUSED_EMB_GEN_INSTANCES(, float)
}  // namespace ntk
