// General-activation instantiations of the stage kernels (fused_kernels.cuh, ERF = 2) for double:
// stages that contain Gelu / Sin / Cos / Rbf layers (elementwise.py:195-400), alone or mixed with ABRelu / Erf.
#include "instantiate.cuh"
namespace ntk {
NTK_FUSED_GEN_INSTANCES(, double)
}  // namespace ntk
