// Embedded-size instantiations of the stage kernels (fused_kernels.cuh, EMB = true) for double:
// images of any H x W <= 32 x 32 with 1 or 3 channels (MNIST 28x28x1, ...), pure ABRelu stages.
#include "instantiate.cuh"
namespace ntk {
NTK_FUSED_EMB_INSTANCES(, double)
}  // namespace ntk
