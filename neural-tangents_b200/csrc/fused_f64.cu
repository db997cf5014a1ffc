// Stage kernels (fused_kernels.cuh) for double: one translation unit per dtype (parallel build).
#include "instantiate.cuh"
namespace ntk {
NTK_FUSED_ERF_INSTANCES(extern, double)
NTK_FUSED_EMB_INSTANCES(extern, double)
NTK_FUSED_EMB_GEN_INSTANCES(extern, double)
NTK_FUSED_GEN_INSTANCES(extern, double)
NTK_FUSED_INSTANCES(, double)
}  // namespace ntk
