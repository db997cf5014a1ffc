// Stage kernels (fused_kernels.cuh) for float: one translation unit per dtype (parallel build).
#include "instantiate.cuh"
namespace ntk {
NTK_FUSED_ERF_INSTANCES(extern, float)
NTK_FUSED_EMB_INSTANCES(extern, float)
NTK_FUSED_EMB_GEN_INSTANCES(extern, float)
NTK_FUSED_GEN_INSTANCES(extern, float)
NTK_FUSED_INSTANCES(, float)
}  // namespace ntk
