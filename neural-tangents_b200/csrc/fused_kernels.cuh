// Fused diagonal-marching fast path (placeholder until the kernels land).
#pragma once

#include "common.cuh"
#include "generic_kernels.cuh"

namespace ntk {

struct FusedPlan {
  bool ok = false;
};

inline FusedPlan plan_fused(const std::vector<ntk_op_t>&, int, int) { return FusedPlan(); }

template <typename T>
bool fused_supported(const FusedPlan&, int, int, int) {
  return false;
}

template <typename T>
int fused_gram(const FusedPlan&, Arena&, cudaStream_t, int64_t*, const T*, int, const T*, int, bool,
               int, int, int, bool, T*, T*, long long) {
  return fail(NTK_EUNSUPPORTED, "fused path not built");
}

inline int fused_configure_device() { return NTK_OK; }

// FCN input Gram x1 x2^T / d  (requirements.py:585-638 for 2-D inputs).
template <typename T>
int fcn_input_gram(bool dry, cudaStream_t stream, int64_t* launches, const T* x1, int t1, const T* x2,
                   int t2, int C, T* out) {
  (*launches)++;
  if (!dry) {
    const long long P = (long long)t1 * t2;
    k_rowdot<T><<<grid_for(P * 32), kThreads, 0, stream>>>(x1, x2, out, P, PairMap{t2, 0}, C,
                                                           (T)(1.0 / (double)C));
    NTK_CUDA(cudaGetLastError());
  }
  return NTK_OK;
}

}  // namespace ntk
