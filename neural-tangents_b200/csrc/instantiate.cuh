// Explicit-instantiation plumbing: the heavy kernel families are compiled in their own
// translation units (fused_f32.cu, fused_f64.cu, res_f32.cu, res_f64.cu) so that `make -j`
// builds them in parallel; executor.cu only sees `extern template` declarations.
#pragma once

#include "fused_kernels.cuh"
#include "res_kernels.cuh"

// launch_stage_k<T, NTK, ERF = true> (Erf-capable unpacked stage kernels) lives in fused_*_erf.cu
#define NTK_FUSED_ERF_INSTANCES(KW, T)                                                                     \
  KW template int launch_stage_k<T, true, true>(cudaStream_t, int64_t*, int, int, int, int, int,           \
                                                const StageArgs<T>&);                                      \
  KW template int launch_stage_k<T, false, true>(cudaStream_t, int64_t*, int, int, int, int, int,          \
                                                 const StageArgs<T>&);                                     \
  KW template int fused_configure_device<T, true>();

// launch_stage_k<T, NTK, ERF = 2> (general family: ABRelu / Erf / Gelu / Sin / Rbf) lives in fused_*_gen.cu
#define NTK_FUSED_GEN_INSTANCES(KW, T)                                                                     \
  KW template int launch_stage_k<T, true, 2>(cudaStream_t, int64_t*, int, int, int, int, int,              \
                                             const StageArgs<T>&);                                         \
  KW template int launch_stage_k<T, false, 2>(cudaStream_t, int64_t*, int, int, int, int, int,             \
                                              const StageArgs<T>&);                                        \
  KW template int fused_configure_device<T, 2>();

// launch_stage_emb<T, NTK> (embedded-size scalar stage kernels, C in {1, 3}) lives in fused_*_emb.cu
#define NTK_FUSED_EMB_INSTANCES(KW, T)                                                                     \
  KW template int launch_stage_emb<T, true, 0>(cudaStream_t, int64_t*, int, int, int, int, int,            \
                                               const StageArgs<T>&);                                       \
  KW template int launch_stage_emb<T, false, 0>(cudaStream_t, int64_t*, int, int, int, int, int,           \
                                                const StageArgs<T>&);

// launch_stage_emb<T, NTK, 2> (embedded sizes, general activation family) lives in fused_*_emb_gen.cu
#define NTK_FUSED_EMB_GEN_INSTANCES(KW, T)                                                                 \
  KW template int launch_stage_emb<T, true, 2>(cudaStream_t, int64_t*, int, int, int, int, int,            \
                                               const StageArgs<T>&);                                       \
  KW template int launch_stage_emb<T, false, 2>(cudaStream_t, int64_t*, int, int, int, int, int,           \
                                                const StageArgs<T>&);

#define NTK_FUSED_INSTANCES(KW, T)                                                                        \
  KW template int fused_gram<T>(const FusedPlan&, Arena&, cudaStream_t, int64_t*, StageProfile*, const T*, \
                                int, const T*, int, bool, int, int, int, bool, T*, T*, long long, bool, bool); \
  KW template int fused_configure_device<T, false>();

// launch_res<T, NTK, ERF = true> (the Erf-capable residual kernels) lives in res_*_erf.cu
#define NTK_RES_ERF_INSTANCES(KW, T)                                                                       \
  KW template int launch_res<T, true, true>(cudaStream_t, int64_t*, int, bool, const ResArgs<T>&);         \
  KW template int launch_res<T, false, true>(cudaStream_t, int64_t*, int, bool, const ResArgs<T>&);

#define NTK_RES_INSTANCES(KW, T)                                                                         \
  KW template int res_gram<T>(const ResPlan&, Arena&, cudaStream_t, int64_t*, const T*, int, const T*,    \
                              int, bool, int, int, bool, T*, T*, long long, bool);                        \
  KW template int diag_gram<T>(const DiagPlan&, Arena&, cudaStream_t, int64_t*, const T*, int, const T*, \
                               int, bool, int, int, bool, T*, T*, long long, bool);
