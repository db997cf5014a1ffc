// Instantiations + launcher of the packed-FP32 stage kernel (stage_packed.cuh): fp32, 32x32 and 16x16.
// Compiled twice: as is (pure-ABRelu stages) and from stage_packed_erf.cu with NTK_PACKED_ERF = 1 (Erf-capable
// instantiations, their own copy of the constant tables).
#include "stage_packed.cuh"

#ifndef NTK_PACKED_ERF
#define NTK_PACKED_ERF 0
#endif
#if NTK_PACKED_ERF
#define NTK_PACKED_LAUNCH launch_stage_packed_erf
#define NTK_PACKED_CONFIGURE stage_packed_erf_configure
#else
#define NTK_PACKED_LAUNCH launch_stage_packed
#define NTK_PACKED_CONFIGURE stage_packed_configure
#endif

namespace ntk {

namespace {

inline int variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("NTK_B200_PVAR");
    v = e ? atoi(e) : 0;
  }
  return v;
}

template <int S, int L, int IN, int EPI, bool NTK, int CIN, bool RC, int LAG = 1, int MINB = 2, bool Q2P = true,
          int VAR = 0, int WPT = 8>
int launch_p_impl(cudaStream_t stream, int64_t* launches, const StageArgs<float>& a) {
  using G = PGeom<S, WPT>;
  auto kern = k_stage_p<S, L, IN, EPI, NTK, CIN, RC, LAG, MINB, Q2P, NTK_PACKED_ERF != 0, VAR, WPT>;
  const size_t smem = stage_p_smem_bytes<S, L, IN, EPI, NTK, CIN, Q2P, WPT>();
  NTK_TRY(ensure_dynamic_smem((const void*)kern, smem));
  const long long blocks = (a.P + G::GROUPS - 1) / G::GROUPS;
  (*launches)++;
  kern<<<(unsigned)blocks, G::NT, smem, stream>>>(a);
  NTK_CUDA(cudaGetLastError());
  return NTK_OK;
}

template <int S, int L, int IN, bool NTK, int CIN>
int launch_p_epi(cudaStream_t stream, int64_t* launches, int epi, const StageArgs<float>& a) {
  if (!NTK && a.col_count != S) {  // self-pair pipeline (nngp only): partial column range
    switch (epi) {
      case EPI_STORE:
        return launch_p_impl<S, L, IN, EPI_STORE, false, CIN, true>(stream, launches, a);
      case EPI_POOL:
        return launch_p_impl<S, L, IN, EPI_POOL, false, CIN, true>(stream, launches, a);
      default:
        return fail(NTK_EINVAL, "partial column range with a GAP epilogue");
    }
  }
  switch (epi) {
    case EPI_STORE:
      return launch_p_impl<S, L, IN, EPI_STORE, NTK, CIN, false>(stream, launches, a);
    case EPI_POOL:
      // Variants A/B-timed on B200 for the dominant kernel (L = 3, FROM_X, POOL, ntk; 9216 pairs):
      //   LAG 1, 2 CTAs/SM, paired q2 (default) 37.87 ms | LAG 2: 37.73 | 3 CTAs/SM + planar q2: 37.29
      //   LAG 2 + 3 CTAs (spills): 39.42 | planar q2 at 2 CTAs: 38.05 | pooled gather consumed one step later (its
      //   LDS latency hidden, +18 registers): 0.5 % slower | K and U box-filter instructions interleaved to share
      //   masks through the reuse cache: ptxas sets no .reuse, 0.9 % slower.  All within 2 %: the kernel is bound
      //   by register-file read bandwidth (profiles/microbench/rf_bandwidth.cu), not latency or occupancy.
      if constexpr (!NTK_PACKED_ERF && L == 3 && IN == IN_FROM_X && NTK && S == 32) {
        if (variant() == 2) return launch_p_impl<S, L, IN, EPI_POOL, NTK, CIN, false, 1, 3, false>(stream, launches, a);
      }
      // VAR bit 1 = bias-free convs (b_std = 0, every Myrtle config): the predicated bias adds are compiled out.
      // Measured on B200 for the dominant kernel: 38.43 -> 37.07 ms per 9216 pairs, results bit-identical
      // (profiles/check_variant.py), so it is the default whenever the stage has no bias; NTK_B200_PVAR=7 forces
      // the general instantiation for A/B runs.
      // VAR bit 0 = degree-7 fit of G (NTK_B200_PVAR=4, or 6 together with bit 1).  Measured in round 2 on B200
      // (bench lines of the same box, 5 steps): default 37.10 ms | PVAR=6 36.60 ms (-1.3 %) | PVAR=2 (3 CTAs / SM,
      // planar q2) 37.81 | PVAR=7 (with bias adds) 38.46 | PVAR=4 38.37.  1.3 % does not pay for 5x the error of G and the
      // loss of the exact duplicate-pair diagonal (the self-pair runs keep the degree-8 kernel): not the default.
      if constexpr (!NTK_PACKED_ERF && L == 3 && IN == IN_FROM_X && NTK && S == 32) {
        const bool no_bias = a.lp[0].bias == 0.f && a.lp[1].bias == 0.f && a.lp[2].bias == 0.f;
        if (variant() == 20 && no_bias)  // 4 w per thread: 256 threads per pair, 4 warps per scheduler
          return launch_p_impl<S, L, IN, EPI_POOL, NTK, CIN, false, 1, 2, true, 2, 4>(stream, launches, a);
        if (variant() == 21 && no_bias)
          return launch_p_impl<S, L, IN, EPI_POOL, NTK, CIN, false, 2, 2, true, 2, 4>(stream, launches, a);
        if (variant() == 4) return launch_p_impl<S, L, IN, EPI_POOL, NTK, CIN, false, 1, 2, true, 1>(stream, launches, a);
        if (variant() == 6 && no_bias)
          return launch_p_impl<S, L, IN, EPI_POOL, NTK, CIN, false, 1, 2, true, 3>(stream, launches, a);
        if (variant() != 7 && no_bias)
          return launch_p_impl<S, L, IN, EPI_POOL, NTK, CIN, false, 1, 2, true, 2>(stream, launches, a);
      }
      // the same bias-free instantiation for the two-layer first stage of Myrtle-5 / Myrtle-7 (BASELINE configs[1], [2])
      if constexpr (!NTK_PACKED_ERF && L == 2 && IN == IN_FROM_X && NTK && S == 32) {
        if (variant() != 7 && a.lp[0].bias == 0.f && a.lp[1].bias == 0.f)
          return launch_p_impl<S, L, IN, EPI_POOL, NTK, CIN, false, 1, 2, true, 2>(stream, launches, a);
      }
      return launch_p_impl<S, L, IN, EPI_POOL, NTK, CIN, false>(stream, launches, a);
    default:
      return launch_p_impl<S, L, IN, EPI_GAP, NTK, CIN, false>(stream, launches, a);
  }
}

template <int S, int IN, bool NTK, int CIN>
int launch_p_L(cudaStream_t stream, int64_t* launches, int L, int epi, const StageArgs<float>& a) {
  switch (L) {
    case 1:
      return launch_p_epi<S, 1, IN, NTK, CIN>(stream, launches, epi, a);
    case 2:
      return launch_p_epi<S, 2, IN, NTK, CIN>(stream, launches, epi, a);
    default:
      return launch_p_epi<S, 3, IN, NTK, CIN>(stream, launches, epi, a);
  }
}

}  // namespace

int NTK_PACKED_LAUNCH(cudaStream_t stream, int64_t* launches, int S, int L, int from_x, int epi, bool ntk,
                      const StageArgs<float>& a) {
  if (S == 32) {
    if (from_x)
      return ntk ? launch_p_L<32, IN_FROM_X, true, 3>(stream, launches, L, epi, a)
                 : launch_p_L<32, IN_FROM_X, false, 3>(stream, launches, L, epi, a);
    return ntk ? launch_p_L<32, IN_LOAD, true, 1>(stream, launches, L, epi, a)
               : launch_p_L<32, IN_LOAD, false, 1>(stream, launches, L, epi, a);
  }
  if (S == 16) {
    if (from_x)
      return ntk ? launch_p_L<16, IN_FROM_X, true, 3>(stream, launches, L, epi, a)
                 : launch_p_L<16, IN_FROM_X, false, 3>(stream, launches, L, epi, a);
    return ntk ? launch_p_L<16, IN_LOAD, true, 1>(stream, launches, L, epi, a)
               : launch_p_L<16, IN_LOAD, false, 1>(stream, launches, L, epi, a);
  }
  return fail(NTK_EINVAL, "packed stage kernel is instantiated for S == 32 and 16");
}

int NTK_PACKED_CONFIGURE() {
  for (int S : {32, 16}) {
    std::vector<float4> m((size_t)S * S);
    for (int ch = 0; ch < S; ++ch)
      for (int h = 0; h < S; ++h) {
        const int h2 = (h + ch) % S;
        const float vu = (h > 0 && h2 != 0) ? 1.f : 0.f, vd = (h < S - 1 && h2 != S - 1) ? 1.f : 0.f;
        m[(size_t)ch * S + h] = make_float4(vu, vu, vd, vd);
      }
    const size_t bytes = m.size() * sizeof(float4);
    if (S == 32) NTK_CUDA(cudaMemcpyToSymbol(c_vmaskp32, m.data(), bytes));
    if (S == 16) NTK_CUDA(cudaMemcpyToSymbol(c_vmaskp16, m.data(), bytes));
  }
  return NTK_OK;
}

}  // namespace ntk
