// Packed-FP32 stage kernel: the fp32 / 32x32 instance of the fused diagonal-marching kernel of
// fused_kernels.cuh, re-expressed on Blackwell's two-wide FP32 instructions (PTX fma/mul/add
// .f32x2 -> SASS FFMA2 / FMUL2 / FADD2).
//
// Why: k_stage is issue-bound (profiles/README.md: 42 warp instructions per element-layer, issue
// slots 73 % busy, FMA pipe 51 %).  profiles/microbench/fp32x2_pipe.cu shows that a packed
// instruction moves the same lanes per clock as two scalar ones (0.5 FFMA2 / clk / SMSP with two
// register sources, 0.33 with three: the limit is register-file bandwidth either way) but takes ONE
// issue slot.  So packing halves the issue pressure of the arithmetic and lets the FMA pipe fill.
//
// What changes against k_stage (same math, same reference rules, same external tensor formats):
//  * A thread still owns 8 consecutive w at fixed (ch, cw); they are held as 4 pairs (i, i+4).
//    With this stride every horizontal neighbour pair is itself a held pair except one per side,
//    which is assembled from the shuffled halo: no register shuffling inside the thread.
//  * The NTK tensor is carried as U = T + K:  conv(U) + b = conv(T) + (conv(K) + b) is exactly the
//    reference's `ntk <- conv(ntk) + nngp_new` (linear.py:1396-1398) without the cross-lane add,
//    and the activation emits U' = kd * T_in + K' with one FMA.  T = U - K is formed only where a
//    tensor leaves the kernel.
//  * 1/sqrt(q1 q2) comes from MUFU.RSQ of the product instead of a second stored map: half the
//    q-map shared memory and one multiply less on the (binding) FMA pipe.  The XU pipe has room.
//  * q-maps are re-laid out in shared memory so that every operand pair is one aligned 64-bit
//    (q2) or 128-bit (q1) load: q1 rows are permuted inside each 8-block, q2 rows hold the
//    (e, e+4) pairs explicitly (and negated, so that K^2 - q1 q2 is a plain packed add).
//  * The nngp lane performs exactly the operations of k_stage / k_qmaps in the same order, so a
//    duplicate pair still sees q1 q2 - K^2 == 0 on its diagonal (SURVEY §7 "FP32 accuracy").
#pragma once

#include "fused_kernels.cuh"

namespace ntk {

static __constant__ float4 c_vmaskp32[32 * 32];  // (vU, vU, vD, vD) per marched row
static __constant__ float4 c_vmaskp16[16 * 16];

template <int S>
__device__ __forceinline__ float4 vmaskp_at(int r) {
  if (S == 32) return c_vmaskp32[r];
  return c_vmaskp16[r];
}

__device__ __forceinline__ float2 f2(float lo, float hi) { return make_float2(lo, hi); }
__device__ __forceinline__ float2 f2s(float v) { return make_float2(v, v); }

__device__ __forceinline__ float rsq_fast(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ float copysign_bits(float mag, float sgn) {
  return __uint_as_float((__float_as_uint(mag) & 0x7fffffffu) | (__float_as_uint(sgn) & 0x80000000u));
}

__device__ __forceinline__ float lds_f1(unsigned addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

__device__ __forceinline__ float2 lds_f2(unsigned addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}

// -G(c) = -acos(c)/sqrt(1-c^2) on [0,1]: the degree-8 fit of acos_over_sin with negated
// coefficients, two elements per instruction.
__device__ __forceinline__ float2 neg_acos_over_sin2(float2 c) {
  float2 g = f2s(-0.017638931050896645f);
  g = __ffma2_rn(g, c, f2s(0.09791343659162521f));
  g = __ffma2_rn(g, c, f2s(-0.2531546652317047f));
  g = __ffma2_rn(g, c, f2s(0.42450031638145447f));
  g = __ffma2_rn(g, c, f2s(-0.5569935441017151f));
  g = __ffma2_rn(g, c, f2s(0.6610828042030334f));
  g = __ffma2_rn(g, c, f2s(-0.7848954796791077f));
  g = __ffma2_rn(g, c, f2s(0.9999822378158569f));
  g = __ffma2_rn(g, c, f2s(-1.570796251296997f));
  return g;
}

// Degree-7 minimax fit of -G (|err| < 5.6e-7, the size of one fp32 rounding of G): one FFMA2 less per pair.
// NOT the default -- kept for an A/B measurement (NTK_B200_PVAR=4): the degree-8 fit keeps the fp32 path at
// 1.8e-7 of the float64 result.
__device__ __forceinline__ float2 neg_acos_over_sin2_deg7(float2 c) {
  float2 g = f2s(0.028090565068060184f);
  g = __ffma2_rn(g, c, f2s(-0.14106766428166093f));
  g = __ffma2_rn(g, c, f2s(0.33101607627600144f));
  g = __ffma2_rn(g, c, f2s(-0.5139260110419339f));
  g = __ffma2_rn(g, c, f2s(0.6503503954074816f));
  g = __ffma2_rn(g, c, f2s(-0.7835891524309508f));
  g = __ffma2_rn(g, c, f2s(0.9999221177978991f));
  g = __ffma2_rn(g, c, f2s(-1.5707957734418327f));
  return g;
}

// ABRelu on two elements (elementwise.py:444-455; see act_point in fused_kernels.cuh).
//   K         conv(K_in) + b            (the reference's nngp after the conv)
//   U         conv(U_in) + b = conv(T_in) + K, i.e. the reference's ntk after the conv (linear.py:1396-1398)
//   q1, nq2   q1 and -q2 of the two elements
//   Ko = coef*s + kd*K  (K'),   Uo = kd*U + Ko  (= T' + K', the carried form of the next layer's ntk)
template <bool NTK, bool POLY7 = false>
__device__ __forceinline__ void act_pair(float2 K, float2 U, float2 q1, float2 nq2, float2 coef2,
                                         float2 hab2, float2& Ko, float2& Uo) {
  const float2 np = __fmul2_rn(q1, nq2);                 // -(q1 q2), exact negation of the product
  // K^2 - q1 q2 must stay UNFUSED (an exact 0 for duplicate pairs).  ptxas contracts a packed
  // mul.rn.f32x2 feeding an add.rn.f32x2 into FFMA2 (even under -fmad=false), so the subtraction
  // is done with two scalar adds, which it does not merge with a packed multiply.
  const float2 kk = __fmul2_rn(K, K);
  float2 d;
  d.x = __fadd_rn(kk.x, np.x);
  d.y = __fadd_rn(kk.y, np.y);
  float2 rb, s;
  rb.x = rsq_fast(fabsf(np.x));
  rb.y = rsq_fast(fabsf(np.y));
  s.x = sqrt_fast(fabsf(d.x));
  s.y = sqrt_fast(fabsf(d.y));
  const float2 sn = __fmul2_rn(s, rb);
  float2 ac;
  ac.x = __fmul_rn(fabsf(K.x), rb.x);
  ac.y = __fmul_rn(fabsf(K.y), rb.y);
  const float2 u = __ffma2_rn(sn, POLY7 ? neg_acos_over_sin2_deg7(ac) : neg_acos_over_sin2(ac), f2s(kHalfPiF));
  float2 us;
  us.x = copysign_bits(u.x, K.x);
  us.y = copysign_bits(u.y, K.y);
  const float2 kd = __ffma2_rn(coef2, us, hab2);
  Ko = __ffma2_rn(kd, K, __fmul2_rn(coef2, s));
  if (NTK) Uo = __ffma2_rn(kd, U, Ko);
}

// Measured and rejected in round 2 (profiles/ab_variants_r02.log): G from the texture unit (piecewise-linear table,
// one TEX per element instead of the 8-step Horner chain; 25.1 instead of 26.9 instructions per element-layer, same
// accuracy) ran 2.8x SLOWER (103.8 vs 37.1 ms): fp32 filtered fetches run at ~1 texel / clk / SM on B200.

// Erf on two elements (elementwise.py:67-112; erf_act_point in fused_kernels.cuh).  The q-map of an Erf layer
// holds D = 1 + 2 b^2 q, so q1 = D1 and nq2 = -D2 here; with Kh = 2 b^2 K:
//   s = sqrt(D1 D2 - Kh^2) (>= 1),  c = Kh / sqrt(D1 D2),  K' = eA asin(c) + eC,  T' = eT T_conv / s,
// asin(c) = copysign(pi/2 - sqrt(1-c^2) G(|c|), c) with the same fit as the ABRelu path; U' = T' + K'.
template <bool NTK>
__device__ __forceinline__ void act_pair_erf(float2 K, float2 U, float2 q1, float2 nq2, float e_in, float eA,
                                             float eT, float eC, float2& Ko, float2& Uo) {
  const float2 Kh = __fmul2_rn(f2s(e_in), K);
  const float2 np = __fmul2_rn(q1, nq2);               // -(D1 D2)
  const float2 d = __ffma2_rn(Kh, Kh, np);             // Kh^2 - D1 D2 <= -1: no cancellation to protect
  float2 rb, rs;
  rb.x = rsq_fast(fabsf(np.x));
  rb.y = rsq_fast(fabsf(np.y));
  rs.x = rsq_fast(fabsf(d.x));
  rs.y = rsq_fast(fabsf(d.y));
  float2 ad;
  ad.x = fabsf(d.x);
  ad.y = fabsf(d.y);
  const float2 s = __fmul2_rn(ad, rs);
  const float2 sn = __fmul2_rn(s, rb);
  float2 ac;
  ac.x = fminf(__fmul_rn(fabsf(Kh.x), rb.x), 1.f);
  ac.y = fminf(__fmul_rn(fabsf(Kh.y), rb.y), 1.f);
  const float2 u = __ffma2_rn(sn, neg_acos_over_sin2(ac), f2s(kHalfPiF));
  float2 us;
  us.x = copysign_bits(u.x, Kh.x);
  us.y = copysign_bits(u.y, Kh.y);
  Ko = __ffma2_rn(f2s(eA), us, f2s(eC));
  if (NTK) Uo = __ffma2_rn(__fmul2_rn(f2s(eT), rs), U, Ko);
}

template <int S, int WPT_ = 8>
struct PGeom {
  static constexpr int WPT = WPT_;
  static constexpr int TPP = S * S / WPT;          // threads per pair
  static constexpr int NT = TPP < 128 ? 128 : TPP; // threads per CTA
  static constexpr int GROUPS = NT / TPP;
  static constexpr int LPG = TPP < 32 ? TPP : 32;
  static constexpr int NWB = S / WPT;
  static constexpr int LW = LPG / NWB;
  static constexpr int NR = S * S;
};

// position of column w inside a permuted q1 row: pairs (i, i + WPT/2) of each WPT-block are adjacent
template <int WPT = 8>
__host__ __device__ __forceinline__ int perm8(int w) {
  constexpr int NP = WPT / 2;
  const int i = w % WPT;
  return (w - i) + ((i % NP) << 1) + (i / NP);
}

template <int S, int L, int IN, int EPI, bool NTK, int CIN, bool Q2P = true, int WPT = 8>
size_t stage_p_smem_bytes() {
  using G = PGeom<S, WPT>;
  const int xs1 = IN == IN_FROM_X ? S * S * CIN : 0;
  const int xs2 = IN == IN_FROM_X ? S * S * 4 : 0;
  const int q1 = L * S * S, q2 = (Q2P ? 2 : 1) * L * S * S;
  const int stg = EPI == EPI_POOL ? 2 * (NTK ? 2 : 1) * S * (S + 1) : 0;
  return (size_t)G::GROUPS * (xs1 + xs2 + q1 + q2 + stg) * sizeof(float);
}

// Variant knobs (kept as template parameters so that they can be A/B-timed on the device):
//   LAG   rows of lag between consecutive fused layers: 1 = a layer consumes the row its predecessor
//         produced in the same step; 2 = the row of the previous step, so the L layer blocks of a step
//         are independent (more ILP, more live registers)
//   MINB  __launch_bounds__ min CTAs per SM (2 or 3)
//   Q2P   q2 rows hold explicit (e, e+4) pairs (one LDS.64 per pair, 2x shared memory) or are planar
//         (two LDS.32 per pair)
//   ERF   the stage may contain Erf layers (runtime branch per layer); ERF = false carries no Erf code
//   VAR   bit 1: the convs have no bias (the predicated bias adds are compiled out; bit-identical results, used by
//         default for the dominant kernel when b_std = 0); bit 0: degree-7 fit of G, an unmeasured round-2
//         candidate reachable only through NTK_B200_PVAR (stage_packed.cu)
template <int S, int L, int IN, int EPI, bool NTK, int CIN, bool RC, int LAG = 1, int MINB = 2, bool Q2P = true,
          bool ERF = false, int VAR = 0, int WPT = 8>
__global__ void __launch_bounds__(PGeom<S, WPT>::NT, MINB)
k_stage_p(const StageArgs<float> a) {
  using G = PGeom<S, WPT>;
  constexpr int NP = WPT / 2;  // NP pairs (i, i + NP)
  constexpr int TPP = G::TPP, LPG = G::LPG, NWB = G::NWB, LW = G::LW, NR = G::NR;
  constexpr int SO = S / 2;
  constexpr int SP = S + 1;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int XS1 = IN == IN_FROM_X ? S * S * CIN : 0;
  constexpr int XS2 = IN == IN_FROM_X ? S * S * 4 : 0;
  constexpr int Q1 = L * S * S, Q2 = (Q2P ? 2 : 1) * L * S * S;
  constexpr int STG = EPI == EPI_POOL ? 2 * (NTK ? 2 : 1) * S * SP : 0;
  constexpr int PER_GROUP = XS1 + XS2 + Q1 + Q2 + STG;
  static_assert((XS1 % 4) == 0 && (XS2 % 4) == 0 && (Q1 % 4) == 0, "16-byte alignment of the smem parts");
  const int tid = threadIdx.x;
  const int grp = tid / TPP, tg = tid % TPP;
  float* sm = reinterpret_cast<float*>(smem_raw) + (size_t)grp * PER_GROUP;
  float* x1s = sm;
  float* x2s = x1s + XS1;
  float* q1A = x2s + XS2;   // [L][S][perm8(w)]
  float* q2B = q1A + Q1;    // [L][S][e] -> (-q[e], -q[(e+4)%S])
  float* stg = q2B + Q2;

  const int lig = tg % LPG, wig = tg / 32;
  const int wblk = lig / LW, cwsub = lig % LW;
  const int cw = wig * LW + cwsub;
  const int w0 = wblk * WPT;

  long long p = (long long)blockIdx.x * G::GROUPS + grp;
  const bool live = p < a.P;
  if (!live) p = a.P - 1;
  int si, sj;
  if (a.self) {
    si = sj = (int)p;
  } else if (a.tri) {
    int off;
    tri_unrank(p, a.n2, si, off);
    sj = si + off;
  } else {
    si = (int)(p / a.n2);
    sj = (int)(p % a.n2);
  }

  // ---- POOL: zero this pair's output (accumulated into with red.global.add below) -----------
  if (EPI == EPI_POOL && a.zero_out && live) {
    constexpr int OUT4 = SO * SO * SO * SO / 4;
    float4* zk = reinterpret_cast<float4*>(a.outK + p * (long long)(SO * SO * SO * SO));
    for (int e = tg; e < OUT4; e += TPP) zk[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (NTK) {
      float4* zt = reinterpret_cast<float4*>(a.outT + p * (long long)(SO * SO * SO * SO));
      for (int e = tg; e < OUT4; e += TPP) zt[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  // ---- stage the two samples and their q-maps -------------------------------------------
  {
    if (IN == IN_FROM_X) {
      const float* g1 = a.x1 + (long long)si * S * S * CIN;
      const float* g2 = a.x2 + (long long)sj * S * S * CIN;
      for (int e = tg; e < S * S * CIN; e += TPP) x1s[e] = __fmul_rn(g1[e], a.in_scale);
      for (int e = tg; e < S * S; e += TPP) {
#pragma unroll
        for (int c = 0; c < 4; ++c) x2s[e * 4 + c] = c < CIN ? g2[e * CIN + c] : 0.f;
      }
    }
    // global q-maps: [n][L][S][S] x (q, 1/sqrt(q)); only q is used here.  A tiny floor keeps
    // rsqrt(q1 q2) finite for all-zero receptive fields (K is 0 there, so the result is unchanged).
    constexpr float kQFloor = 1e-18f;
    const float2* g1 = reinterpret_cast<const float2*>(a.qm1) + (long long)si * L * S * S;
    const float2* g2 = reinterpret_cast<const float2*>(a.qm2) + (long long)sj * L * S * S;
    for (int e = tg; e < L * S * S; e += TPP) {
      const int w = e % S, row = e / S;
      q1A[row * S + perm8<WPT>(w)] = fmaxf(__ldg(&g1[e].x), kQFloor);
      const float nq = -fmaxf(__ldg(&g2[e].x), kQFloor);
      if (Q2P) {
        q2B[(row * S + w) * 2] = nq;
        q2B[(row * S + ((w + S - NP) % S)) * 2 + 1] = nq;
      } else {
        q2B[row * S + w] = nq;
      }
    }
  }
  if (TPP > 32)
    __syncthreads();
  else
    __syncwarp();

  // ---- per-thread constants --------------------------------------------------------------
  // lk[i]: link between w0+i-1 and w0+i is intact (both inside the image, w' does not wrap)
  float lk[WPT + 1];
#pragma unroll
  for (int i = 0; i <= WPT; ++i) {
    const int wl = w0 + i - 1, wr = w0 + i;
    lk[i] = (wl >= 0 && wr <= S - 1 && ((wl + cw) % S) != S - 1) ? 1.f : 0.f;
  }
  float2 mL[NP + 1];  // mL[j] = (lk[j], lk[j+4]): masks of the left links of pair j == right links of pair j-1
#pragma unroll
  for (int j = 0; j <= NP; ++j) mL[j] = f2(lk[j], lk[j + NP]);
  int off2[WPT];      // (w0 + i + cw) mod S
#pragma unroll
  for (int i = 0; i < WPT; ++i) off2[i] = (w0 + i + cw) % S;
  const unsigned q2base = (unsigned)__cvta_generic_to_shared(q2B);

  float2 RK[L][2][NP], RU[L][2][NP];
#pragma unroll
  for (int l = 0; l < L; ++l)
#pragma unroll
    for (int sl = 0; sl < 2; ++sl)
#pragma unroll
      for (int j = 0; j < NP; ++j) {
        RK[l][sl][j] = f2s(0.f);
        RU[l][sl][j] = f2s(0.f);
      }

  const int nrows = RC ? a.col_count * S : NR;
  // RC: the marched columns are {0..k} then {S-k..S-1} (k = col_start), i.e. the needed columns in
  // the SAME relative order as the full march, so the pooled outputs they feed are accumulated in the
  // same order as in a cross-pair run and a duplicate pair reproduces the self-pair diagonal exactly.
  const int col_k = RC ? a.col_start : 0;
  auto full_row = [&](int r) {
    if (!RC) return r;
    const int c = r / S;
    return (c <= col_k ? c : S - (a.col_count - c)) * S + (r % S);
  };

  float2 gap_k = f2s(0.f), gap_u = f2s(0.f);
  const float* inK = IN == IN_LOAD ? a.inK + p * (long long)NR * S * S + (long long)w0 * S + cw : nullptr;
  const bool has_inT = IN == IN_LOAD && NTK && a.inT != nullptr;  // nullptr: sheared input covariance, ntk = 0
  const float* inT = has_inT ? a.inT + p * (long long)NR * S * S + (long long)w0 * S + cw : nullptr;

  float nK[WPT], nT[WPT];
  auto fetch = [&](int r) {
    if (IN == IN_LOAD) {
      const int rc = full_row(r < nrows ? r : nrows - 1);
      const float* gk = inK + (long long)rc * S * S;
#pragma unroll
      for (int i = 0; i < WPT; ++i) nK[i] = __ldg(gk + i * S);
      if (NTK) {
        if (has_inT) {
          const float* gt = inT + (long long)rc * S * S;
#pragma unroll
          for (int i = 0; i < WPT; ++i) nT[i] = __ldg(gt + i * S);
        } else {
#pragma unroll
          for (int i = 0; i < WPT; ++i) nT[i] = 0.f;
        }
      }
    }
  };
  fetch(0);

  float2 BK[L][NP], BU[L][NP];
#pragma unroll
  for (int l = 0; l < L; ++l)
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      BK[l][j] = f2s(0.f);
      BU[l][j] = f2s(0.f);
    }

  auto step = [&](const int t, auto par_c) {
    constexpr int par = decltype(par_c)::value;
    float2 PK[NP], PU[NP];
    // ---- input row r = t of layer 1 ------------------------------------------------------
    if (IN == IN_FROM_X) {
      const int r = full_row(t < nrows ? t : nrows - 1);
      const int ch = r / S, h = r % S;
      const int h2 = (h + ch) % S;
      const float* xa = x1s + (h * S + w0) * CIN;
      const float4* xb = reinterpret_cast<const float4*>(x2s) + h2 * S;
      float acc[WPT];
#pragma unroll
      for (int i = 0; i < WPT; ++i) {
        const float4 b4 = xb[off2[i]];
        const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
        float v = __fmul_rn(xa[i * CIN], bb[0]);
#pragma unroll
        for (int c = 1; c < CIN; ++c) v = __fmaf_rn(xa[i * CIN + c], bb[c], v);
        acc[i] = v;
      }
#pragma unroll
      for (int j = 0; j < NP; ++j) PK[j] = f2(acc[j], acc[j + NP]);
    } else {
#pragma unroll
      for (int j = 0; j < NP; ++j) {
        PK[j] = f2(nK[j], nK[j + NP]);
        if (NTK) PU[j] = f2(__fadd_rn(nK[j], nT[j]), __fadd_rn(nK[j + NP], nT[j + NP]));
      }
      fetch(t + 1);
    }
#pragma unroll
    for (int li = 0; li < L; ++li) {
      // LAG == 2: last layer first, so BK[l-1] still holds the previous step's row
      const int l = LAG == 2 ? L - 1 - li : li;
      const int lm = l == 0 ? 0 : l - 1;
      const int slot = (par + (LAG == 1 ? l : 0)) & 1;  // == (t - LAG*l) & 1, compile-time
      const bool has_u = NTK && (l > 0 || IN == IN_LOAD);
#define INK(j) (l == 0 ? PK[j] : BK[lm][j])
#define INU(j) (l == 0 ? PU[j] : BU[lm][j])
      int r_out = t - LAG * l - 1;
      r_out = full_row(r_out < 0 ? 0 : (r_out > nrows - 1 ? nrows - 1 : r_out));
      const int ch = r_out / S, h = r_out % S;
      const int h2 = (h + ch) % S;
      const float4 vm = vmaskp_at<S>(r_out);
      const float2 vU = f2(vm.x, vm.y), vD = f2(vm.z, vm.w);
      // ---- vertical taps of the two OLD rows (last use of the oldest row) ------------------
      float2 tk[NP], tu[NP];
#pragma unroll
      for (int j = 0; j < NP; ++j) {
        tk[j] = __ffma2_rn(vU, RK[l][slot][j], RK[l][slot ^ 1][j]);
        if (has_u) tu[j] = __ffma2_rn(vU, RU[l][slot][j], RU[l][slot ^ 1][j]);
      }
      // ---- horizontal 3-tap of the new row into the freed ring slot --------------------------
      {
        float left = 0.f, right = 0.f, leftU = 0.f, rightU = 0.f;
        if (NWB > 1) {
          left = __shfl_up_sync(0xffffffffu, INK(NP - 1).y, LW);   // x[7] of the left w-block
          right = __shfl_down_sync(0xffffffffu, INK(0).x, LW);     // x[0] of the right w-block
          if (has_u) {
            leftU = __shfl_up_sync(0xffffffffu, INU(NP - 1).y, LW);
            rightU = __shfl_down_sync(0xffffffffu, INU(0).x, LW);
          }
        }
        // hsum3(Pm, P0, Pp) = fma(mR, Pp, fma(mL, Pm, P0)), pair-wise
        const float2 Lp0 = f2(left, INK(NP - 1).x), Rp3 = f2(INK(0).y, right);
#pragma unroll
        for (int j = 0; j < NP; ++j) {
          const float2 Pm = j == 0 ? Lp0 : INK(j == 0 ? 0 : j - 1);
          const float2 Pp = j == NP - 1 ? Rp3 : INK(j == NP - 1 ? j : j + 1);
          RK[l][slot][j] = __ffma2_rn(mL[j + 1], Pp, __ffma2_rn(mL[j], Pm, INK(j)));
        }
        if (has_u) {
          const float2 Lu0 = f2(leftU, INU(NP - 1).x), Ru3 = f2(INU(0).y, rightU);
#pragma unroll
          for (int j = 0; j < NP; ++j) {
            const float2 Pm = j == 0 ? Lu0 : INU(j == 0 ? 0 : j - 1);
            const float2 Pp = j == NP - 1 ? Ru3 : INU(j == NP - 1 ? j : j + 1);
            RU[l][slot][j] = __ffma2_rn(mL[j + 1], Pp, __ffma2_rn(mL[j], Pm, INU(j)));
          }
        }
      }
      // ---- finish the vertical sum, add the bias, apply the activation ---------------------
      {
        const float4* q1r = reinterpret_cast<const float4*>(q1A + (l * S + h) * S + w0);
        float2 q1p[NP];
#pragma unroll
        for (int j = 0; j < NP; j += 2) {
          const float4 qa = q1r[j / 2];
          q1p[j] = f2(qa.x, qa.y);
          q1p[j + 1] = f2(qa.z, qa.w);
        }
        const unsigned q2row = q2base + (unsigned)((l * S + h2) * S * (Q2P ? 8 : 4));
        const float2 coef2 = f2s(a.lp[l].coef), hab2 = f2s(a.lp[l].hab2);
        const float bias = a.lp[l].bias;
        const bool has_bias = !(VAR & 2) && bias != 0.f;
#pragma unroll
        for (int j = 0; j < NP; ++j) {
          float2 ck = __ffma2_rn(vD, RK[l][slot][j], tk[j]);
          if (has_bias) ck = __fadd2_rn(ck, f2s(bias));
          float2 cu = ck;  // T0 == 0 for the first layer of a FROM_X stage: U = K
          if (has_u) {
            cu = __ffma2_rn(vD, RU[l][slot][j], tu[j]);
            if (has_bias) cu = __fadd2_rn(cu, f2s(bias));
          }
          float2 nq2;
          if (Q2P) {
            nq2 = lds_f2(q2row + (unsigned)off2[j] * 8u);
          } else {
            nq2.x = lds_f1(q2row + (unsigned)off2[j] * 4u);
            nq2.y = lds_f1(q2row + (unsigned)off2[j + NP] * 4u);
          }
          if (ERF && a.lp[l].kind == ACT_ERF)
            act_pair_erf<NTK>(ck, cu, q1p[j], nq2, a.lp[l].e_in, a.lp[l].eA, a.lp[l].eT, a.lp[l].eC, BK[l][j], BU[l][j]);
          else
            act_pair<NTK, (VAR & 1) != 0>(ck, cu, q1p[j], nq2, coef2, hab2, BK[l][j], BU[l][j]);
        }
      }
#undef INK
#undef INU
    }
    // ---- epilogue on the finished row of the last layer ---------------------------------------
    const int r_fin = t - LAG * (L - 1) - 1;
    if (r_fin >= 0 && r_fin < nrows) {
      const int rf = full_row(r_fin);
      const int ch = rf / S, h = rf % S;
      if (EPI == EPI_STORE) {
        if (live) {
          const long long base = (p * NR + rf) * (long long)(S * S) + (long long)w0 * S + cw;
#pragma unroll
          for (int j = 0; j < NP; ++j) {
            a.outK[base + (long long)j * S] = BK[L - 1][j].x;
            a.outK[base + (long long)(j + NP) * S] = BK[L - 1][j].y;
            if (NTK) {
              a.outT[base + (long long)j * S] = __fsub_rn(BU[L - 1][j].x, BK[L - 1][j].x);
              a.outT[base + (long long)(j + NP) * S] = __fsub_rn(BU[L - 1][j].y, BK[L - 1][j].y);
            }
          }
        }
      } else if (EPI == EPI_GAP) {
#pragma unroll
        for (int j = 0; j < NP; ++j) {
          gap_k = __fadd2_rn(gap_k, BK[L - 1][j]);
          if (NTK) gap_u = __fadd2_rn(gap_u, BU[L - 1][j]);
        }
      } else {  // EPI_POOL: AvgPool 2x2/2 of both members (linear.py:3499-3572); U is pooled, T = U - K after
        float* sK = stg + (r_fin & 1) * ((NTK ? 2 : 1) * S * SP);
        float* sU = sK + S * SP;
#pragma unroll
        for (int j = 0; j < NP; ++j) {
          sK[(w0 + j) * SP + cw] = BK[L - 1][j].x;
          sK[(w0 + j + NP) * SP + cw] = BK[L - 1][j].y;
          if (NTK) {
            sU[(w0 + j) * SP + cw] = BU[L - 1][j].x;
            sU[(w0 + j + NP) * SP + cw] = BU[L - 1][j].y;
          }
        }
        if (TPP > 32)
          __syncthreads();
        else
          __syncwarp();
        const int ii = h & 1;
        int Ch;
        if ((ch & 1) == 0)
          Ch = ch >> 1;
        else
          Ch = (ii == 0 ? (ch - 1) >> 1 : ((ch + 1) >> 1) % SO);
        const long long obase = ((p * SO + Ch) * SO + (h >> 1)) * (long long)(SO * SO);
#pragma unroll
        for (int k = 0; k < (SO * SO) / TPP; ++k) {
          const int o = tg + k * TPP;
          const int b = o / SO, Cw = o % SO;
          const int c0 = 2 * Cw, c1 = (2 * Cw + 1) % S, cm = (2 * Cw + S - 1) % S;
          const float* r0 = sK + (2 * b) * SP;
          const float* r1 = sK + (2 * b + 1) * SP;
          const float v = __fadd_rn(__fadd_rn(r0[c0], r1[c0]), __fadd_rn(r0[c1], r1[cm]));
          if (live) atomicAdd(a.outK + obase + o, __fmul_rn(v, a.epi_scale));
          if (NTK) {
            const float* t0 = sU + (2 * b) * SP;
            const float* t1 = sU + (2 * b + 1) * SP;
            const float u = __fadd_rn(__fadd_rn(t0[c0], t1[c0]), __fadd_rn(t0[c1], t1[cm]));
            if (live) atomicAdd(a.outT + obase + o, __fmul_rn(__fsub_rn(u, v), a.epi_scale));
          }
        }
      }
    }
  };

  const int NSTEPS0 = nrows + LAG * (L - 1) + 1;
  const int NSTEPS = NSTEPS0 + (NSTEPS0 & 1);
  for (int t0 = 0; t0 < NSTEPS; t0 += 2) {
    step(t0, std::integral_constant<int, 0>{});
    step(t0 + 1, std::integral_constant<int, 1>{});
  }

  if (EPI == EPI_GAP) {
    float vk = __fadd_rn(gap_k.x, gap_k.y);
    float vu = __fadd_rn(gap_u.x, gap_u.y);
#pragma unroll
    for (int o = LPG / 2; o > 0; o >>= 1) {
      vk = __fadd_rn(vk, __shfl_down_sync(0xffffffffu, vk, o));
      if (NTK) vu = __fadd_rn(vu, __shfl_down_sync(0xffffffffu, vu, o));
    }
    if (TPP > 32) {
      __shared__ float red[2][G::NT / 32];
      if ((tid & 31) == 0) {
        red[0][tid >> 5] = vk;
        red[1][tid >> 5] = vu;
      }
      __syncthreads();
      if (tg == 0) {
        float sk = 0.f, su = 0.f;
        for (int w = 0; w < TPP / 32; ++w) {
          sk = __fadd_rn(sk, red[0][grp * (TPP / 32) + w]);
          su = __fadd_rn(su, red[1][grp * (TPP / 32) + w]);
        }
        vk = sk;
        vu = su;
      }
    }
    if (tg == 0 && live) {
      a.outK[p] = __fmul_rn(vk, a.epi_scale);
      if (NTK) a.outT[p] = __fmul_rn(__fsub_rn(vu, vk), a.epi_scale);
    }
  }
}

}  // namespace ntk
