// Fused "diagonal-marching" kernels: the B200 fast path for stacks of
//   [Conv 3x3 / stride 1 / SAME  ->  ABRelu] x L   (+ AvgPool 2x2/2 | GlobalAvgPool)
// (the Myrtle configurations of BASELINE.json).  Math: SURVEY.md Appendix A,
// reference rules `_src/stax/linear.py:3341-3378` (Conv), `_src/stax/elementwise.py:444-455`
// (ABRelu), `_src/stax/linear.py:3499-3572` (AvgPool), `:1771-1801` (GlobalAvgPool).
//
// Key facts exploited
//  * Every tap of the covariance-space conv moves BOTH members of a pixel pair by the
//    same offset, so (dh, dw) = (h'-h, w'-w) is invariant: the [h,h',w,w'] tensor of one
//    sample pair is a stack of independent 2-D images, each box-filtered 3x3.
//  * "Circular shear" layout: element (h,h',w,w') lives at [ch][h][w][cw] with
//    ch = (h'-h) mod S, cw = (w'-w) mod S.  It is dense (S^4 elements), the innermost
//    index is contiguous, and every conv tap is a pure (h,w) shift at fixed (ch,cw).
//    Zero padding becomes a per-link 0/1 mask (a link is cut where h' or w' wraps).
//  * A thread owns WPT consecutive w at fixed (ch,cw): the w-direction taps are
//    register-local (one shuffle pair per row for the block halo), the h-direction taps
//    are a 2-row sliding window in registers while the thread group marches over h.
//    L layers are software-pipelined with one row of lag each, so L Conv+ABRelu layers
//    cost ZERO HBM traffic; only the pooled (16x smaller) tensor is written.
//  * The first stage computes the input covariance K0 = x1.x2/C on the fly from the
//    two samples held in shared memory (K = C = 3 outer product; nothing materialised).
//
// HBM traffic per pair for Myrtle-10/fp32: 2 x 512 KB + 2 x 32 KB written+read, vs the
// 36.9 MB "one round trip per layer" model used for the roofline.
#pragma once

#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "gemm_kernels.cuh"
#include "generic_kernels.cuh"

namespace ntk {

constexpr int kMaxFusedLayers = 3;
// Rows of lag between consecutive fused layers of k_stage.  1: a layer consumes the row its predecessor
// produced in the same step (fewest live registers).  2: the row of the previous step, so the L layer blocks
// of a step are independent.  Measured on B200 (Myrtle-10 stage 0): fp32 is bound by issue / register-file
// bandwidth and gains nothing from the extra ILP; fp64 is bound by DFMA latency (31 % fixed-latency waits at
// 8 warps per SM) and gains 12 % (134.2 -> 117.9 ms) at 216 registers.
template <typename T>
struct StageLag {
  static constexpr int value = sizeof(T) == 8 ? 2 : 1;
};

enum { IN_FROM_X = 0, IN_LOAD = 1 };
enum { EPI_STORE = 0, EPI_POOL = 1, EPI_GAP = 2 };

// Per-layer constants, already multiplied by the scale of the NEXT conv (alpha = W^2/9),
// so the box filter itself is a plain sum.
template <typename T>
struct FLayer {
  T coef;     // alpha_next * (a-b)^2 / (2 pi)
  T half_ab;  // alpha_next * (a^2+b^2) / 2
  T hab2;     // half_ab - coef * pi/2
  T bias;     // b_std^2 of this layer's conv
  // Erf(a, b, c) (elementwise.py:67-112), used when kind == ACT_ERF.  The q-map of such a layer
  // holds D = 1 + 2 b^2 q and 1/sqrt(D) instead of (q, 1/sqrt q).
  int kind;   // ACT_ABRELU | ACT_ERF
  T e_in;     // 2 b^2
  T eA;       // alpha_next * a^2 * 2/pi
  T eT;       // alpha_next * a^2 b^2 * 4/pi
  T eC;       // alpha_next * c^2
};
// ACT_GELU / ACT_SIN / ACT_RBF (elementwise.py:195-400) reuse the Erf fields as generic parameters:
//   Gelu: eA = alpha_next;   Sin(a, b, c): e_in = b^2, eA = alpha_next a^2 / 2, eT = cos(2c);   Rbf: e_in = gamma, eA = alpha_next
// and their q-maps hold the plain diagonal variance q.
enum { ACT_ABRELU = 0, ACT_ERF = 1, ACT_GELU = 2, ACT_SIN = 3, ACT_RBF = 4 };

template <typename T>
struct StageArgs {
  const T* x1;    // FROM_X: [n1, S, S, C]
  const T* x2;    // FROM_X: [n2, S, S, C]
  const T* inK;   // LOAD: sheared [P, S, S, S, S]
  const T* inT;
  T* outK;        // STORE: sheared [P,S^4]; POOL: sheared [P,(S/2)^4] (pre-zeroed); GAP: [P]
  T* outT;
  const T* qm1;   // [n1][L][S][S][2]  (q, 1/sqrt(q)) per fused layer
  const T* qm2;   // [n2][L][S][S][2]
  long long P;
  int n2;         // pair p -> (p / n2, p % n2), or (p, p) when self
  int self;
  int tri;        // pairs enumerate the upper triangle (j >= i) of an n2 x n2 block, row-major
  int col_start;  // RC: k -- the marched ch columns are 0..k and S-k..S-1 (col_count = 2k+1 of them)
  int col_count;
  int zero_out;   // POOL epilogue: the kernel zeroes its pair's output itself (no host memset)
  T in_scale;     // FROM_X: alpha_1 / C folded into x1
  T epi_scale;    // POOL: alpha_next/16; GAP: 1/S^4; STORE: unused (folded in coef)
  FLayer<T> lp[kMaxFusedLayers];
  // EMB kernels: the RH x RW image sits in the top-left corner of the S x S shear (MNIST 28x28 in 32x32, any
  // H, W <= 32); positions outside are cut off by the link masks and never reach a result.  RH == RW == S otherwise.
  int RH, RW;
  // VALID convs (round 2): the kernels compute the SAME stencil everywhere; the meaningful region after l VALID layers is
  // the box [l, RH0 - l) x [l, RW0 - l) in both members -- its entries read only meaningful entries of the layer before
  // (linear.py:3341-3378 without padding), everything outside is finite don't-care data.  The GAP epilogue sums the box
  // of the stage's last layer: origin box_o (both axes), size box_h x box_w.  SAME networks: box_o = 0, box = RH x RW.
  int box_o, box_h, box_w;
};

// Upper-triangular pair enumeration of a W x W block: row l holds the W - l pairs (l, l..W-1)
// and starts at prefix(l) = l*W - l(l-1)/2.  Returns l and the offset inside the row.
__host__ __device__ __forceinline__ long long tri_prefix(long long l, long long W) {
  return l * W - l * (l - 1) / 2;
}
__device__ __forceinline__ void tri_unrank(long long p, long long W, int& l, int& off) {
  const double b = 2.0 * (double)W + 1.0;
  long long li = (long long)((b - sqrt(b * b - 8.0 * (double)p)) * 0.5);
  if (li < 0) li = 0;
  if (li > W - 1) li = W - 1;
  while (li > 0 && tri_prefix(li, W) > p) --li;
  while (li + 1 < W && tri_prefix(li + 1, W) <= p) ++li;
  l = (int)li;
  off = (int)(p - tri_prefix(li, W));
}

// ---------------------------------------------------------------------------------------
// arithmetic shared by the stage kernel and the q-map kernels (identical rounding, so a
// duplicate pair x1[i] == x2[j] sees q1*q2 - K^2 == 0 exactly; SURVEY §7 "FP32 accuracy").
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }

// 8/16-byte shared-memory load from a 32-bit shared-window address
template <typename T>
__device__ __forceinline__ typename std::conditional<sizeof(T) == 4, float2, double2>::type lds_v2(
    unsigned addr);
template <>
__device__ __forceinline__ float2 lds_v2<float>(unsigned addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
template <>
__device__ __forceinline__ double2 lds_v2<double>(unsigned addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}

// R = P0 + mL*Pm + mR*Pp
template <typename T>
__device__ __forceinline__ T hsum3(T Pm, T P0, T Pp, T mL, T mR) {
  return fma_t(mR, Pp, fma_t(mL, Pm, P0));
}
// C = R0 + vU*Ru + vD*Rd + bias
template <typename T>
__device__ __forceinline__ T vsum3(T Ru, T R0, T Rd, T vU, T vD, T bias) {
  return add_rn(fma_t(vD, Rd, fma_t(vU, Ru, R0)), bias);
}

__device__ __forceinline__ float sqrt_fast(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rsqrt_t(float x) { return rsqrtf(x); }
__device__ __forceinline__ double rsqrt_t(double x) { return 1.0 / sqrt(x); }  // q-maps only (cold)

// acos(c)/sqrt(1-c^2) on [0,1], degree-8 fit (max abs error 1.0e-7 before rounding).
__device__ __forceinline__ float acos_over_sin(float c) {
  float g = 0.017638931050896645f;
  g = __fmaf_rn(g, c, -0.09791343659162521f);
  g = __fmaf_rn(g, c, 0.2531546652317047f);
  g = __fmaf_rn(g, c, -0.42450031638145447f);
  g = __fmaf_rn(g, c, 0.5569935441017151f);
  g = __fmaf_rn(g, c, -0.6610828042030334f);
  g = __fmaf_rn(g, c, 0.7848954796791077f);
  g = __fmaf_rn(g, c, -0.9999822378158569f);
  g = __fmaf_rn(g, c, 1.570796251296997f);
  return g;
}

// ABRelu on one element (elementwise.py:444-455), constants pre-scaled by alpha_next.
//   s = sqrt(max(q1 q2 - K^2, 0)); theta = atan2(s, K);
//   kd = half_ab - coef*theta;  K' = coef*s + kd*K;  T' = kd*T
// fp32: with sin(theta) = s/sqrt(q1 q2) and cos(theta) = K/sqrt(q1 q2) from the stored
// 1/sqrt(q) maps,  acos|cos| = sin * G(|cos|)  (G = acos(c)/sqrt(1-c^2), a degree-8 fit) and
//   kd = hab2 + coef * copysign(pi/2 - acos|cos|, cos),   hab2 = half_ab - coef*pi/2:
// one MUFU (sqrt) per element, no division, no atan2, no branch.
constexpr float kHalfPiF = 1.57079632679489661923f;

// kd of an exact-duplicate element (theta == 0): shared with the q-map kernel so that both
// see bit-identical diagonals.
__device__ __forceinline__ float kd_zero_angle(float coef, float half_ab, float hab2) {
  (void)half_ab;
  return __fmaf_rn(coef, kHalfPiF, hab2);
}
__device__ __forceinline__ double kd_zero_angle(double coef, double half_ab, double hab2) {
  (void)half_ab;
  return __fma_rn(coef, 1.5707963267948966, hab2);
}

__device__ __forceinline__ void act_point(float K, float Tn, float q1, float b1, float q2, float b2,
                                          float coef, float half_ab, float hab2, float& Ko,
                                          float& To) {
  (void)half_ab;
  const float p = __fmul_rn(q1, q2);
  const float rb = __fmul_rn(b1, b2);
  // |.|: a slightly negative difference is rounding noise of the same size as a slightly
  // positive one; exact duplicates give exactly 0 either way.
  const float s = sqrt_fast(fabsf(__fsub_rn(p, __fmul_rn(K, K))));
  const float sn = __fmul_rn(s, rb);
  const float c = __fmul_rn(K, rb);
  const float u = __fmaf_rn(-sn, acos_over_sin(fabsf(c)), kHalfPiF);
  const float kd = __fmaf_rn(coef, copysignf(u, c), hab2);
  Ko = __fmaf_rn(kd, K, __fmul_rn(coef, s));
  To = __fmul_rn(kd, Tn);
}

// fp64: same structure with a degree-16 fit of G (|err| < 7e-14) and a Newton square root
// seeded by MUFU.RSQ64H -- about 45 DP instructions per element instead of ~120 for
// sqrt + atan2.
__device__ __forceinline__ double acos_over_sin(double c) {
  // degree-16 interpolant of G at the Chebyshev nodes of [0,1] (|err| < 7e-14, 3 decades below the
  // 1e-10 budget after 10 layers), split into even and odd parts: two independent Horner chains in
  // c^2 -- 16 FMAs + 1 multiply at a dependent depth of 10.  The fp64 stage kernel is bound by DFMA
  // latency and DP issue, so both the shorter chains and the lower degree pay (ncu: profiles/).
  const double y = __dmul_rn(c, c);
  double e = 0.0006085619653744684;
  e = __fma_rn(e, y, 0.026115987357625138);
  e = __fma_rn(e, y, 0.1515026854745219);
  e = __fma_rn(e, y, 0.32168777107072677);
  e = __fma_rn(e, y, 0.42153491130233556);
  e = __fma_rn(e, y, 0.49055899159932126);
  e = __fma_rn(e, y, 0.5890456114068265);
  e = __fma_rn(e, y, 0.7853981595839817);
  e = __fma_rn(e, y, 1.5707963267948286);
  double o = -0.005808597590519145;
  o = __fma_rn(o, y, -0.07419324994425795);
  o = __fma_rn(o, y, -0.24155677597211828);
  o = __fma_rn(o, y, -0.3804374854745279);
  o = __fma_rn(o, y, -0.45529076094959686);
  o = __fma_rn(o, y, -0.5332956171404206);
  o = __fma_rn(o, y, -0.6666665195234605);
  o = __fma_rn(o, y, -0.9999999999606062);
  return __fma_rn(c, o, e);
}

// sqrt(x) for x >= 0 to ~1 ulp: rsqrt seed + two coupled Newton steps (x == 0 -> ~1e-150).
__device__ __forceinline__ double sqrt_fast(double x) {
  x = fmax(x, 1e-300);
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));  // ~2^-22 relative
  double g = __dmul_rn(x, y), h = __dmul_rn(0.5, y);
  double r = __fma_rn(-g, h, 0.5);  // coupled Newton step: errors ~1e-13
  g = __fma_rn(g, r, g);
  h = __fma_rn(h, r, h);
  r = __fma_rn(-g, h, 0.5);         // final correction of g: below 1 ulp
  return __fma_rn(g, r, g);
}

__device__ __forceinline__ void act_point(double K, double Tn, double q1, double b1, double q2,
                                          double b2, double coef, double half_ab, double hab2,
                                          double& Ko, double& To) {
  (void)half_ab;
  const double p = __dmul_rn(q1, q2);
  const double rb = __dmul_rn(b1, b2);
  const double s = sqrt_fast(fabs(__dsub_rn(p, __dmul_rn(K, K))));
  const double sn = __dmul_rn(s, rb);
  const double c = __dmul_rn(K, rb);
  const double u = __fma_rn(-sn, acos_over_sin(fmin(fabs(c), 1.0)), 1.5707963267948966);
  const double kd = __fma_rn(coef, copysign(u, c), hab2);
  Ko = __fma_rn(kd, K, __dmul_rn(coef, s));
  To = __dmul_rn(kd, Tn);
}

// Erf on one element.  With Kh = 2 b^2 K and D = 1 + 2 b^2 q:
//   s = sqrt(D1 D2 - Kh^2)  (>= 1: the reference's `square_root`),  c = Kh / sqrt(D1 D2),
//   K' = a^2 (2/pi) asin(c) + c_^2,   T' = a^2 b^2 (4/pi) T / s,
// asin(c) = copysign(pi/2 - sqrt(1-c^2) G(|c|), c) with the same fit G as the ABRelu path.
// The q-map kernels call this very function on (K = q, D, D), so a duplicate pair reproduces its
// diagonal bit for bit.
__device__ __forceinline__ float rsqrt_acc(float d, float& s) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  s = __fmul_rn(d, r);
  return r;
}
__device__ __forceinline__ double rsqrt_acc(double x, double& s) {
  // division-free (a DP divide is a subroutine call, which would put the whole kernel on the ABI
  // register budget): rsqrt seed + coupled Newton steps for g ~ sqrt(x), h ~ 1/(2 sqrt(x))
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double g = __dmul_rn(x, y), h = __dmul_rn(0.5, y);
  double r = __fma_rn(-g, h, 0.5);
  g = __fma_rn(g, r, g);
  h = __fma_rn(h, r, h);
  r = __fma_rn(-g, h, 0.5);
  s = __fma_rn(g, r, g);
  return __dmul_rn(2.0, __fma_rn(h, r, h));
}
__device__ __forceinline__ float min1(float c) { return fminf(c, 1.f); }
__device__ __forceinline__ double min1(double c) { return fmin(c, 1.0); }
__device__ __forceinline__ float copysign_t(float a, float b) { return copysignf(a, b); }
__device__ __forceinline__ double copysign_t(double a, double b) { return copysign(a, b); }

template <typename T>
__device__ __forceinline__ void erf_act_point(T K, T Tn, T D1, T rD1, T D2, T rD2, T e_in, T eA, T eT, T eC,
                                              T& Ko, T& To) {
  const T Kh = mul_rn(e_in, K);
  const T p = mul_rn(D1, D2);
  const T rb = mul_rn(rD1, rD2);
  T d = sub_rn(p, mul_rn(Kh, Kh));
  d = d > (T)0.25 ? d : (T)0.25;  // analytically >= 1
  T s;
  const T rs = rsqrt_acc(d, s);
  const T sn = mul_rn(s, rb);
  const T c = mul_rn(Kh, rb);
  const T ac = min1(c < (T)0 ? -c : c);
  const T u = fma_t(-sn, acos_over_sin(ac), (T)1.57079632679489661923);
  Ko = fma_t(eA, copysign_t(u, c), eC);
  To = mul_rn(mul_rn(eT, rs), Tn);
}

__device__ __forceinline__ float exp_g(float x) { return __expf(x); }
__device__ __forceinline__ double exp_g(double x) { return exp(x); }
__device__ __forceinline__ float atan2_g(float y, float x) { return atan2f(y, x); }
__device__ __forceinline__ double atan2_g(double y, double x) { return atan2(y, x); }
__device__ __forceinline__ float sqrt_g(float x) { return sqrtf(x); }
__device__ __forceinline__ double sqrt_g(double x) { return sqrt(x); }

// Gelu / Sin / Rbf on one element (elementwise.py:225-240, 294-300, 375-379), outputs pre-scaled by the next conv's
// alpha (FLayer.eA).  The q-map kernel calls it on (K = q, q, q) for the diagonal, so duplicate pairs agree.
template <typename T>
__device__ __forceinline__ void gen_act_point(int kind, T K, T Tn, T q1, T q2, T e_in, T eA, T eT, T& Ko, T& To) {
  if (kind == ACT_GELU) {
    const T prod = q1 * q2, prod_plus_1 = (q1 + (T)1) * (q2 + (T)1);
    const T delta_squared = prod_plus_1 - K * K;
    const T delta = sqrt_g(delta_squared > (T)0 ? delta_squared : (T)0);
    const T angles = atan2_g(K, delta);
    const T inv_2pi = (T)0.15915494309189533577;
    T nk = (K * K + prod * delta_squared) / (prod_plus_1 * delta);
    nk = (nk + K * angles) * inv_2pi + (T)0.25 * K;
    T first = (T)1 / delta_squared + ((T)1 - prod) / prod_plus_1 + (T)1;
    first *= K / delta * inv_2pi;
    const T dot = first + (T)0.25 + angles * inv_2pi;
    Ko = eA * nk;
    To = eA * dot * Tn;
  } else if (kind == ACT_SIN) {
    const T sum_ = q1 + q2;
    const T s1 = exp_g(e_in * ((T)-0.5 * sum_ + K));
    const T s2 = exp_g(e_in * ((T)-0.5 * sum_ - K)) * eT;
    Ko = eA * (s1 - s2);
    To = eA * e_in * (s1 + s2) * Tn;
  } else {  // ACT_RBF
    const T k = exp_g(e_in * ((T)2 * K - (q1 + q2)));
    Ko = eA * k;
    To = eA * (T)2 * e_in * k * Tn;
  }
}

// Vertical link masks per row r = ch*S + h of a pair (row-major march order):
//   .x = vU: rows r-1 and r are linked (h > 0 and h' = (h+ch) mod S did not wrap),
//   .y = vD: rows r and r+1 are linked.
// Held in constant memory so that one uniform load replaces ~18 uniform-datapath
// instructions per layer and step.  Entry 0 / NR+1.. (clamped rows) reuse valid rows.
// `static`: every translation unit that instantiates the stage kernels owns (and uploads) its copy.
static __constant__ float2 c_vmask32[32 * 32];
static __constant__ float2 c_vmask16[16 * 16];
static __constant__ float2 c_vmask8[8 * 8];

template <int S>
__device__ __forceinline__ float2 vmask_at(int r) {
  if (S == 32) return c_vmask32[r];
  if (S == 16) return c_vmask16[r];
  return c_vmask8[r];
}

template <typename T>
struct Vec2;
template <>
struct Vec2<float> {
  using type = float2;
};
template <>
struct Vec2<double> {
  using type = double2;
};

// ---------------------------------------------------------------------------------------
// q-maps: the per-sample diagonal variances q^l[h,w] = cov^l[h,h,w,w] of every fused layer
// (requirements.py:1057-1117) and their reciprocal square roots.  Between pools the
// diagonal image (ch = cw = 0) evolves on its own: conv -> box filter with border-only
// cuts; ABRelu on the diagonal is q -> (a^2+b^2)/2 q.  One CTA per sample.
//   src_mode 0: diag0[h,w] = in_scale * sum_c x[h,w,c]^2      (FROM_X stages)
//   src_mode 1: diag0[h,w] = selfK[n][ch=0][h][w][cw=0]        (sheared self-pair tensor)
// ---------------------------------------------------------------------------------------
// Sheared input covariance for inputs whose channel count has no FROM_X instantiation (C not in {1, 3}; round 2):
//   out[p][ch][h][w][cw] = sum_c (x1[i, h, w, c] * in_scale) * x2[j, (h + ch) % S, (w + cw) % S, c]     (0 outside RH x RW)
// in the layout the LOADing stage kernels read, with the roundings of the FROM_X kernels and of k_qmaps (requirements.py:
// 542-553; a duplicate pair reproduces its diagonal bit for bit).  One CTA row per (pair, ch, h); a thread owns (w, cw).
// One WARP per row (pair, ch, h): it stages the two image rows in its own slice of shared memory and writes the S x S
// (w, cw) plane (S = 32: one w per iteration, lane = cw, 128-byte stores); the warps of a CTA are independent, so the
// short global reads of one row hide behind the stores of the others (the first version -- one CTA per row, two CTA
// barriers per 4 KB of output -- ran at 1.7 TB/s).
template <typename T, int S>
__global__ void k_input_shear(const T* __restrict__ x1, const T* __restrict__ x2, T* __restrict__ out, long long P,
                              int n2, int self, int tri, int RH, int RW, int C, T in_scale) {
  extern __shared__ __align__(16) unsigned char shear_raw[];
  const int CP = C | 1;  // odd row pitch: the w' = (w + cw) % S gather is bank-conflict free
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  T* xa = reinterpret_cast<T*>(shear_raw) + (size_t)warp * 2 * S * CP;  // [S][CP]  x1 row h, pre-scaled, zero outside
  T* xb = xa + S * CP;                                                  // [S][CP]  x2 row h'
  const long long nrows = P * S * S;
  for (long long row = (long long)blockIdx.x * nwarps + warp; row < nrows; row += (long long)gridDim.x * nwarps) {
    const long long p = row / (S * S);
    const int r = (int)(row - p * (S * S));
    const int ch = r / S, h = r % S;
    const int h2 = (h + ch) % S;
    int si, sj;
    if (self) {
      si = sj = (int)p;
    } else if (tri) {
      int off;
      tri_unrank(p, n2, si, off);
      sj = si + off;
    } else {
      si = (int)(p / n2);
      sj = (int)(p % n2);
    }
    const T* a = x1 + ((long long)si * RH + h) * RW * C;
    const T* b = x2 + ((long long)sj * RH + h2) * RW * C;
    __syncwarp();
    for (int e = lane; e < S * C; e += 32) {
      const int w = e / C, c = e - w * C;
      xa[w * CP + c] = (h < RH && w < RW) ? mul_rn(a[e], in_scale) : (T)0;
      xb[w * CP + c] = (h2 < RH && w < RW) ? b[e] : (T)0;
    }
    __syncwarp();
    T* o = out + row * (long long)(S * S);
    for (int e = lane; e < S * S; e += 32) {
      const int w = e / S, cw = e % S;
      const T* pa = xa + w * CP;
      const T* pb = xb + ((w + cw) % S) * CP;
      T v = mul_rn(pa[0], pb[0]);
      for (int c = 1; c < C; ++c) v = fma_t(pa[c], pb[c], v);
      o[e] = v;
    }
  }
}

template <typename T>
int launch_input_shear(cudaStream_t stream, int64_t* launches, int S, const T* x1, const T* x2, T* out, long long P, int n2,
                       int self, int tri, int RH, int RW, int C, T in_scale) {
  (*launches)++;
  const size_t per_warp = (size_t)2 * S * (C | 1) * sizeof(T);
  int nwarps = 8;
  while (nwarps > 1 && per_warp * nwarps > (size_t)96 * 1024) nwarps >>= 1;
  const size_t smem = per_warp * nwarps;
  if (smem > (size_t)200 * 1024) return fail(NTK_EUNSUPPORTED, "k_input_shear: %d channels do not fit shared memory", C);
  const unsigned grid = (unsigned)std::min<long long>((P * S * S + nwarps - 1) / nwarps, (long long)kNumSMs * 8);
  if (S == 32) {
    NTK_TRY(ensure_dynamic_smem((const void*)k_input_shear<T, 32>, smem));
    k_input_shear<T, 32><<<grid, 32 * nwarps, smem, stream>>>(x1, x2, out, P, n2, self, tri, RH, RW, C, in_scale);
  } else if (S == 16) {
    NTK_TRY(ensure_dynamic_smem((const void*)k_input_shear<T, 16>, smem));
    k_input_shear<T, 16><<<grid, 32 * nwarps, smem, stream>>>(x1, x2, out, P, n2, self, tri, RH, RW, C, in_scale);
  } else {
    NTK_TRY(ensure_dynamic_smem((const void*)k_input_shear<T, 8>, smem));
    k_input_shear<T, 8><<<grid, 32 * nwarps, smem, stream>>>(x1, x2, out, P, n2, self, tri, RH, RW, C, in_scale);
  }
  NTK_CUDA(cudaGetLastError());
  return NTK_OK;
}

// RH x RW: the real image size (<= S; the image sits in the top-left corner of the S x S map, see StageArgs).
template <typename T, int S>
__global__ void k_qmaps(const T* __restrict__ src, int src_mode, int C, T in_scale, int L,
                        FLayer<T> l0, FLayer<T> l1, FLayer<T> l2, T* __restrict__ qm, int RH, int RW) {
  __shared__ T P[S * S];
  __shared__ T R[S * S];
  const int n = blockIdx.x;
  const FLayer<T> lp[3] = {l0, l1, l2};
  for (int e = threadIdx.x; e < S * S; e += blockDim.x) {
    const int h = e / S, w = e % S;
    T v = (T)0;
    if (h < RH && w < RW) {
      if (src_mode == 0) {
        const T* x = src + ((long long)n * RH * RW + h * RW + w) * C;
        v = mul_rn(mul_rn(x[0], in_scale), x[0]);
        for (int c = 1; c < C; ++c) v = fma_t(mul_rn(x[c], in_scale), x[c], v);
      } else {
        v = src[(((long long)n * S + 0) * S + h) * S * S + (long long)w * S + 0];
      }
    }
    P[e] = v;
  }
  __syncthreads();
  typename Vec2<T>::type* out = reinterpret_cast<typename Vec2<T>::type*>(qm) + (long long)n * L * S * S;
  for (int l = 0; l < L; ++l) {
    for (int e = threadIdx.x; e < S * S; e += blockDim.x) {
      const int h = e / S, w = e % S;
      const T mL = (w > 0 && w < RW) ? (T)1 : (T)0, mR = w < RW - 1 ? (T)1 : (T)0;
      R[e] = hsum3<T>(P[h * S + (w > 0 ? w - 1 : w)], P[e], P[h * S + (w < S - 1 ? w + 1 : w)], mL, mR);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < S * S; e += blockDim.x) {
      const int h = e / S, w = e % S;
      const T vU = (h > 0 && h < RH) ? (T)1 : (T)0, vD = h < RH - 1 ? (T)1 : (T)0;
      const T q = vsum3<T>(R[(h > 0 ? h - 1 : h) * S + w], R[e], R[(h < S - 1 ? h + 1 : h) * S + w], vU,
                           vD, lp[l].bias);
      typename Vec2<T>::type o;
      if (lp[l].kind == ACT_ERF) {
        o.x = fma_t(lp[l].e_in, q, (T)1);
        o.y = rsqrt_t(o.x);
        T ko, to;
        erf_act_point<T>(q, (T)0, o.x, o.y, o.x, o.y, lp[l].e_in, lp[l].eA, lp[l].eT, lp[l].eC, ko, to);
        P[e] = ko;
      } else if (lp[l].kind >= ACT_GELU) {
        o.x = q;
        o.y = (T)0;
        T ko, to;
        gen_act_point<T>(lp[l].kind, q, (T)0, q, q, lp[l].e_in, lp[l].eA, lp[l].eT, ko, to);
        P[e] = ko;
      } else {
        o.x = q;
        o.y = q > (T)0 ? rsqrt_t(q) : (T)0;
        P[e] = mul_rn(kd_zero_angle(lp[l].coef, lp[l].half_ab, lp[l].hab2), q);  // theta == 0
      }
      out[(long long)l * S * S + e] = o;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
// The stage kernel.
//   S    spatial size (32, 16, 8);  WPT  consecutive w per thread;  L fused layers (1..3)
//   CIN  input channels for FROM_X (x2 is padded to 4 in shared memory)
// A *group* of TPP = S*S/WPT threads owns one sample pair and marches over its S*S rows
// (ch-major: row r = ch*S + h).  NT threads per CTA hold NT/TPP groups.
// ---------------------------------------------------------------------------------------
template <int S, int WPT, int SH = 1>
struct StageGeom {
  static constexpr int TPP = S * S / WPT;               // threads per plane == per group
  static constexpr int NT = TPP < 128 ? 128 : TPP * SH; // threads per CTA
  static constexpr int GROUPS = NT / TPP;
  static constexpr int LPG = TPP < 32 ? TPP : 32;       // lanes of one group inside a warp
  static constexpr int NWB = S / WPT;                   // w-blocks
  static constexpr int LW = LPG / NWB;                  // cw lanes per w-block inside a warp
  static constexpr int NR = S * S;                      // rows per pair
};

// SH > 1: the SH groups of a CTA work on SH consecutive columns of the same Gram row, so the
// row sample x1[i] and its q-maps are staged once and shared (more resident warps per SM).
// RC ("runtime columns"): march only the ch columns [col_start, col_start + col_count) -- used by
// the self-pair pipeline; the cross-pair kernels keep compile-time trip counts.
// ERF: 0 = pure ABRelu (no other code compiled in, the hot path); 1 = the stage may contain Erf layers (runtime branch
// per layer); 2 = any of ABRelu / Erf / Gelu / Sin / Rbf (the general family, fused_*_gen.cu).
// EMB: the image is RH x RW <= S x S (StageArgs): staging, link masks and the GAP epilogue use the real size and
// the vertical masks are computed instead of read from the constant tables (which hold RH == S).
template <typename T, int S, int WPT, int L, int IN, int EPI, bool NTK, int CIN, int SH, bool RC, int ERF,
          bool EMB = false>
__global__ void __launch_bounds__(StageGeom<S, WPT, SH>::NT)
k_stage(const StageArgs<T> a) {
  using G = StageGeom<S, WPT, SH>;
  using V2 = typename Vec2<T>::type;
  constexpr int TPP = G::TPP, LPG = G::LPG, NWB = G::NWB, LW = G::LW, NR = G::NR;
  constexpr int LAG = StageLag<T>::value;
  constexpr int SO = S / 2;
  constexpr int SP = S + 1;  // padded staging row

  extern __shared__ __align__(16) unsigned char smem_raw[];
  // per-group shared memory carve-up
  constexpr int XS1 = IN == IN_FROM_X ? S * S * CIN : 0;
  constexpr int XS2 = IN == IN_FROM_X ? S * S * 4 : 0;
  constexpr int QM = L * S * S * 2;
  constexpr int STG = EPI == EPI_POOL ? 2 * (NTK ? 2 : 1) * S * SP : 0;
  constexpr bool SHARED_ROW = SH > 1;
  constexpr int ROW_PART = XS1 + QM;        // x1 sample + its q-maps
  constexpr int COL_PART = XS2 + QM + STG;  // x2 sample + its q-maps + pool staging
  const int tid = threadIdx.x;
  const int grp = tid / TPP, tg = tid % TPP;
  T* sm_row = reinterpret_cast<T*>(smem_raw) + (SHARED_ROW ? 0 : (size_t)grp * (ROW_PART + COL_PART));
  T* sm_col = reinterpret_cast<T*>(smem_raw) +
              (SHARED_ROW ? (size_t)ROW_PART + (size_t)grp * COL_PART
                          : (size_t)grp * (ROW_PART + COL_PART) + ROW_PART);
  T* x1s = sm_row;
  V2* q1m = reinterpret_cast<V2*>(x1s + XS1);
  T* x2s = sm_col;
  V2* q2m = reinterpret_cast<V2*>(x2s + XS2);
  T* stg = reinterpret_cast<T*>(q2m + L * S * S);

  const int lig = tg % LPG, wig = tg / 32;
  const int wblk = lig / LW, cwsub = lig % LW;
  const int cw = wig * LW + cwsub;
  const int w0 = wblk * WPT;

  long long p;
  bool live;
  int si, sj;
  if (SHARED_ROW) {
    const int bpr = (a.n2 + SH - 1) / SH;  // CTAs per Gram row
    si = blockIdx.x / bpr;
    sj = (blockIdx.x % bpr) * SH + grp;
    live = sj < a.n2;
    if (!live) sj = a.n2 - 1;
    p = (long long)si * a.n2 + sj;
  } else {
    p = (long long)blockIdx.x * G::GROUPS + grp;
    live = p < a.P;
    if (!live) p = a.P - 1;
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
  }

  // ---- stage the two samples and their q-maps in shared memory -------------------------
  {
    // the row part is loaded by the whole CTA when shared, else by its group
    const int rt = SHARED_ROW ? tid : tg, rn = SHARED_ROW ? G::NT : TPP;
    if (IN == IN_FROM_X && !EMB) {
      const T* g1 = a.x1 + (long long)si * S * S * CIN;
      const T* g2 = a.x2 + (long long)sj * S * S * CIN;
      for (int e = rt; e < S * S * CIN; e += rn) x1s[e] = mul_rn(g1[e], a.in_scale);
      for (int e = tg; e < S * S; e += TPP) {
#pragma unroll
        for (int c = 0; c < 4; ++c) x2s[e * 4 + c] = c < CIN ? g2[e * CIN + c] : (T)0;
      }
    }
    if (IN == IN_FROM_X && EMB) {  // [RH, RW, CIN] samples, zero outside
      const T* g1 = a.x1 + (long long)si * a.RH * a.RW * CIN;
      const T* g2 = a.x2 + (long long)sj * a.RH * a.RW * CIN;
      for (int e = rt; e < S * S; e += rn) {
        const int h = e / S, w = e % S;
        const bool in = h < a.RH && w < a.RW;
#pragma unroll
        for (int c = 0; c < CIN; ++c) x1s[e * CIN + c] = in ? mul_rn(g1[(h * a.RW + w) * CIN + c], a.in_scale) : (T)0;
      }
      for (int e = tg; e < S * S; e += TPP) {
        const int h = e / S, w = e % S;
        const bool in = h < a.RH && w < a.RW;
#pragma unroll
        for (int c = 0; c < 4; ++c) x2s[e * 4 + c] = (c < CIN && in) ? g2[(h * a.RW + w) * CIN + c] : (T)0;
      }
    }
    const V2* g1 = reinterpret_cast<const V2*>(a.qm1) + (long long)si * L * S * S;
    const V2* g2 = reinterpret_cast<const V2*>(a.qm2) + (long long)sj * L * S * S;
    if (TPP >= 128) {
      // q-maps are contiguous 16-byte-aligned blocks: one elected thread hands both copies to
      // the TMA bulk-copy engine (cp.async.bulk, SASS UBLKCP) and the group waits on an mbarrier.
      __shared__ __align__(8) unsigned long long qbar[G::GROUPS];
      constexpr unsigned kBytes = (unsigned)(L * S * S * sizeof(V2));
      const unsigned bar = (unsigned)__cvta_generic_to_shared(&qbar[grp]);
      const bool copy_row = !SHARED_ROW || grp == 0;
      if (tg == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
                     "r"(copy_row ? 2 * kBytes : kBytes));
        if (copy_row)
          asm volatile(
              "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                  (unsigned)__cvta_generic_to_shared(q1m)),
              "l"(g1), "r"(kBytes), "r"(bar)
              : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                (unsigned)__cvta_generic_to_shared(q2m)),
            "l"(g2), "r"(kBytes), "r"(bar)
            : "memory");
      }
      __syncthreads();  // barrier initialised (and, for SHARED_ROW, group 0 issued the row copy)
      unsigned done = 0;
      while (!done) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(bar)
            : "memory");
      }
    } else {
      for (int e = rt; e < L * S * S; e += rn) q1m[e] = g1[e];
      for (int e = tg; e < L * S * S; e += TPP) q2m[e] = g2[e];
    }
  }
  if (TPP > 32)
    __syncthreads();
  else
    __syncwarp();

  // ---- per-thread constants -------------------------------------------------------------
  // lk[i]: link between w0+i-1 and w0+i is intact (both inside the image, w' does not wrap)
  T lk[WPT + 1];
  int off2[WPT];  // (w0 + i + cw) mod S (column of the second member), as a byte offset into a V2 row
  const int RWc = EMB ? a.RW : S, RHc = EMB ? a.RH : S;
#pragma unroll
  for (int i = 0; i <= WPT; ++i) {
    const int wl = w0 + i - 1, wr = w0 + i;
    lk[i] = (wl >= 0 && wr <= RWc - 1 && ((wl + cw) % S) + 1 <= RWc - 1) ? (T)1 : (T)0;
  }
  bool inw[WPT];  // EMB + GAP: element (w0 + i, w' = w0 + i + cw) lies inside the box of the last layer
#pragma unroll
  for (int i = 0; i < WPT; ++i) {
    const int wa = w0 + i - a.box_o, wb = ((w0 + i + cw) % S) - a.box_o;
    inw[i] = !EMB || (wa >= 0 && wa < a.box_w && wb >= 0 && wb < a.box_w);
  }
#pragma unroll
  for (int i = 0; i < WPT; ++i) off2[i] = ((w0 + i + cw) % S) * (int)sizeof(V2);
  const unsigned q2base = (unsigned)__cvta_generic_to_shared(q2m);

  // sliding windows: RK[l][slot][i] = horizontally summed input row of layer l.
  // The pipeline runs unconditionally: rows outside [0, NR) are computed from clamped,
  // finite data and never reach a result (their links are cut: vU = 0 at h = 0, vD = 0 at
  // h = S-1), so the rings need no fill/drain special cases.
  T RK[L][2][WPT];
  T RT[L][2][WPT];
#pragma unroll
  for (int l = 0; l < L; ++l)
#pragma unroll
    for (int sl = 0; sl < 2; ++sl)
#pragma unroll
      for (int i = 0; i < WPT; ++i) {
        RK[l][sl][i] = (T)0;
        RT[l][sl][i] = (T)0;
      }

  // Rows marched: nrows = col_count * S; marched row r -> column ch (see full_row below),
  // full row index rf = ch*S + h (position in the sheared tensor and in the mask table).
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

  T gap_k = (T)0, gap_t = (T)0;
  const T* inK = IN == IN_LOAD ? a.inK + p * (long long)NR * S * S + (long long)w0 * S + cw : nullptr;
  // inT == nullptr on a LOADing ntk stage: the input is the sheared input covariance of k_input_shear (ntk = 0)
  const bool has_inT = IN == IN_LOAD && NTK && a.inT != nullptr;
  const T* inT = has_inT ? a.inT + p * (long long)NR * S * S + (long long)w0 * S + cw : nullptr;

  // next input row (software prefetch for LOAD)
  T nK[WPT], nT[WPT];
  auto fetch = [&](int r) {
    if (IN == IN_LOAD) {
      const int rc = full_row(r < nrows ? r : nrows - 1);
      const T* gk = inK + (long long)rc * S * S;
#pragma unroll
      for (int i = 0; i < WPT; ++i) nK[i] = __ldg(gk + i * S);
      if (NTK) {
        if (has_inT) {
          const T* gt = inT + (long long)rc * S * S;
#pragma unroll
          for (int i = 0; i < WPT; ++i) nT[i] = __ldg(gt + i * S);
        } else {
#pragma unroll
          for (int i = 0; i < WPT; ++i) nT[i] = (T)0;
        }
      }
    }
  };
  fetch(0);

  // Layer-to-layer hand-over rows.  With LAG == 2 a layer consumes the row its predecessor
  // produced in the PREVIOUS step, so the L layer blocks of one step are independent and the
  // scheduler can overlap their latencies (shuffle -> taps -> sqrt -> polynomial chains).
  T BK[L][WPT], BT[L][WPT];
#pragma unroll
  for (int l = 0; l < L; ++l)
#pragma unroll
    for (int i = 0; i < WPT; ++i) {
      BK[l][i] = (T)0;
      BT[l][i] = (T)0;
    }

  auto step = [&](const int t, auto par_c) {
    // `par` = t & 1 as a compile-time constant: the ring slot of each layer alternates.
    constexpr int par = decltype(par_c)::value;
    T PK[WPT], PT[WPT];
    // ---- input row r = t of layer 1 ----------------------------------------------------
    if (IN == IN_FROM_X) {
      const int r = full_row(t < nrows ? t : nrows - 1);
      const int ch = r / S, h = r % S;
      const int h2 = (h + ch) % S;
      const T* xa = x1s + (h * S + w0) * CIN;
      const char* xb = reinterpret_cast<const char*>(x2s + h2 * S * 4);
#pragma unroll
      for (int i = 0; i < WPT; ++i) {
        const T* b4 = reinterpret_cast<const T*>(xb + off2[i] * 2);  // 4 T per pixel = 2 V2
        T acc = mul_rn(xa[i * CIN], b4[0]);
#pragma unroll
        for (int c = 1; c < CIN; ++c) acc = fma_t(xa[i * CIN + c], b4[c], acc);
        PK[i] = acc;
      }
    } else {
#pragma unroll
      for (int i = 0; i < WPT; ++i) {
        PK[i] = nK[i];
        PT[i] = nT[i];
      }
      fetch(t + 1);
    }
#pragma unroll
    for (int li = 0; li < L; ++li) {
      // LAG == 2: last layer first, so BK[l-1] still holds the previous step's row
      const int l = LAG == 2 ? L - 1 - li : li;
      const int lm = l == 0 ? 0 : l - 1;
      const int slot = (par + (LAG == 1 ? l : 0)) & 1;  // == (t - LAG*l) & 1, compile-time
      const bool has_t = NTK && (l > 0 || IN == IN_LOAD);
#define INK(i) (l == 0 ? PK[i] : BK[lm][i])
#define INT(i) (l == 0 ? PT[i] : BT[lm][i])
      // ---- conv row r_out = t - LAG*l - 1 (clamped outside the image) ---------------------
      int r_out = t - LAG * l - 1;
      r_out = full_row(r_out < 0 ? 0 : (r_out > nrows - 1 ? nrows - 1 : r_out));
      const int ch = r_out / S, h = r_out % S;
      const int h2 = (h + ch) % S;
      T vU, vD;
      if (EMB) {
        vU = (h > 0 && h2 > 0 && h < RHc && h2 < RHc) ? (T)1 : (T)0;
        vD = (h + 1 < RHc && h2 + 1 < RHc) ? (T)1 : (T)0;
      } else {
        const float2 vm = vmask_at<S>(r_out);
        vU = (T)vm.x;
        vD = (T)vm.y;
      }
      // ---- vertical taps of the two OLD rows first (last use of the oldest row) ----------
      T tk[WPT], tt[WPT];
#pragma unroll
      for (int i = 0; i < WPT; ++i) {
        tk[i] = fma_t(vU, RK[l][slot][i], RK[l][slot ^ 1][i]);
        if (has_t) tt[i] = fma_t(vU, RT[l][slot][i], RT[l][slot ^ 1][i]);
      }
      // ---- horizontal 3-tap of the new row, written straight into the freed ring slot ----
      {
        T left = (T)0, right = (T)0, leftT = (T)0, rightT = (T)0;
        if (NWB > 1) {
          left = __shfl_up_sync(0xffffffffu, INK(WPT - 1), LW);
          right = __shfl_down_sync(0xffffffffu, INK(0), LW);
          if (has_t) {
            leftT = __shfl_up_sync(0xffffffffu, INT(WPT - 1), LW);
            rightT = __shfl_down_sync(0xffffffffu, INT(0), LW);
          }
        }
#pragma unroll
        for (int i = 0; i < WPT; ++i) {
          RK[l][slot][i] = hsum3<T>(i == 0 ? left : INK(i == 0 ? 0 : i - 1), INK(i),
                                    i == WPT - 1 ? right : INK(i == WPT - 1 ? i : i + 1), lk[i],
                                    lk[i + 1]);
          if (has_t)
            RT[l][slot][i] = hsum3<T>(i == 0 ? leftT : INT(i == 0 ? 0 : i - 1), INT(i),
                                      i == WPT - 1 ? rightT : INT(i == WPT - 1 ? i : i + 1), lk[i],
                                      lk[i + 1]);
        }
      }
      // ---- finish the vertical sum, add the bias, apply the activation ---------------------
      {
        const V2* q1r = q1m + (l * S + h) * S + w0;
        const unsigned q2row = q2base + (unsigned)((l * S + h2) * S * (int)sizeof(V2));
        const T coef = a.lp[l].coef, half_ab = a.lp[l].half_ab, hab2 = a.lp[l].hab2,
                bias = a.lp[l].bias;
#pragma unroll
        for (int i = 0; i < WPT; ++i) {
          const T ck = add_rn(fma_t(vD, RK[l][slot][i], tk[i]), bias);
          T ct = (T)0;
          if (NTK) {
            // linear.py:1396-1398 (T0 == 0 for the first layer of a FROM_X stage)
            ct = has_t ? add_rn(fma_t(vD, RT[l][slot][i], tt[i]), ck) : ck;
          }
          const V2 qa = q1r[i];
          const V2 qb = lds_v2<T>(q2row + off2[i]);
          if (ERF == 2 && a.lp[l].kind >= ACT_GELU)
            gen_act_point<T>(a.lp[l].kind, ck, ct, qa.x, qb.x, a.lp[l].e_in, a.lp[l].eA, a.lp[l].eT, BK[l][i], BT[l][i]);
          else if (ERF && a.lp[l].kind == ACT_ERF)
            erf_act_point<T>(ck, ct, qa.x, qa.y, qb.x, qb.y, a.lp[l].e_in, a.lp[l].eA, a.lp[l].eT, a.lp[l].eC,
                             BK[l][i], BT[l][i]);
          else
            act_point(ck, ct, qa.x, qa.y, qb.x, qb.y, coef, half_ab, hab2, BK[l][i], BT[l][i]);
        }
      }
#undef INK
#undef INT
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
          for (int i = 0; i < WPT; ++i) {
            a.outK[base + (long long)i * S] = BK[L - 1][i];
            if (NTK) a.outT[base + (long long)i * S] = BT[L - 1][i];
          }
        }
      } else if (EPI == EPI_GAP) {
        const int ha = h - a.box_o, hb = ((h + ch) % S) - a.box_o;
        const bool row_in = !EMB || (ha >= 0 && ha < a.box_h && hb >= 0 && hb < a.box_h);
#pragma unroll
        for (int i = 0; i < WPT; ++i) {
          if (EMB && !(row_in && inw[i])) continue;
          gap_k = add_rn(gap_k, BK[L - 1][i]);
          if (NTK) gap_t = add_rn(gap_t, BT[L - 1][i]);
        }
      } else {  // EPI_POOL: AvgPool 2x2/2 of both members (linear.py:3499-3572)
        T* sK = stg + (r_fin & 1) * ((NTK ? 2 : 1) * S * SP);
        T* sT = sK + S * SP;
#pragma unroll
        for (int i = 0; i < WPT; ++i) {
          sK[(w0 + i) * SP + cw] = BK[L - 1][i];
          if (NTK) sT[(w0 + i) * SP + cw] = BT[L - 1][i];
        }
        if (TPP > 32)
          __syncthreads();
        else
          __syncwarp();
        // this row feeds pooled row a = h/2 of column Ch (see DESIGN.md "pooling in the shear")
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
          const T* r0 = sK + (2 * b) * SP;
          const T* r1 = sK + (2 * b + 1) * SP;
          T v = add_rn(add_rn(r0[c0], r1[c0]), add_rn(r0[c1], r1[cm]));
          if (live) atomicAdd(a.outK + obase + o, mul_rn(v, a.epi_scale));
          if (NTK) {
            const T* t0 = sT + (2 * b) * SP;
            const T* t1 = sT + (2 * b + 1) * SP;
            T u = add_rn(add_rn(t0[c0], t1[c0]), add_rn(t0[c1], t1[cm]));
            if (live) atomicAdd(a.outT + obase + o, mul_rn(u, a.epi_scale));
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
    // deterministic group reduction: lanes of the group inside each warp, then warps
    T vk = gap_k, vt = gap_t;
#pragma unroll
    for (int o = LPG / 2; o > 0; o >>= 1) {
      vk = add_rn(vk, __shfl_down_sync(0xffffffffu, vk, o));
      if (NTK) vt = add_rn(vt, __shfl_down_sync(0xffffffffu, vt, o));
    }
    if (TPP > 32) {
      __shared__ T red[2][G::NT / 32];
      if ((tid & 31) == 0) {
        red[0][tid >> 5] = vk;
        red[1][tid >> 5] = vt;
      }
      __syncthreads();
      if (tg == 0) {
        T sk = (T)0, st = (T)0;
        for (int w = 0; w < TPP / 32; ++w) {
          sk = add_rn(sk, red[0][grp * (TPP / 32) + w]);
          st = add_rn(st, red[1][grp * (TPP / 32) + w]);
        }
        vk = sk;
        vt = st;
      }
    }
    if (tg == 0 && live) {
      a.outK[p] = mul_rn(vk, a.epi_scale);
      if (NTK) a.outT[p] = mul_rn(vt, a.epi_scale);
    }
  }
}

// ---------------------------------------------------------------------------------------
// sheared <-> canonical layout conversion (only when a fused segment meets the general path)
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void k_shear(const T* __restrict__ in, T* __restrict__ out, long long P, int S, int to_shear) {
  const long long per = (long long)S * S * S * S;
  const long long total = P * per;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long p = idx / per;
    int r = (int)(idx % per);
    // idx enumerates the sheared tensor [ch][h][w][cw]
    int cw = r % S;
    r /= S;
    int w = r % S;
    r /= S;
    int h = r % S;
    int ch = r / S;
    int h2 = (h + ch) % S, w2 = (w + cw) % S;
    long long can = p * per + (((long long)h * S + h2) * S + w) * S + w2;
    if (to_shear)
      out[idx] = in[can];
    else
      out[can] = in[idx];
  }
}

// triangular tile result -> matrix: pair p = (l, l + off) of the block starting at (r0, r0)
template <typename T>
__global__ void k_scatter_tri(const T* __restrict__ src, T* __restrict__ dst, long long P, int W,
                              long long ld, int r0) {
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < P;
       p += (long long)gridDim.x * blockDim.x) {
    int l, off;
    tri_unrank(p, W, l, off);
    dst[(long long)(r0 + l) * ld + r0 + l + off] = src[p];
  }
}

// m[i, j] = m[j, i] for j < i (fills the strictly lower triangle from the upper one)
template <typename T>
__global__ void k_mirror(T* __restrict__ m, int n, long long ld) {
  const long long total = (long long)n * n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx / n), j = (int)(idx % n);
    if (j < i) m[(long long)i * ld + j] = m[(long long)j * ld + i];
  }
}

// ---------------------------------------------------------------------------------------
// host side: plan + driver
// ---------------------------------------------------------------------------------------
struct FusedStage {
  int L = 0;
  double w2[kMaxFusedLayers], b2[kMaxFusedLayers], a[kMaxFusedLayers], b[kMaxFusedLayers];
  double c[kMaxFusedLayers];          // Erf only
  int kind[kMaxFusedLayers] = {0};    // ACT_ABRELU | ACT_ERF
  int epi = EPI_STORE;  // what follows this chunk
  // the pool behind this chunk (epi == EPI_POOL): SAME padding (== VALID on even sizes only) and
  // SumPool (no division by the window, `_src/stax/linear.py:1503`: x16 for a 2x2 window on both members)
  bool pool_same = false;
  double pool_mul = 1.0;
};

struct FusedPlan {
  bool ok = false;
  std::vector<FusedStage> stages;    // chunks of <= 3 layers; resolution halves after EPI_POOL
  std::vector<ntk_op_t> dense_tail;  // Dense ops applied to the [n1,n2] result
  // tail = `tail_pools` 2x2/2 pools, then GlobalAvgPool / GlobalSumPool / Flatten: a global mean times `tail_mul`
  // (x16 per SumPool; GlobalSumPool multiplies by the final map size at run time)
  int tail_pools = 0;
  bool tail_flatten = false, tail_gsum = false, tail_same = false;
  double tail_mul = 1.0;
  bool valid_convs = false;  // every conv is 3x3 / 1 / VALID (all SAME otherwise; mixtures are not planned)
};

// Meaningful region of the maps: origin `o` (the same on both axes) and size h x w, in the coordinates of the stage's shear.
struct FusedBox {
  int o, h, w;
};

// Walks the boxes through the stages.  in_box[s]: region entering stage s; *last: region after the last stage's layers.
// false: a map would be empty, or a pool would not be aligned with the box (odd origin).
inline bool fused_walk_boxes(const FusedPlan& plan, int H, int W, std::vector<FusedBox>* in_box, FusedBox* last) {
  FusedBox b{0, H, W};
  for (size_t s = 0; s + 1 < plan.stages.size(); ++s) {
    const FusedStage& st = plan.stages[s];
    if (in_box) in_box->push_back(b);
    if (plan.valid_convs) {
      b.o += st.L;
      b.h -= 2 * st.L;
      b.w -= 2 * st.L;
      if (b.h < 1 || b.w < 1) return false;
    }
    if (st.epi == EPI_POOL) {
      if (b.o & 1) return false;  // the kernels pool rows / columns (2b, 2b + 1) of the shear
      if (b.h < 2 || b.w < 2) return false;
      b = FusedBox{b.o / 2, b.h / 2, b.w / 2};
    }
  }
  if (last) *last = b;
  return true;
}

// AvgPool / SumPool op flags: i[5] bit 0 = normalize_edges, bit 1 = SumPool; GAP: i[0] = 1 for GlobalSumPool
inline bool pool_is_sum(const ntk_op_t& o) { return (o.i[5] & 2) != 0; }
inline bool pool_normalize_edges(const ntk_op_t& o) { return (o.i[5] & 1) != 0; }

// Recognises:  ([Conv3x3/1/SAME, ABRelu]+  AvgPool2x2/2?)+  (AvgPool2x2/2)* (GAP | Flatten@1x1)  Dense*
inline FusedPlan plan_fused(const std::vector<ntk_op_t>& ops, int n_slots, int out_slot,
                            int max_layers = kMaxFusedLayers) {
  FusedPlan plan;
  const int n = (int)ops.size();
  // must be a linear chain ending in out_slot
  for (int k = 0; k < n; ++k) {
    if (ops[k].kind == NTK_OP_FANINSUM) return plan;
    if (ops[k].src != (k == 0 ? 0 : ops[k - 1].dst)) return plan;
  }
  if (n == 0 || ops[n - 1].dst != out_slot) return plan;
  int pad0 = -1;  // padding of the first conv; every conv must use it
  for (int k = 0; k < n; ++k)
    if (ops[k].kind == NTK_OP_CONV) {
      pad0 = ops[k].i[4];
      break;
    }
  if (pad0 != NTK_PAD_SAME && pad0 != NTK_PAD_VALID) return plan;
  plan.valid_convs = pad0 == NTK_PAD_VALID;
  auto is_conv = [pad0](const ntk_op_t& o) {
    return o.kind == NTK_OP_CONV && o.i[0] == 3 && o.i[1] == 3 && o.i[2] == 1 && o.i[3] == 1 && o.i[4] == pad0;
  };
  auto is_act = [](const ntk_op_t& o) {
    return (o.kind == NTK_OP_ABRELU && o.i[0] == 0) || o.kind == NTK_OP_ERF || o.kind == NTK_OP_GELU ||
           o.kind == NTK_OP_SIN || o.kind == NTK_OP_RBF;
  };
  auto is_pool = [](const ntk_op_t& o) {
    return o.kind == NTK_OP_AVGPOOL && o.i[0] == 2 && o.i[1] == 2 && o.i[2] == 2 && o.i[3] == 2 &&
           o.i[4] != NTK_PAD_CIRCULAR && !(o.i[4] == NTK_PAD_SAME && pool_normalize_edges(o) && !pool_is_sum(o));
  };
  int k = 0;
  int pending_pools = 0;  // pools seen after the last layer run
  bool pend_same = false;
  double pend_mul = 1.0;
  while (k < n) {
    if (!(is_conv(ops[k]) && k + 1 < n && is_act(ops[k + 1]))) break;
    if (!plan.stages.empty()) {
      if (pending_pools > 1) return FusedPlan();
      plan.stages.back().epi = pending_pools == 1 ? EPI_POOL : EPI_STORE;
      plan.stages.back().pool_same = pend_same;
      plan.stages.back().pool_mul = pend_mul;
    }
    // gather the run of conv+act layers and cut it into chunks of <= kMaxFusedLayers
    std::vector<std::pair<int, int>> run;
    while (k + 1 < n && is_conv(ops[k]) && is_act(ops[k + 1])) {
      run.push_back({k, k + 1});
      k += 2;
    }
    for (size_t s = 0; s < run.size(); s += max_layers) {
      FusedStage st;
      st.L = (int)std::min<size_t>(max_layers, run.size() - s);
      for (int l = 0; l < st.L; ++l) {
        const ntk_op_t& c = ops[run[s + l].first];
        const ntk_op_t& act = ops[run[s + l].second];
        st.w2[l] = c.f[0];
        st.b2[l] = c.i[5] ? c.f[1] : 0.0;
        st.a[l] = act.f[0];
        st.b[l] = act.f[1];
        st.c[l] = (act.kind == NTK_OP_ERF || act.kind == NTK_OP_SIN) ? act.f[2] : 0.0;
        st.kind[l] = act.kind == NTK_OP_ERF ? ACT_ERF : act.kind == NTK_OP_GELU ? ACT_GELU
                     : act.kind == NTK_OP_SIN ? ACT_SIN : act.kind == NTK_OP_RBF ? ACT_RBF : ACT_ABRELU;
      }
      st.epi = EPI_STORE;
      plan.stages.push_back(st);
    }
    pending_pools = 0;
    pend_same = false;
    pend_mul = 1.0;
    while (k < n && is_pool(ops[k])) {
      ++pending_pools;
      pend_same = pend_same || ops[k].i[4] == NTK_PAD_SAME;
      if (pool_is_sum(ops[k])) pend_mul *= 16.0;
      ++k;
    }
  }
  if (plan.stages.empty() || k >= n) return FusedPlan();
  if (ops[k].kind != NTK_OP_GAP && ops[k].kind != NTK_OP_FLATTEN) return FusedPlan();
  // tail: `pending_pools` 2x2/2 pools then GAP (always a global mean) or Flatten (a global
  // mean iff the pools reduced the map to 1x1 -- checked against S at run time).
  plan.stages.back().epi = EPI_GAP;
  const bool flatten = ops[k].kind == NTK_OP_FLATTEN;
  ++k;
  for (; k < n; ++k) {
    if (ops[k].kind != NTK_OP_DENSE) return FusedPlan();
    plan.dense_tail.push_back(ops[k]);
  }
  plan.ok = true;
  plan.tail_pools = pending_pools;
  plan.tail_flatten = flatten;
  plan.tail_gsum = !flatten && ops[k - 1 - (int)plan.dense_tail.size()].i[0] == 1;
  plan.tail_same = pend_same;
  plan.tail_mul = pend_mul;
  // L == 0 marker stage closes the list (keeps `stages.size() - 1` == number of real stages)
  FusedStage tail;
  tail.L = 0;
  plan.stages.push_back(tail);
  return plan;
}

// Shear size of an H x W image: the smallest instantiated S in {8, 16, 32} that holds it (0: none).
inline int fused_shear_size(int H, int W) {
  const int m = H > W ? H : W;
  return m <= 8 ? 8 : (m <= 16 ? 16 : (m <= 32 ? 32 : 0));
}

// Images that are not S x S RGB run the EMB family of the scalar stage kernels (pure ABRelu stages only).
// Other channel counts (round 2) go through k_input_shear and LOADing first stages, which exist in every family.
// Which kernel family and input mode a plan runs with.  FROM_X stages exist for C = 3 (every family) and C = 1 (EMB family).
//   emb      images that are not a shear size, VALID convs (box-aware epilogues), and grey inputs of pure-ABRelu networks
//            (measured on B200, grey 32 x 32 Myrtle-10 / Myrtle-5: EMB FROM_X 181 k / 299 k entries/s against 116 k / 142 k
//            through the round-2 pre-pass + packed kernels, `profiles/grey_ab_r02.log`)
//   prepass  every other channel count: k_input_shear + LOADing first stage; also grey shear-size inputs of networks the
//            ABRelu-only EMB family cannot run (Erf, Gelu, ...)
struct FusedRoute {
  bool emb, prepass;
};
inline FusedRoute fused_route(const FusedPlan& plan, int H, int W, int C) {
  const bool square = H == W && (H == 32 || H == 16 || H == 8);
  bool pure_abrelu = true;
  for (size_t s = 0; s + 1 < plan.stages.size(); ++s)
    for (int l = 0; l < plan.stages[s].L; ++l) pure_abrelu = pure_abrelu && plan.stages[s].kind[l] == ACT_ABRELU;
  FusedRoute r;
  r.emb = !square || plan.valid_convs || (C == 1 && pure_abrelu);
  r.prepass = r.emb ? (C != 1 && C != 3) : C != 3;
  return r;
}

template <typename T>
bool fused_supported(const FusedPlan& plan, int H, int W, int C) {
  if (!plan.ok || H < 1 || W < 1) return false;
  int S = fused_shear_size(H, W);
  if (S == 0) return false;
  if (C < 1) return false;  // C = 1, 3: FROM_X stages; any other C: k_input_shear + LOADing first stage
  // the pre-pass stages two image rows per warp in shared memory (launch_input_shear); sized for float64 so that both dtypes
  // of a program land on the same path
  if (fused_route(plan, H, W, C).prepass && (size_t)2 * S * (C | 1) * sizeof(double) > (size_t)200 * 1024) return false;
  FusedBox last{0, H, W};
  std::vector<FusedBox> in_box;
  if (!fused_walk_boxes(plan, H, W, &in_box, &last)) return false;
  for (size_t s = 0; s + 1 < plan.stages.size(); ++s) {
    const FusedStage& st = plan.stages[s];
    if (st.epi == EPI_POOL) {
      if (S <= 8) return false;  // S = 4 stages are not instantiated
      // VALID drops an odd last row (floor); SAME pads it with zeros the garbage outside the image cannot provide
      const int bh = in_box[s].h - (plan.valid_convs ? 2 * st.L : 0), bw = in_box[s].w - (plan.valid_convs ? 2 * st.L : 0);
      if (st.pool_same && ((bh | bw) & 1)) return false;
      S /= 2;
    }
  }
  const int RH = last.h, RW = last.w;
  // the tail pools + global reduction are computed as one global mean: every tail pool must tile the map exactly
  const int pools = plan.tail_pools;
  if (RH % (1 << pools) != 0 || RW % (1 << pools) != 0) return false;
  if (plan.tail_flatten && ((RH >> pools) != 1 || (RW >> pools) != 1)) return false;  // Flatten needs a 1x1 map
  return true;
}

template <typename T, int S>
struct StageCfg {
  // fp32: 8 w per thread (128-thread groups at S = 32); fp64: 4 (register pressure)
  static constexpr int WPT = sizeof(T) == 4 ? 8 : 4;
};

template <typename T, int S, int WPT, int L, int IN, int EPI, bool NTK, int CIN, int SH>
size_t stage_smem_bytes() {
  using G = StageGeom<S, WPT, SH>;
  const int xs1 = IN == IN_FROM_X ? S * S * CIN : 0;
  const int xs2 = IN == IN_FROM_X ? S * S * 4 : 0;
  const int qm = L * S * S * 2;
  const int stg = EPI == EPI_POOL ? 2 * (NTK ? 2 : 1) * S * (S + 1) : 0;
  if (SH > 1) return (size_t)((xs1 + qm) + G::GROUPS * (xs2 + qm + stg)) * sizeof(T);
  return (size_t)G::GROUPS * (xs1 + xs2 + 2 * qm + stg) * sizeof(T);
}

template <typename T, int S, int WPT, int L, int IN, int EPI, bool NTK, int CIN, int ERF, int SH = 1, bool RC = false,
          bool EMB = false>
int launch_stage_impl(cudaStream_t stream, int64_t* launches, const StageArgs<T>& a) {
  using G = StageGeom<S, WPT, SH>;
  auto kern = k_stage<T, S, WPT, L, IN, EPI, NTK, CIN, SH, RC, ERF, EMB>;
  const size_t smem = stage_smem_bytes<T, S, WPT, L, IN, EPI, NTK, CIN, SH>();
  NTK_TRY(ensure_dynamic_smem((const void*)kern, smem));
  long long blocks = (a.P + G::GROUPS - 1) / G::GROUPS;
  if (SH > 1) blocks = (a.P / a.n2) * ((a.n2 + SH - 1) / SH);
  (*launches)++;
  kern<<<(unsigned)blocks, G::NT, smem, stream>>>(a);
  NTK_CUDA(cudaGetLastError());
  return NTK_OK;
}

template <typename T, int S, int L, int IN, bool NTK, int CIN, int ERF, bool EMB = false>
int launch_stage_epi(cudaStream_t stream, int64_t* launches, int epi, const StageArgs<T>& a) {
  constexpr int WPT = StageCfg<T, S>::WPT;
  // (An SH = 3 variant -- three column samples sharing one row sample per CTA, 12 instead of 8 resident
  // warps per SM -- was measured on B200 in round 1: no gain, the issue rate stayed at ~73 %.  The
  // kernel keeps the SH template parameter; the variant is no longer instantiated.)
  if (!NTK && a.col_count != S) {  // self-pair pipeline (nngp only): partial column range
    switch (epi) {
      case EPI_STORE:
        return launch_stage_impl<T, S, WPT, L, IN, EPI_STORE, false, CIN, ERF, 1, true, EMB>(stream, launches, a);
      case EPI_POOL:
        return launch_stage_impl<T, S, WPT, L, IN, EPI_POOL, false, CIN, ERF, 1, true, EMB>(stream, launches, a);
      default:
        return fail(NTK_EINVAL, "partial column range with a GAP epilogue");
    }
  }
  switch (epi) {
    case EPI_STORE:
      return launch_stage_impl<T, S, WPT, L, IN, EPI_STORE, NTK, CIN, ERF, 1, false, EMB>(stream, launches, a);
    case EPI_POOL:
      return launch_stage_impl<T, S, WPT, L, IN, EPI_POOL, NTK, CIN, ERF, 1, false, EMB>(stream, launches, a);
    default:
      return launch_stage_impl<T, S, WPT, L, IN, EPI_GAP, NTK, CIN, ERF, 1, false, EMB>(stream, launches, a);
  }
}

template <typename T, int S, int IN, bool NTK, int CIN, int ERF, bool EMB = false>
int launch_stage_L(cudaStream_t stream, int64_t* launches, int L, int epi, const StageArgs<T>& a) {
  switch (L) {
    case 1:
      return launch_stage_epi<T, S, 1, IN, NTK, CIN, ERF, EMB>(stream, launches, epi, a);
    case 2:
      return launch_stage_epi<T, S, 2, IN, NTK, CIN, ERF, EMB>(stream, launches, epi, a);
    default:
      return launch_stage_epi<T, S, 3, IN, NTK, CIN, ERF, EMB>(stream, launches, epi, a);
  }
}

// EMB family: images of any size RH x RW <= S x S (FROM_X: C in {1, 3}).  ERF = 0: pure ABRelu (fused_*_emb.cu); ERF = 2: the
// general activation family (ABRelu / Erf / Gelu / Sin / Rbf, fused_*_emb_gen.cu) for MNIST-sized inputs and VALID stacks.
template <typename T, bool NTK, int ERF = 0>
int launch_stage_emb(cudaStream_t stream, int64_t* launches, int S, int L, int from_x, int C, int epi,
                     const StageArgs<T>& a) {
  if (from_x) {
    if (C == 1) {
      if (S == 32) return launch_stage_L<T, 32, IN_FROM_X, NTK, 1, ERF, true>(stream, launches, L, epi, a);
      if (S == 16) return launch_stage_L<T, 16, IN_FROM_X, NTK, 1, ERF, true>(stream, launches, L, epi, a);
      return launch_stage_L<T, 8, IN_FROM_X, NTK, 1, ERF, true>(stream, launches, L, epi, a);
    }
    if (C != 3) return fail(NTK_EUNSUPPORTED, "fused FROM_X stages are instantiated for C == 1 and C == 3");
    if (S == 32) return launch_stage_L<T, 32, IN_FROM_X, NTK, 3, ERF, true>(stream, launches, L, epi, a);
    if (S == 16) return launch_stage_L<T, 16, IN_FROM_X, NTK, 3, ERF, true>(stream, launches, L, epi, a);
    return launch_stage_L<T, 8, IN_FROM_X, NTK, 3, ERF, true>(stream, launches, L, epi, a);
  }
  if (S == 32) return launch_stage_L<T, 32, IN_LOAD, NTK, 1, ERF, true>(stream, launches, L, epi, a);
  if (S == 16) return launch_stage_L<T, 16, IN_LOAD, NTK, 1, ERF, true>(stream, launches, L, epi, a);
  return launch_stage_L<T, 8, IN_LOAD, NTK, 1, ERF, true>(stream, launches, L, epi, a);
}

// Packed-FP32 (FFMA2) instance for fp32 at 32x32: stage_packed.cuh / stage_packed.cu.
int launch_stage_packed(cudaStream_t stream, int64_t* launches, int S, int L, int from_x, int epi, bool ntk,
                        const StageArgs<float>& a);
int stage_packed_configure();
// the same kernels with the Erf closed form compiled in (stage_packed_erf.cu)
int launch_stage_packed_erf(cudaStream_t stream, int64_t* launches, int S, int L, int from_x, int epi, bool ntk,
                            const StageArgs<float>& a);
int stage_packed_erf_configure();

inline bool packed_enabled() {
  static const bool on = getenv("NTK_B200_NO_PACKED") == nullptr;
  return on;
}
inline int launch_stage_packed_any(cudaStream_t, int64_t*, int, int, int, int, bool, bool, const StageArgs<double>&) {
  return fail(NTK_EINVAL, "packed stage kernel is fp32 only");
}
inline int launch_stage_packed_any(cudaStream_t stream, int64_t* launches, int S, int L, int from_x, int epi,
                                   bool ntk, bool erf, const StageArgs<float>& a) {
  if (erf) return launch_stage_packed_erf(stream, launches, S, L, from_x, epi, ntk, a);
  return launch_stage_packed(stream, launches, S, L, from_x, epi, ntk, a);
}

// The unpacked stage kernels: ERF = false (pure ABRelu) and ERF = true (Erf-capable) families are
// instantiated in separate translation units (fused_*.cu / fused_*_erf.cu).
template <typename T, bool NTK, int ERF>
int launch_stage_k(cudaStream_t stream, int64_t* launches, int S, int L, int from_x, int C, int epi,
                   const StageArgs<T>& a) {
  if (from_x) {
    if (C != 3) return fail(NTK_EUNSUPPORTED, "fused FROM_X stages are instantiated for C == 3");
    if (S == 32) return launch_stage_L<T, 32, IN_FROM_X, NTK, 3, ERF>(stream, launches, L, epi, a);
    if (S == 16) return launch_stage_L<T, 16, IN_FROM_X, NTK, 3, ERF>(stream, launches, L, epi, a);
    return launch_stage_L<T, 8, IN_FROM_X, NTK, 3, ERF>(stream, launches, L, epi, a);
  }
  if (S == 32) return launch_stage_L<T, 32, IN_LOAD, NTK, 1, ERF>(stream, launches, L, epi, a);
  if (S == 16) return launch_stage_L<T, 16, IN_LOAD, NTK, 1, ERF>(stream, launches, L, epi, a);
  return launch_stage_L<T, 8, IN_LOAD, NTK, 1, ERF>(stream, launches, L, epi, a);
}

// Which stages run on the packed kernels: fp32 at 32x32 / 16x16 whose activations are all ABRelu or all Erf.
// (A stage that mixes them stays on the scalar kernel, whose Erf arithmetic is the one k_qmaps uses for the
// diagonal: an ABRelu layer behind an Erf layer needs that bit-exact agreement on duplicate pairs.)
template <typename T>
bool stage_is_packed(int S, int C, int n_erf, int L) {
  return sizeof(T) == 4 && (S == 32 || S == 16) && C == 3 && packed_enabled() && (n_erf == 0 || n_erf == L);
}

template <typename T, bool NTK>
int launch_stage(cudaStream_t stream, int64_t* launches, int S, int L, int from_x, int C, int epi,
                 const StageArgs<T>& a, bool emb = false) {
  int n_erf = 0, n_gen = 0;
  for (int l = 0; l < L; ++l) {
    n_erf += a.lp[l].kind == ACT_ERF;
    n_gen += a.lp[l].kind >= ACT_GELU;
  }
  const bool any_erf = n_erf > 0;
  if (emb) {
    if (any_erf || n_gen) return launch_stage_emb<T, NTK, 2>(stream, launches, S, L, from_x, C, epi, a);
    return launch_stage_emb<T, NTK>(stream, launches, S, L, from_x, C, epi, a);
  }
  if (n_gen) return launch_stage_k<T, NTK, 2>(stream, launches, S, L, from_x, C, epi, a);
  if (stage_is_packed<T>(S, from_x ? C : 3, n_erf, L))
    return launch_stage_packed_any(stream, launches, S, L, from_x, epi, NTK, any_erf, a);
  if (any_erf) return launch_stage_k<T, NTK, 1>(stream, launches, S, L, from_x, C, epi, a);
  return launch_stage_k<T, NTK, 0>(stream, launches, S, L, from_x, C, epi, a);
}

template <typename T>
int launch_qmaps(cudaStream_t stream, int64_t* launches, int S, const T* src, int src_mode, int n,
                 int C, T in_scale, int L, const FLayer<T>* lp, T* qm, int RH, int RW) {
  (*launches)++;
  FLayer<T> z{(T)0, (T)0, (T)0, (T)0};
  FLayer<T> l0 = lp[0], l1 = L > 1 ? lp[1] : z, l2 = L > 2 ? lp[2] : z;
  if (S == 32)
    k_qmaps<T, 32><<<n, 256, 0, stream>>>(src, src_mode, C, in_scale, L, l0, l1, l2, qm, RH, RW);
  else if (S == 16)
    k_qmaps<T, 16><<<n, 256, 0, stream>>>(src, src_mode, C, in_scale, L, l0, l1, l2, qm, RH, RW);
  else
    k_qmaps<T, 8><<<n, 64, 0, stream>>>(src, src_mode, C, in_scale, L, l0, l1, l2, qm, RH, RW);
  NTK_CUDA(cudaGetLastError());
  return NTK_OK;
}

// Uploads the vertical link masks to the current device (once per context).  Templated on the
// dtype and kernel family so that each translation unit (fused_f32.cu, fused_f32_erf.cu, ...)
// uploads its own copy of the `static __constant__` tables.
template <typename T, int ERF>
int fused_configure_device() {
  for (int S : {32, 16, 8}) {
    std::vector<float2> m((size_t)S * S);
    for (int ch = 0; ch < S; ++ch)
      for (int h = 0; h < S; ++h) {
        const int h2 = (h + ch) % S;
        m[(size_t)ch * S + h].x = (h > 0 && h2 != 0) ? 1.f : 0.f;
        m[(size_t)ch * S + h].y = (h < S - 1 && h2 != S - 1) ? 1.f : 0.f;
      }
    const size_t bytes = m.size() * sizeof(float2);
    if (S == 32) NTK_CUDA(cudaMemcpyToSymbol(c_vmask32, m.data(), bytes));
    if (S == 16) NTK_CUDA(cudaMemcpyToSymbol(c_vmask16, m.data(), bytes));
    if (S == 8) NTK_CUDA(cudaMemcpyToSymbol(c_vmask8, m.data(), bytes));
  }
  return NTK_OK;
}

// Per-stage layer constants: fold the NEXT conv's alpha = W^2/9 into this layer's ABRelu.
template <typename T>
void stage_constants(const FusedPlan& plan, size_t s, FLayer<T>* lp, double* next_alpha_out) {
  const FusedStage& st = plan.stages[s];
  const double two_pi = 2.0 * 3.14159265358979323846;
  for (int l = 0; l < st.L; ++l) {
    double alpha_next = 1.0;
    if (l + 1 < st.L)
      alpha_next = st.w2[l + 1] / 9.0;
    else if (st.epi != EPI_GAP && s + 1 < plan.stages.size() && plan.stages[s + 1].L > 0)
      alpha_next = plan.stages[s + 1].w2[0] / 9.0;
    const double d = st.a[l] - st.b[l];
    double coef = d * d / two_pi, half_ab = (st.a[l] * st.a[l] + st.b[l] * st.b[l]) / 2.0;
    // the scale feeding a pooled boundary is applied by the pool epilogue instead
    const bool via_epi = (l + 1 == st.L) && st.epi == EPI_POOL;
    if (!via_epi) {
      coef *= alpha_next;
      half_ab *= alpha_next;
    }
    lp[l].coef = (T)coef;
    lp[l].half_ab = (T)half_ab;
    lp[l].hab2 = (T)(half_ab - coef * 1.57079632679489661923);
    lp[l].bias = (T)st.b2[l];
    {
      const double an = via_epi ? 1.0 : alpha_next, pi = 3.14159265358979323846;
      const double ea = st.a[l], eb = st.b[l], ec = st.c[l];
      lp[l].kind = st.kind[l];
      lp[l].e_in = (T)(2.0 * eb * eb);
      lp[l].eA = (T)(an * ea * ea * 2.0 / pi);
      lp[l].eT = (T)(an * ea * ea * eb * eb * 4.0 / pi);
      lp[l].eC = (T)(an * ec * ec);
      if (st.kind[l] == ACT_GELU) {
        lp[l].eA = (T)an;
      } else if (st.kind[l] == ACT_SIN) {  // a sin(b x + c)
        lp[l].e_in = (T)(eb * eb);
        lp[l].eA = (T)(an * ea * ea / 2.0);
        lp[l].eT = (T)cos(2.0 * ec);
      } else if (st.kind[l] == ACT_RBF) {  // gamma travels in `a`
        lp[l].e_in = (T)ea;
        lp[l].eA = (T)an;
      }
    }
    if (l + 1 == st.L) *next_alpha_out = alpha_next;
  }
}

// Whole Gram block through the fused stages.  x1/x2 are device pointers to all samples;
// the pair grid is tiled so that the stage boundaries fit in the arena.
template <typename T>
int fused_gram(const FusedPlan& plan, Arena& arena, cudaStream_t stream, int64_t* launches,
               StageProfile* prof, const T* x1, int n1, const T* x2, int n2, bool symmetric, int H0,
               int W0, int C, bool want_ntk, T* out_nngp, T* out_ntk, long long ld,
               bool full_square = false, bool upper = false) {
  const int S0 = fused_shear_size(H0, W0);
  const bool emb = fused_route(plan, H0, W0, C).emb;
  const bool prepass = fused_route(plan, H0, W0, C).prepass;  // stage 0 LOADs the sheared input covariance of k_input_shear
  const size_t xrow = (size_t)H0 * W0 * C;  // elements per input sample
  // `upper` (NTK_FLAG_UPPER_ONLY): x1 holds the same samples as x2[0:n1]; only entries (i, j >= i) are wanted.
  upper = upper && !symmetric && n2 >= n1;
  const bool triangular = (symmetric && !full_square && n1 == n2 && n1 > 1) || upper;
  const bool share_q = symmetric || upper;  // the q-maps of x1 are (a prefix of) those of x2
  const size_t n_st = plan.stages.size() - 1;  // real stages (last entry is the tail marker)
  // ---- 1. q-maps for every stage and both sample sets (self-pair pipeline) --------------
  // qm[s][set]: [n][L][S][S][2]
  std::vector<T*> qm1(n_st), qm2(n_st);
  std::vector<int> Ss(n_st), RHs(n_st), RWs(n_st);
  // RHs / RWs: extent of the region the link masks keep connected at stage s (SAME: the image; VALID: origin + size
  // of the box entering the stage -- the masks are don't-care there, see StageArgs::box_o)
  int RH_last = H0, RW_last = W0;
  FusedBox last_box{0, H0, W0};
  {
    std::vector<FusedBox> in_box;
    if (!fused_walk_boxes(plan, H0, W0, &in_box, &last_box)) return fail(NTK_EINVAL, "fused plan does not fit the input size");
    int S = S0;
    for (size_t s = 0; s < n_st; ++s) {
      Ss[s] = S;
      RHs[s] = in_box[s].o + in_box[s].h;
      RWs[s] = in_box[s].o + in_box[s].w;
      if (plan.stages[s].epi == EPI_POOL) S /= 2;
    }
    RH_last = last_box.h;
    RW_last = last_box.w;
  }
  for (size_t s = 0; s < n_st; ++s) {
    const size_t per = (size_t)plan.stages[s].L * Ss[s] * Ss[s] * 2 * sizeof(T);
    qm2[s] = (T*)arena.alloc(per * n2);
    qm1[s] = share_q ? qm2[s] : (T*)arena.alloc(per * n1);
    if (!qm1[s] || !qm2[s]) return fail(NTK_ENOMEM, "workspace too small for the q-maps");
  }
  const double alpha0 = plan.stages[0].w2[0] / 9.0;
  const T in_scale = (T)(alpha0 / (double)C);
  // Self pairs only feed the q-maps, i.e. the diagonal column (ch = cw = 0) of each boundary.
  // Conv layers never mix ch columns and AvgPool couples Ch with {2Ch-1, 2Ch, 2Ch+1}, so the
  // self-pair run of stage s has to march only the columns [-k_in[s], k_in[s]] (mod S).
  std::vector<int> k_in(n_st, 0);
  for (int s = (int)n_st - 2, k_out = 0; s >= 0; --s) {
    k_in[s] = plan.stages[s].epi == EPI_POOL ? 2 * k_out + 1 : k_out;
    k_out = k_in[s];
  }
  for (int set = share_q ? 1 : 0; set < 2; ++set) {
    const T* x = set == 0 ? x1 : x2;
    const int n = set == 0 ? n1 : n2;
    // self tensors are processed in chunks to bound the workspace
    const int chunk = std::min(n, 512);
    T* bufA = nullptr;
    T* bufB = nullptr;
    size_t cap = 0;
    for (size_t s = 0; s < n_st; ++s) cap = std::max(cap, (size_t)Ss[s] * Ss[s] * Ss[s] * Ss[s]);
    // only boundaries after stage 0 are materialised; the largest is at (S0/2 or S0)
    size_t need = 0;
    {
      for (size_t s = 0; s + 1 < n_st; ++s) {
        const int So = plan.stages[s].epi == EPI_POOL ? Ss[s] / 2 : Ss[s];
        need = std::max(need, (size_t)So * So * So * So);
      }
    }
    T* bufX = nullptr;  // prepass: sheared self covariance entering stage 0
    if (need > 0) {
      bufA = (T*)arena.alloc(need * chunk * sizeof(T));
      bufB = (T*)arena.alloc(need * chunk * sizeof(T));
      if (prepass) bufX = (T*)arena.alloc((size_t)S0 * S0 * S0 * S0 * chunk * sizeof(T));
      if (!bufA || !bufB || (prepass && !bufX)) return fail(NTK_ENOMEM, "workspace too small for the self-pair pipeline");
    }
    for (int c0 = 0; c0 < n; c0 += chunk) {
      const int m = std::min(chunk, n - c0);
      T* cur = nullptr;  // sheared self tensors entering stage s
      for (size_t s = 0; s < n_st; ++s) {
        FLayer<T> lp[kMaxFusedLayers];
        double next_alpha = 1.0;
        stage_constants<T>(plan, s, lp, &next_alpha);
        T* qm = (set == 0 ? qm1[s] : qm2[s]) + (size_t)c0 * plan.stages[s].L * Ss[s] * Ss[s] * 2;
        const int S = Ss[s];
        if (s == 0)
          NTK_TRY(launch_qmaps<T>(stream, launches, S, x + (size_t)c0 * xrow, 0, m, C, in_scale,
                                  plan.stages[s].L, lp, qm, RHs[s], RWs[s]));
        else
          NTK_TRY(launch_qmaps<T>(stream, launches, S, cur, 1, m, C, (T)1, plan.stages[s].L, lp, qm, RHs[s],
                                  RWs[s]));
        if (s + 1 == n_st) break;  // the last stage's self tensors are never needed
        if (s == 0 && prepass) {
          NTK_TRY(launch_input_shear<T>(stream, launches, S, x + (size_t)c0 * xrow, x + (size_t)c0 * xrow, bufX, m, 1, 1, 0,
                                        RHs[0], RWs[0], C, in_scale));
          cur = bufX;
        }
        // run the stage on the self pairs (nngp only) to get the next boundary
        StageArgs<T> a{};
        a.RH = RHs[s];
        a.RW = RWs[s];
        a.box_o = 0;  // self pairs never end in a GAP epilogue
        a.box_h = RHs[s];
        a.box_w = RWs[s];
        a.x1 = a.x2 = x + (size_t)c0 * xrow;
        a.inK = cur;
        a.inT = nullptr;
        T* nxt = (cur == bufA) ? bufB : bufA;
        const int epi = plan.stages[s].epi;
        const int So = epi == EPI_POOL ? S / 2 : S;
        a.outK = nxt;
        a.outT = nullptr;
        a.qm1 = a.qm2 = qm;
        a.P = m;
        a.n2 = 1;
        a.self = 1;
        a.col_count = std::min(S, 2 * k_in[s] + 1);
        a.col_start = a.col_count == S ? 0 : k_in[s];
        a.in_scale = in_scale;
        a.epi_scale = (T)(next_alpha * plan.stages[s].pool_mul / 16.0);
        for (int l = 0; l < plan.stages[s].L; ++l) a.lp[l] = lp[l];
        if (epi == EPI_POOL)
          NTK_CUDA(cudaMemsetAsync(nxt, 0, (size_t)m * So * So * So * So * sizeof(T), stream));
        NTK_TRY((launch_stage<T, false>(stream, launches, S, plan.stages[s].L, s == 0 && !prepass, prepass ? 3 : C, epi, a,
                                        emb)));
        cur = nxt;
      }
    }
    if (bufX) arena.release(bufX);
    if (bufA) arena.release(bufA);
    if (bufB) arena.release(bufB);
  }

  // ---- 2. cross pairs, tiled --------------------------------------------------------------
  // bytes per pair of the materialised boundaries (two live at a time: in + out)
  size_t max_b = 0, sum2 = 0;
  {
    std::vector<size_t> bsz;
    if (prepass) bsz.push_back((size_t)S0 * S0 * S0 * S0 * sizeof(T));  // the sheared input covariance (nngp only)
    for (size_t s = 0; s + 1 < n_st; ++s) {
      const int So = plan.stages[s].epi == EPI_POOL ? Ss[s] / 2 : Ss[s];
      bsz.push_back((size_t)So * So * So * So * sizeof(T) * (want_ntk ? 2 : 1));
    }
    for (size_t i = 0; i < bsz.size(); ++i) {
      max_b = std::max(max_b, bsz[i]);
      sum2 = std::max(sum2, bsz[i] + (i + 1 < bsz.size() ? bsz[i + 1] : 0));
    }
  }
  // Tile size from the free workspace: two boundary buffers of max_b bytes per pair plus the [pairs] results.
  // (Large tiles matter: every stage launch ends in a partial wave of ~1 ms CTAs; 40k-pair tiles lose ~3 % to
  // those tails, 150k-pair tiles < 1 %.)
  long long tile_pairs = (long long)n1 * n2;
  if (sum2 > 0) {
    const size_t per_pair = 2 * max_b + 2 * sizeof(T);
    const size_t avail = arena.largest_free();
    const long long fit = avail > ((size_t)8 << 20) ? (long long)((avail - ((size_t)8 << 20)) / per_pair) : 1;
    tile_pairs = std::max<long long>(1, std::min<long long>(tile_pairs, fit));
  }
  T* bnd[2] = {nullptr, nullptr};
  size_t bnd_bytes = 0;
  if (sum2 > 0) {
    for (;;) {
      bnd_bytes = (size_t)tile_pairs * max_b;
      bnd[0] = (T*)arena.alloc(bnd_bytes);
      bnd[1] = bnd[0] ? (T*)arena.alloc(bnd_bytes) : nullptr;
      if (bnd[0] && bnd[1]) break;
      if (bnd[0]) arena.release(bnd[0]);
      bnd[0] = bnd[1] = nullptr;
      if (tile_pairs <= 1) return fail(NTK_ENOMEM, "workspace too small for one pair");
      tile_pairs = (tile_pairs + 1) / 2;
    }
  }
  // tile = t1 rows x n2 columns when possible, else 1 row x t2 columns
  int t1, t2;
  if (tile_pairs >= n2) {
    t2 = n2;
    t1 = (int)std::min<long long>(n1, tile_pairs / n2);
  } else {
    t1 = 1;
    t2 = (int)tile_pairs;
  }
  T* resK = (T*)arena.alloc((size_t)t1 * t2 * sizeof(T));
  T* resT = want_ntk ? (T*)arena.alloc((size_t)t1 * t2 * sizeof(T)) : nullptr;
  if (!resK || (want_ntk && !resT)) return fail(NTK_ENOMEM, "workspace too small");

  // tail: pools + GlobalAvgPool / Flatten == global mean over the last stage's map; SumPool / GlobalSumPool scale it
  double tail_scale = plan.tail_mul / ((double)RH_last * RH_last * RW_last * RW_last);
  if (plan.tail_gsum) {
    const double fh = (double)(RH_last >> plan.tail_pools), fw = (double)(RW_last >> plan.tail_pools);
    tail_scale *= fh * fh * fw * fw;
  }
  for (int r0 = 0; r0 < n1; r0 += t1) {
    const int a1 = std::min(t1, n1 - r0);
    // K(x, x) is symmetric: only pairs (i, j >= i) are computed and the strictly lower triangle
    // is mirrored afterwards (the reference computes the full square, `_src/batching.py:370`).
    // Full-width tiles enumerate the triangle pair by pair; single-row tiles start at c0 = r0.
    const bool tri_tile = triangular && t2 == n2;
    for (int c0 = triangular ? r0 : 0; c0 < n2; c0 += t2) {
      const int a2 = tri_tile ? n2 - r0 : std::min(t2, n2 - c0);
      const long long P = tri_tile ? tri_prefix(a1, a2) : (long long)a1 * a2;
      T* cur = nullptr;
      if (prepass) {
        NTK_TRY(launch_input_shear<T>(stream, launches, S0, x1 + (size_t)r0 * xrow, x2 + (size_t)c0 * xrow, bnd[0], P, a2, 0,
                                      tri_tile ? 1 : 0, RHs[0], RWs[0], C, in_scale));
        cur = bnd[0];
      }
      for (size_t s = 0; s < n_st; ++s) {
        const int S = Ss[s];
        FLayer<T> lp[kMaxFusedLayers];
        double next_alpha = 1.0;
        stage_constants<T>(plan, s, lp, &next_alpha);
        StageArgs<T> a{};
        a.RH = RHs[s];
        a.RW = RWs[s];
        a.box_o = last_box.o;  // read by the GAP epilogue (last stage) only
        a.box_h = last_box.h;
        a.box_w = last_box.w;
        a.x1 = x1 + (size_t)r0 * xrow;
        a.x2 = x2 + (size_t)c0 * xrow;
        const int epi = plan.stages[s].epi;
        const int So = epi == EPI_POOL ? S / 2 : S;
        const size_t in_per = (size_t)S * S * S * S, out_per = (size_t)So * So * So * So;
        a.inK = cur;
        a.inT = (cur && want_ntk && !(s == 0 && prepass)) ? cur + (size_t)P * in_per : nullptr;  // prepass: ntk = 0
        T* nxt = (cur == bnd[0]) ? bnd[1] : bnd[0];
        if (epi == EPI_GAP) {
          a.outK = resK;
          a.outT = resT;
          a.epi_scale = (T)tail_scale;
        } else {
          a.outK = nxt;
          a.outT = want_ntk ? nxt + (size_t)P * out_per : nullptr;
          a.epi_scale = (T)(next_alpha * plan.stages[s].pool_mul / 16.0);
          if (epi == EPI_POOL) {
            // the packed kernels zero their own (pre-accumulation) output while other CTAs compute
            // -- DRAM is idle in this kernel -- instead of a 4.8 GB memset in front of every launch
            int n_erf = 0, n_gen = 0;
            for (int l = 0; l < plan.stages[s].L; ++l) {
              n_erf += plan.stages[s].kind[l] == ACT_ERF;
              n_gen += plan.stages[s].kind[l] >= ACT_GELU;
            }
            a.zero_out = (!emb && !n_gen && stage_is_packed<T>(S, (s == 0 && !prepass) ? C : 3, n_erf, plan.stages[s].L)) ? 1 : 0;
            if (!a.zero_out)
              NTK_CUDA(cudaMemsetAsync(nxt, 0, (size_t)P * out_per * sizeof(T) * (want_ntk ? 2 : 1), stream));
          }
        }
        a.qm1 = qm1[s] + (size_t)r0 * plan.stages[s].L * S * S * 2;
        a.qm2 = qm2[s] + (size_t)c0 * plan.stages[s].L * S * S * 2;
        a.P = P;
        a.n2 = a2;
        a.self = 0;
        a.tri = tri_tile ? 1 : 0;
        a.col_start = 0;
        a.col_count = S;
        a.in_scale = in_scale;
        for (int l = 0; l < plan.stages[s].L; ++l) a.lp[l] = lp[l];
        cudaEvent_t ev0 = nullptr, ev1 = nullptr;
        const bool timed = prof && prof->enabled && s < (size_t)StageProfile::kMaxStages;
        if (timed) {
          NTK_CUDA(cudaEventCreate(&ev0));
          NTK_CUDA(cudaEventCreate(&ev1));
          NTK_CUDA(cudaEventRecord(ev0, stream));
        }
        const int from_x = (s == 0 && !prepass) ? 1 : 0, cin = prepass ? 3 : C;
        if (want_ntk)
          NTK_TRY((launch_stage<T, true>(stream, launches, S, plan.stages[s].L, from_x, cin, epi, a, emb)));
        else
          NTK_TRY((launch_stage<T, false>(stream, launches, S, plan.stages[s].L, from_x, cin, epi, a, emb)));
        if (timed) {
          NTK_CUDA(cudaEventRecord(ev1, stream));
          prof->pending[s].push_back({ev0, ev1});
          prof->pending_pairs[s].push_back(P);
        }
        cur = nxt;
      }
      // Dense tail on the [a1, a2] scalars (linear.py:899-926), then scatter into the result
      for (const ntk_op_t& d : plan.dense_tail) {
        (*launches)++;
        k_dense<T><<<grid_for(P), kThreads, 0, stream>>>(resK, resT, P, (T)d.f[0],
                                                         (T)(d.i[0] ? d.f[1] : 0.0), 0);
        NTK_CUDA(cudaGetLastError());
      }
      (*launches)++;
      if (tri_tile)
        k_scatter_tri<T><<<grid_for(P), kThreads, 0, stream>>>(resK, out_nngp, P, a2, ld, r0);
      else
        k_scatter<T><<<grid_for(P), kThreads, 0, stream>>>(resK, out_nngp, a1, a2, 1LL, ld, r0, c0);
      NTK_CUDA(cudaGetLastError());
      if (want_ntk) {
        (*launches)++;
        if (tri_tile)
          k_scatter_tri<T><<<grid_for(P), kThreads, 0, stream>>>(resT, out_ntk, P, a2, ld, r0);
        else
          k_scatter<T><<<grid_for(P), kThreads, 0, stream>>>(resT, out_ntk, a1, a2, 1LL, ld, r0, c0);
        NTK_CUDA(cudaGetLastError());
      }
      if (tri_tile) break;  // one triangular tile covers all columns of this row block
    }
  }
  if (triangular && !upper) {
    (*launches)++;
    k_mirror<T><<<grid_for((long long)n1 * n1), kThreads, 0, stream>>>(out_nngp, n1, ld);
    NTK_CUDA(cudaGetLastError());
    if (want_ntk) {
      (*launches)++;
      k_mirror<T><<<grid_for((long long)n1 * n1), kThreads, 0, stream>>>(out_ntk, n1, ld);
      NTK_CUDA(cudaGetLastError());
    }
  }
  return NTK_OK;
}

// FCN input Gram x1 x2^T / d  (requirements.py:585-638 for 2-D inputs): tensor cores
// (gemm_kernels.cuh) unless the feature dimension is tiny or NTK_B200_NO_TC is set.
template <typename T>
int fcn_input_gram(bool dry, cudaStream_t stream, int64_t* launches, const T* x1, int t1, const T* x2,
                   int t2, int C, T* out) {
  (*launches)++;
  if (dry) return NTK_OK;
  static const bool no_tc = getenv("NTK_B200_NO_TC") != nullptr;
  if (C >= 16 && !no_tc) return launch_gram_tc(stream, x1, t1, x2, t2, C, out);
  const long long P = (long long)t1 * t2;
  k_rowdot<T><<<grid_for(P * 32), kThreads, 0, stream>>>(x1, x2, out, P, PairMap{t2, 0}, C,
                                                         (T)(1.0 / (double)C));
  NTK_CUDA(cudaGetLastError());
  return NTK_OK;
}

}  // namespace ntk
