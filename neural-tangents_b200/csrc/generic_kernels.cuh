// One-kernel-per-layer path: straightforward, coalesced, grid-stride kernels on
// the canonical zipped layout [pair, h, h', w, w'].  This is the general path
// (any filter/stride/padding, Erf, FanInSum, Kernel-in/Kernel-out); the fused
// diagonal-marching kernels in fused_kernels.cuh take over for the stride-1
// 3x3 SAME Conv+ABRelu(+AvgPool 2x2) stacks that dominate the Myrtle configs.
//
// Math follows SURVEY.md Appendix A; each kernel cites the reference rule.
#pragma once

#include "common.cuh"

namespace ntk {

template <typename T>
struct Consts;
template <>
struct Consts<float> {
  static __host__ __device__ constexpr float pi() { return 3.14159265358979323846f; }
};
template <>
struct Consts<double> {
  static __host__ __device__ constexpr double pi() { return 3.14159265358979323846; }
};

__device__ __forceinline__ float fma_t(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double fma_t(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float sqrt_t(float a) { return sqrtf(a); }
__device__ __forceinline__ double sqrt_t(double a) { return sqrt(a); }
__device__ __forceinline__ float atan2_t(float y, float x) { return atan2f(y, x); }
__device__ __forceinline__ double atan2_t(double y, double x) { return atan2(y, x); }
__device__ __forceinline__ float max_t(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double max_t(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float abs_t(float a) { return fabsf(a); }
__device__ __forceinline__ double abs_t(double a) { return fabs(a); }

// Pair index -> (row sample, column sample).  `self`: pair p is (p, p) (cov1/cov2).
struct PairMap {
  int n2;
  int self;
  __device__ __forceinline__ void ij(long long p, int& i, int& j) const {
    if (self) {
      i = j = (int)p;
    } else {
      i = (int)(p / n2);
      j = (int)(p % n2);
    }
  }
};

constexpr int kThreads = 256;

inline int grid_for(long long n, int threads = kThreads) {
  long long b = (n + threads - 1) / threads;
  long long cap = (long long)kNumSMs * 32;  // grid-stride beyond this
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// Row decomposition shared by the per-op kernels: a tensor [P, A, A, B, B] is walked as rows (p, a, a') by groups of
// G = min(32, pow2ceil(B * B)) lanes, the lanes of a group stride over the (b, b') plane.  The 64-bit divisions are paid
// once per row instead of once per element (round 1: ~200 integer instructions per element, the per-op path ran at 12 %
// of its own HBM traffic bound); consecutive lanes still touch consecutive b' (coalesced).
struct RowWalk {
  long long row, nrows, stride;
  int e0, G;
  // `items` = work items per row handled by the lanes of a group (B * B elements, or B * B / 4 quads)
  __device__ __forceinline__ RowWalk(long long P, int A, int B, int items = 0) {
    const int bb = items > 0 ? items : B * B;
    G = 32;
    while (G > 1 && (G >> 1) >= bb) G >>= 1;
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    row = t / G;
    e0 = (int)(t % G);
    stride = (long long)gridDim.x * blockDim.x / G;
    nrows = P * A * A;
  }
};

// ---- input layer: requirements.py:542-553 (cross) / 529-539 (self) -----------
// out[p,h,h',w,w'] = (1/C) sum_c x1[i,h,w,c] * x2[j,h',w',c]
template <typename T>
__global__ void k_input_cov(const T* __restrict__ x1, const T* __restrict__ x2, T* __restrict__ out,
                            long long P, PairMap pm, int H, int W, int C, T inv_c) {
  RowWalk rw(P, H, W);
  const int ww = W * W;
  for (long long row = rw.row; row < rw.nrows; row += rw.stride) {
    const long long p = row / ((long long)H * H);
    const int r = (int)(row - p * ((long long)H * H));
    const int h = r / H, h2 = r - h * H;
    int i, j;
    pm.ij(p, i, j);
    const T* arow = x1 + ((long long)i * H + h) * W * C;
    const T* brow = x2 + ((long long)j * H + h2) * W * C;
    T* orow = out + row * ww;
    for (int e = rw.e0; e < ww; e += rw.G) {
      const int w = e / W, w2 = e - w * W;
      const T* a = arow + w * C;
      const T* b = brow + w2 * C;
      T acc = 0;
      for (int c = 0; c < C; ++c) acc = fma_t(a[c], b[c], acc);
      orow[e] = acc * inv_c;
    }
  }
}

// FCN input ([N,d] inputs): nngp[i,j] = x1[i].x2[j]/d, one warp per output.
// (The tensor-core split-precision GEMM in gemm_kernels.cuh replaces this for
// large d; this kernel is the exact-accumulation fallback for tiny/odd shapes.)
template <typename T>
__global__ void k_rowdot(const T* __restrict__ x1, const T* __restrict__ x2, T* __restrict__ out,
                         long long P, PairMap pm, int C, T inv_c) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long p = warp; p < P; p += nwarps) {
    int i, j;
    pm.ij(p, i, j);
    const T* a = x1 + (long long)i * C;
    const T* b = x2 + (long long)j * C;
    T acc = 0;
    for (int c = lane; c < C; c += 32) acc = fma_t(a[c], b[c], acc);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[p] = acc * inv_c;
  }
}

// ---- Conv rule: linear.py:3341-3378 (diagonal-offset box filter) -------------
// out[p,a,a',b,b'] = scale/(kh kw) * sum_{i<kh,j<kw} in~[sa+i-lo, sa'+i-lo, sb+j-lo, sb'+j-lo]
//                    + shift (+ addend[p,a,a',b,b'])
struct ConvGeom {
  int Hi, Wi, Ho, Wo, kh, kw, sh, sw, loh, low, circular;
};

template <typename T>
__global__ void k_conv(const T* __restrict__ in, const T* __restrict__ addend, T* __restrict__ out,
                       long long P, ConvGeom g, T scale, T shift) {
  const long long per_i = (long long)g.Hi * g.Hi * g.Wi * g.Wi;
  const int ww = g.Wo * g.Wo;
  RowWalk rw(P, g.Ho, g.Wo);
  // Without wrap-around the valid taps of an output are the index ranges [ilo, ihi) x [jlo, jhi) and their addresses an
  // arithmetic progression: tap (i, j) sits at base + i * SI + j * SJ (both members move together along the diagonal).
  const int SI = (g.Hi + 1) * g.Wi * g.Wi, SJ = g.Wi + 1;
  for (long long row = rw.row; row < rw.nrows; row += rw.stride) {
    const long long p = row / ((long long)g.Ho * g.Ho);
    const int r = (int)(row - p * ((long long)g.Ho * g.Ho));
    const int a = r / g.Ho, a2 = r - a * g.Ho;
    const T* src = in + p * per_i;
    const T* arow = addend ? addend + row * ww : nullptr;
    T* orow = out + row * ww;
    if (!g.circular) {
      const int h0 = g.sh * a - g.loh, h20 = g.sh * a2 - g.loh;
      const int ilo = max(0, max(-h0, -h20)), ihi = min(g.kh, min(g.Hi - h0, g.Hi - h20));
      const T* rbase = src + ((long long)h0 * g.Hi + h20) * (g.Wi * g.Wi);  // may point before `src`: only valid taps are read
      for (int e = rw.e0; e < ww; e += rw.G) {
        const int b = e / g.Wo, b2 = e - b * g.Wo;
        const int w0 = g.sw * b - g.low, w20 = g.sw * b2 - g.low;
        const int jlo = max(0, max(-w0, -w20)), jhi = min(g.kw, min(g.Wi - w0, g.Wi - w20));
        const T* ebase = rbase + (w0 * g.Wi + w20);
        T acc = 0;
        for (int i = ilo; i < ihi; ++i) {
          const T* tp = ebase + i * SI + jlo * SJ;
          for (int j = jlo; j < jhi; ++j, tp += SJ) acc += __ldg(tp);
        }
        T v = fma_t(acc, scale, shift);
        if (arow) v += arow[e];
        orow[e] = v;
      }
      continue;
    }
    for (int e = rw.e0; e < ww; e += rw.G) {
      const int b = e / g.Wo, b2 = e - b * g.Wo;
      T acc = 0;
      for (int i = 0; i < g.kh; ++i) {
        const int h = (((g.sh * a + i - g.loh) % g.Hi) + g.Hi) % g.Hi;
        const int h2 = (((g.sh * a2 + i - g.loh) % g.Hi) + g.Hi) % g.Hi;
        const T* plane = src + ((long long)h * g.Hi + h2) * g.Wi * g.Wi;
        for (int j = 0; j < g.kw; ++j) {
          const int w = (((g.sw * b + j - g.low) % g.Wi) + g.Wi) % g.Wi;
          const int w2 = (((g.sw * b2 + j - g.low) % g.Wi) + g.Wi) % g.Wi;
          acc += plane[w * g.Wi + w2];
        }
      }
      T v = fma_t(acc, scale, shift);
      if (arow) v += arow[e];
      orow[e] = v;
    }
  }
}

// 3 x 3 filters without wrap-around (every SAME / VALID 3 x 3 conv, any stride): the nine taps unrolled and predicated, nngp
// and ntk in ONE pass -- the ntk's addend is the nngp value just computed (linear.py:1396-1398), so it is never re-read.
// Same tap order and the same roundings as k_conv.
template <typename T>
__global__ void k_conv3(const T* __restrict__ inK, const T* __restrict__ inT, T* __restrict__ outK,
                        T* __restrict__ outT, long long P, ConvGeom g, T scale, T shift) {
  const int WW = g.Wi * g.Wi;
  const long long per_i = (long long)g.Hi * g.Hi * WW;
  const int ww = g.Wo * g.Wo;
  const int SI = (g.Hi + 1) * WW, SJ = g.Wi + 1;
  const bool has_t = inT != nullptr;
  RowWalk rw(P, g.Ho, g.Wo);
  for (long long row = rw.row; row < rw.nrows; row += rw.stride) {
    const long long p = row / ((long long)g.Ho * g.Ho);
    const int r = (int)(row - p * ((long long)g.Ho * g.Ho));
    const int a = r / g.Ho, a2 = r - a * g.Ho;
    const int h0 = g.sh * a - g.loh, h20 = g.sh * a2 - g.loh;
    bool vi[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) vi[i] = h0 + i >= 0 && h0 + i < g.Hi && h20 + i >= 0 && h20 + i < g.Hi;
    const long long rb = p * per_i + ((long long)h0 * g.Hi + h20) * WW;
    for (int e = rw.e0; e < ww; e += rw.G) {
      const int b = e / g.Wo, b2 = e - b * g.Wo;
      const int w0 = g.sw * b - g.low, w20 = g.sw * b2 - g.low;
      bool vj[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) vj[j] = w0 + j >= 0 && w0 + j < g.Wi && w20 + j >= 0 && w20 + j < g.Wi;
      const long long eb = rb + (w0 * g.Wi + w20);
      T ak = 0, at = 0;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
          if (vi[i] && vj[j]) {
            ak += __ldg(inK + eb + (i * SI + j * SJ));
            if (has_t) at += __ldg(inT + eb + (i * SI + j * SJ));
          }
      const T k = fma_t(ak, scale, shift);
      outK[row * ww + e] = k;
      if (has_t) outT[row * ww + e] = fma_t(at, scale, (T)0) + k;
    }
  }
}

// ---- AvgPool rule: linear.py:3499-3572 (independent offsets on both members) --
struct PoolGeom {
  int Hi, Wi, Ho, Wo, wh, ww, sh, sw, loh, low, circular, normalize_edges;
  int sum;  // SumPool (linear.py:1503): no division
};

template <typename T>
__global__ void k_pool(const T* __restrict__ in, T* __restrict__ out, long long P, PoolGeom g) {
  const long long per_i = (long long)g.Hi * g.Hi * g.Wi * g.Wi;
  const int ww = g.Wo * g.Wo;
  RowWalk rw(P, g.Ho, g.Wo);
  // window members inside the input along one axis (all of them under CIRCULAR padding)
  auto count = [&](int o, int s, int win, int lo, int n) {
    int c = 0;
    for (int i = 0; i < win; ++i) {
      const int v = s * o + i - lo;
      if (g.circular || (v >= 0 && v < n)) ++c;
    }
    return c;
  };
  for (long long row = rw.row; row < rw.nrows; row += rw.stride) {
    const long long p = row / ((long long)g.Ho * g.Ho);
    const int r = (int)(row - p * ((long long)g.Ho * g.Ho));
    const int a = r / g.Ho, a2 = r - a * g.Ho;
    const T* src = in + p * per_i;
    T* orow = out + row * ww;
    const int cnt_h = count(a, g.sh, g.wh, g.loh, g.Hi), cnt_h2 = count(a2, g.sh, g.wh, g.loh, g.Hi);
    for (int e = rw.e0; e < ww; e += rw.G) {
      const int b = e / g.Wo, b2 = e - b * g.Wo;
      T acc = 0;
      for (int i = 0; i < g.wh; ++i) {
        int h = g.sh * a + i - g.loh;
        if (g.circular) h = ((h % g.Hi) + g.Hi) % g.Hi;
        if (h < 0 || h >= g.Hi) continue;
        for (int i2 = 0; i2 < g.wh; ++i2) {
          int h2 = g.sh * a2 + i2 - g.loh;
          if (g.circular) h2 = ((h2 % g.Hi) + g.Hi) % g.Hi;
          if (h2 < 0 || h2 >= g.Hi) continue;
          for (int j = 0; j < g.ww; ++j) {
            int w = g.sw * b + j - g.low;
            if (g.circular) w = ((w % g.Wi) + g.Wi) % g.Wi;
            if (w < 0 || w >= g.Wi) continue;
            const T* rowp = src + (((long long)h * g.Hi + h2) * g.Wi + w) * g.Wi;
            for (int j2 = 0; j2 < g.ww; ++j2) {
              int w2 = g.sw * b2 + j2 - g.low;
              if (g.circular) w2 = ((w2 % g.Wi) + g.Wi) % g.Wi;
              if (w2 < 0 || w2 >= g.Wi) continue;
              acc += rowp[w2];
            }
          }
        }
      }
      T norm = (T)((long long)g.wh * g.wh * g.ww * g.ww);
      if (g.normalize_edges)
        norm = (T)((long long)cnt_h * cnt_h2 * count(b, g.sw, g.ww, g.low, g.Wi) * count(b2, g.sw, g.ww, g.low, g.Wi));
      orow[e] = g.sum ? acc : acc / norm;
    }
  }
}

// ---- diagonal variances: requirements.py:1057-1074 ----------------------------
// q[n,h,w] = cov[n,h,h,w,w]
template <typename T>
__global__ void k_diag(const T* __restrict__ cov, T* __restrict__ q, long long n, int H, int W) {
  const long long total = n * H * W;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long s = idx / ((long long)H * W);
    int r = (int)(idx % ((long long)H * W));
    int h = r / W, w = r % W;
    q[idx] = cov[((((long long)s * H + h) * H + h) * W + w) * W + w];
  }
}

// ---- activations ---------------------------------------------------------------
// ABRelu: elementwise.py:444-455;  Erf: elementwise.py:84-93 + kernel.py:426-439.
struct ActParams {
  int kind;       // NTK_OP_ABRELU, NTK_OP_ERF, NTK_OP_GELU, NTK_OP_SIN (a, b, c) or NTK_OP_RBF (a = gamma)
  double a, b, c;
};

__device__ __forceinline__ float exp_t(float a) { return expf(a); }
__device__ __forceinline__ double exp_t(double a) { return exp(a); }

// Gelu (elementwise.py:225-240): k = nngp, q1 / q2 the diagonal variances of the two members.
template <typename T>
__device__ __forceinline__ void gelu_point(T k, T q1, T q2, T& k_out, T& dot) {
  const T prod = q1 * q2, prod_plus_1 = (q1 + (T)1) * (q2 + (T)1);
  const T delta_squared = prod_plus_1 - k * k;
  const T delta = sqrt_t(max_t(delta_squared, (T)0));
  const T angles = atan2_t(k, delta);
  const T two_pi = (T)2 * Consts<T>::pi();
  T nk = (k * k + prod * delta_squared) / (prod_plus_1 * delta);
  nk = (nk + k * angles) / two_pi + (T)0.25 * k;
  T first = (T)1 / delta_squared + ((T)1 - prod) / prod_plus_1 + (T)1;
  first *= k / delta / two_pi;
  dot = first + (T)0.25 + angles / two_pi;
  k_out = nk;
}

// Sin(a, b, c) (elementwise.py:294-300): sum_ = q1 + q2
template <typename T>
__device__ __forceinline__ void sin_point(T k, T sum_, T half_a2, T b2, T cos2c, T& k_out, T& dot) {
  const T s1 = exp_t(b2 * ((T)-0.5 * sum_ + k));
  const T s2 = exp_t(b2 * ((T)-0.5 * sum_ - k)) * cos2c;
  k_out = half_a2 * (s1 - s2);
  dot = half_a2 * b2 * (s1 + s2);
}

// Rbf(gamma) (elementwise.py:375-379)
template <typename T>
__device__ __forceinline__ void rbf_point(T k, T sum_, T gamma, T& k_out, T& dot) {
  k_out = exp_t(gamma * (-sum_ + (T)2 * k));
  dot = (T)2 * gamma * k_out;
}

template <typename T>
__device__ __forceinline__ void abrelu_point(T k, T prod, T coef_s, T half_ab, T& k_out, T& dot) {
  // s = sqrt(max(prod - k^2, 0)); theta = atan2(s, k) (pi/2 at (0,0));
  // dot = (a^2+b^2)/2 - (a-b)^2/(2 pi) * theta;  k' = (a-b)^2/(2 pi) * s + dot * k
  T s = sqrt_t(max_t(fma_t(-k, k, prod), (T)0));
  T theta = (s == (T)0 && k == (T)0) ? Consts<T>::pi() / 2 : atan2_t(s, k);
  dot = half_ab - coef_s * theta;
  k_out = fma_t(dot, k, coef_s * s);
}

template <typename T>
__device__ __forceinline__ void erf_point(T k, T prod, T& k_out, T& dot) {
  // k, prod already include the b^2 input scale: prod = (1+2 q1)(1+2 q2).
  T s = sqrt_t(max_t(fma_t((T)-4 * k, k, prod), (T)0));
  const T f = (T)2 / Consts<T>::pi();
  k_out = f * atan2_t((T)2 * k, s);
  dot = (T)2 * f / s;
}

// In-place on K (and Tt when non-null).  q1/q2 are the diagonal maps of cov1/cov2
// taken BEFORE this activation.  `stab` (device scalar, nullable) is the
// do_stabilize factor of elementwise.py:430-436.
template <typename T, int VEC>
struct alignas(sizeof(T) * VEC) ActPack {
  T v[VEC];
};

// VEC = 4 (W % 4 == 0): a lane owns 4 consecutive w' (one 16 / 32-byte access per tensor); VEC = 1 otherwise.
template <typename T, int VEC>
__global__ void k_act(T* __restrict__ K, T* __restrict__ Tt, const T* __restrict__ q1,
                      const T* __restrict__ q2, long long P, PairMap pm, int H, int W, ActParams ap,
                      const T* __restrict__ stab) {
  using Pack = ActPack<T, VEC>;
  const T a = (T)ap.a, b = (T)ap.b;
  const T coef_s = (a - b) * (a - b) / ((T)2 * Consts<T>::pi());
  const T half_ab = (a * a + b * b) / (T)2;
  const T bb = b * b, aa = a * a, cc = (T)(ap.c * ap.c);
  const T cos2c = ap.kind == NTK_OP_SIN ? (T)cos(2.0 * ap.c) : (T)0;
  T factor = (T)1;
  if (stab) factor = max_t(*stab, (T)1e-12);
  const bool has_t = Tt != nullptr;
  // one element: k, t in / out
  auto point = [&](T& k, T& t, T v1, T v2) {
    T ko, dot;
    if (ap.kind == NTK_OP_ABRELU) {
      if (stab) {
        k /= factor;
        v1 /= factor;
        v2 /= factor;
      }
      abrelu_point<T>(k, v1 * v2, coef_s, half_ab, ko, dot);
      if (stab) ko *= factor;
      k = ko;
      t *= dot;
    } else if (ap.kind == NTK_OP_ERF) {
      k *= bb;
      T prod = ((T)1 + (T)2 * bb * v1) * ((T)1 + (T)2 * bb * v2);
      erf_point<T>(k, prod, ko, dot);
      k = fma_t(aa, ko, cc);
      t = aa * (bb * t * dot);
    } else if (ap.kind == NTK_OP_LAYERNORM) {
      // linear.py:2566-2584: every kernel is divided by sqrt((eps + q1)(eps + q2)); a = eps
      const T inv = (T)1 / sqrt_t((a + v1) * (a + v2));
      k = k * inv;
      t *= inv;
    } else {
      if (ap.kind == NTK_OP_GELU)
        gelu_point<T>(k, v1, v2, ko, dot);
      else if (ap.kind == NTK_OP_SIN)
        sin_point<T>(k, v1 + v2, aa / (T)2, bb, cos2c, ko, dot);
      else
        rbf_point<T>(k, v1 + v2, a, ko, dot);
      k = ko;
      t *= dot;
    }
  };
  const int items = W * W / VEC;
  RowWalk rw(P, H, W, items);
  for (long long row = rw.row; row < rw.nrows; row += rw.stride) {
    const long long p = row / ((long long)H * H);
    const int r = (int)(row - p * ((long long)H * H));
    const int h = r / H, h2 = r - h * H;
    int i, j;
    pm.ij(p, i, j);
    const T* q1row = q1 + ((long long)i * H + h) * W;
    const T* q2row = q2 + ((long long)j * H + h2) * W;
    Pack* krow = reinterpret_cast<Pack*>(K + row * (long long)(W * W));
    Pack* trow = has_t ? reinterpret_cast<Pack*>(Tt + row * (long long)(W * W)) : nullptr;
    for (int e = rw.e0; e < items; e += rw.G) {
      const int w = (e * VEC) / W, w2 = e * VEC - w * W;
      const T v1 = q1row[w];
      const Pack v2 = *reinterpret_cast<const Pack*>(q2row + w2);
      Pack k = krow[e], t;
      if (has_t) t = trow[e];
#pragma unroll
      for (int c = 0; c < VEC; ++c) {
        T tt = has_t ? t.v[c] : (T)0;
        point(k.v[c], tt, v1, v2.v[c]);
        t.v[c] = tt;
      }
      krow[e] = k;
      if (has_t) trow[e] = t;
    }
  }
}

// max |x| over a tensor -> *out (out must be zeroed first); values are >= 0 so the
// float bit pattern orders like an unsigned integer.
template <typename T>
__global__ void k_absmax(const T* __restrict__ x, long long n, T* __restrict__ out) {
  T m = 0;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n;
       idx += (long long)gridDim.x * blockDim.x)
    m = max_t(m, abs_t(x[idx]));
  for (int o = 16; o > 0; o >>= 1) m = max_t(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) {
    if (sizeof(T) == 4)
      atomicMax(reinterpret_cast<unsigned int*>(out), __float_as_uint((float)m));
    else
      atomicMax(reinterpret_cast<unsigned long long*>(out),
                (unsigned long long)__double_as_longlong((double)m));
  }
}

// ---- reductions to [P]: GlobalAvgPool linear.py:1771-1801, Flatten 1865-1899 ----
// One block per pair; fixed-order tree reduction (deterministic).
template <typename T, bool kFlatten>
__global__ void k_reduce_spatial(const T* __restrict__ in, T* __restrict__ out, long long P, int H,
                                 int W, int sum = 0) {
  __shared__ T sm[kThreads / 32];
  const long long per = (long long)H * H * W * W;
  for (long long p = blockIdx.x; p < P; p += gridDim.x) {
    const T* src = in + p * per;
    T acc = 0;
    if (kFlatten) {
      for (int t = threadIdx.x; t < H * W; t += blockDim.x) {
        int h = t / W, w = t % W;
        acc += src[(((long long)h * H + h) * W + w) * W + w];
      }
    } else {
      for (long long t = threadIdx.x; t < per; t += blockDim.x) acc += src[t];
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
      T v = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : (T)0;
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (threadIdx.x == 0) out[p] = sum ? v : v / (T)(kFlatten ? (long long)H * W : per);
    }
    __syncthreads();
  }
}

// ---- Dense: linear.py:899-926 ----------------------------------------------------
// K <- w2 K + b2;  T <- K_new + w2 T   (T == nullptr: skip; t_zero: T_in == 0)
template <typename T>
__global__ void k_dense(T* __restrict__ K, T* __restrict__ Tt, long long n, T w2, T b2, int t_zero) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n;
       idx += (long long)gridDim.x * blockDim.x) {
    T k = fma_t(w2, K[idx], b2);
    K[idx] = k;
    if (Tt) Tt[idx] = t_zero ? k : fma_t(w2, Tt[idx], k);
  }
}

// ---- fully-connected networks in one launch -------------------------------------------------
// For [N, d] inputs every op after the input Gram is elementwise on the [n1, n2] matrix and on the two
// per-sample variance vectors, so the whole Dense / ABRelu / Erf chain (BASELINE configs[0])
// runs per entry in registers: K0[i,j], q-chains of samples i and j -> nngp[i,j], ntk[i,j].  The arithmetic is that of
// k_dense / k_act (same expressions, same order), so the result is bit-identical to the per-op path.
constexpr int kMaxFcnOps = 48;
struct FcnProg {
  int n;
  int kind[kMaxFcnOps];      // NTK_OP_DENSE | NTK_OP_ABRELU | NTK_OP_ERF
  int has_bias[kMaxFcnOps];  // Dense
  double f0[kMaxFcnOps], f1[kMaxFcnOps], f2[kMaxFcnOps];
};

// Per-sample variance chain: q evolves on its own (Dense: w2 q + b2; activation: the closed form on (q, q, q)),
// and its value IN FRONT of every activation is what the cross entries need.  qs: [n_act][n].
template <typename T>
__global__ void k_fcn_qchain(const T* __restrict__ c, int n, const FcnProg prog, T* __restrict__ qs) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    T q = c[i];
    int act = 0;
    for (int o = 0; o < prog.n; ++o) {
      if (prog.kind[o] == NTK_OP_DENSE) {
        const T w2 = (T)prog.f0[o], b2 = (T)(prog.has_bias[o] ? prog.f1[o] : 0.0);
        q = fma_t(w2, q, b2);
      } else if (prog.kind[o] == NTK_OP_ABRELU) {
        qs[(long long)act++ * n + i] = q;
        const T a = (T)prog.f0[o], b = (T)prog.f1[o];
        const T coef_s = (a - b) * (a - b) / ((T)2 * Consts<T>::pi());
        const T half_ab = (a * a + b * b) / (T)2;
        T d, dot;
        abrelu_point<T>(q, q * q, coef_s, half_ab, d, dot);
        q = d;
      } else {
        qs[(long long)act++ * n + i] = q;
        const T a = (T)prog.f0[o], b = (T)prog.f1[o];
        const T bb = b * b, aa = a * a, cc = (T)(prog.f2[o] * prog.f2[o]);
        T d, dot;
        erf_point<T>(q * bb, ((T)1 + (T)2 * bb * q) * ((T)1 + (T)2 * bb * q), d, dot);
        q = fma_t(aa, d, cc);
      }
    }
  }
}

template <typename T>
__global__ void k_fcn_chain(const T* __restrict__ K0, const T* __restrict__ qs1, const T* __restrict__ qs2,
                            int t1, int t2, long long n1_all, long long n2_all, const FcnProg prog,
                            T* __restrict__ nngp, T* __restrict__ ntk, long long ld) {
  const long long P = (long long)t1 * t2;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < P;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx / t2), j = (int)(idx % t2);
    T k = K0[idx], t = (T)0;
    bool t_zero = true;
    int act = 0;
    for (int o = 0; o < prog.n; ++o) {
      if (prog.kind[o] == NTK_OP_DENSE) {  // linear.py:899-926
        const T w2 = (T)prog.f0[o], b2 = (T)(prog.has_bias[o] ? prog.f1[o] : 0.0);
        k = fma_t(w2, k, b2);
        t = t_zero ? k : fma_t(w2, t, k);
        t_zero = false;
      } else {
        const T q1 = qs1[act * n1_all + i], q2 = qs2[act * n2_all + j];
        ++act;
        if (prog.kind[o] == NTK_OP_ABRELU) {  // elementwise.py:444-455
          const T a = (T)prog.f0[o], b = (T)prog.f1[o];
          const T coef_s = (a - b) * (a - b) / ((T)2 * Consts<T>::pi());
          const T half_ab = (a * a + b * b) / (T)2;
          T ko, dot;
          abrelu_point<T>(k, q1 * q2, coef_s, half_ab, ko, dot);
          k = ko;
          t *= dot;
        } else {  // Erf: elementwise.py:84-93 + kernel.py:426-439
          const T a = (T)prog.f0[o], b = (T)prog.f1[o];
          const T bb = b * b, aa = a * a, cc = (T)(prog.f2[o] * prog.f2[o]);
          T ko, dot;
          erf_point<T>(k * bb, ((T)1 + (T)2 * bb * q1) * ((T)1 + (T)2 * bb * q2), ko, dot);
          k = fma_t(aa, ko, cc);
          t = aa * (bb * t * dot);
        }
      }
    }
    nngp[(long long)i * ld + j] = k;
    if (ntk) ntk[(long long)i * ld + j] = t;
  }
}

// ---- FanInSum: branching.py:87-93 ---------------------------------------------------
template <typename T>
__global__ void k_add(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out,
                      long long n) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n;
       idx += (long long)gridDim.x * blockDim.x)
    out[idx] = a[idx] + b[idx];
}

template <typename T>
__global__ void k_fill(T* __restrict__ out, long long n, T v) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n;
       idx += (long long)gridDim.x * blockDim.x)
    out[idx] = v;
}

// ---- tile -> result matrix --------------------------------------------------------
// src: [t1, t2, per] tile; dst: pair (r0+i, c0+j) at ((r0+i)*ld + (c0+j))*per.
template <typename T>
__global__ void k_scatter(const T* __restrict__ src, T* __restrict__ dst, int t1, int t2,
                          long long per, long long ld, int r0, int c0) {
  const long long total = (long long)t1 * t2 * per;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long p = idx / per;
    long long e = idx % per;
    int i = (int)(p / t2), j = (int)(p % t2);
    dst[((long long)(r0 + i) * ld + (c0 + j)) * per + e] = src[idx];
  }
}

}  // namespace ntk
