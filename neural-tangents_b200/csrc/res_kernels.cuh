// Fused kernels for residual networks built from
//   Conv 3x3 / stride 1|2 / SAME,  ABRelu,  FanOut -> parallel(main, shortcut) -> FanInSum
// (WideResNet, `README.md:192-222`; rules `_src/stax/linear.py:3341-3378`,
// `_src/stax/elementwise.py:444-455`, `_src/stax/branching.py:55-117`).
//
// Same circular-shear column structure as fused_kernels.cuh, plus one more consequence of it:
//   a stride-2 SAME 3x3 conv only reads the EVEN (ch, cw) columns of its input (output column D
//   corresponds to input column 2D) and equals the stride-1 box filter sampled at odd (h, w).
// Conv layers and residual adds never mix columns, so a network whose only spatial reductions
// are strided convs needs, at each resolution, only the columns that are multiples of
// `cws` = product of the strides still to come: 1/16 of the 32x32 tensors and 1/4 of the 16x16
// ones for a WideResNet.  Tensors between kernels are stored compactly as
//   [pair][kh][h][w][kw],  ch = cws*kh, cw = cws*kw,  kh, kw < S/cws.
// A thread group covers `cws` ch-columns at once (the S cw-lanes hold cws sub-columns of S/cws
// cw values each), so all lanes stay busy.
//
// One kernel launch runs one of (flags in ResArgs):
//   stem           x -> conv -> Z                                   (FROM_X, 1 unit)
//   identity block Z -> relu -> conv -> relu -> conv -> + Z         (2 units, RES_INPUT)
//   conv-shortcut  Z -> [relu -> conv -> relu -> conv] + conv_s(Z)  (2 units + side, RES_SIDE)
//   strided half A Z -> [relu -> conv/2], conv_s/2(Z)               (1 unit + side, REPI_SUB)
//   strided half B Y -> relu -> conv -> + S                          (1 unit, RES_STREAM)
// with a STORE / stride-2 SUBsample / GlobalAvgPool epilogue.  The per-sample variances come from
// `k_qprog`, which runs the same network on the diagonal column only (no pools => the diagonal is
// self-contained, no self-pair pipeline).
#pragma once

#include "fused_kernels.cuh"

namespace ntk {

enum { RES_NONE = 0, RES_INPUT = 1, RES_STREAM = 2, RES_SIDE = 3 };
enum { REPI_STORE = 0, REPI_SUB = 1, REPI_GAP = 2 };

template <typename T>
struct ResArgs {
  const T* x1;
  const T* x2;
  const T* inK;
  const T* inT;
  const T* resK;  // RES_STREAM: second input, same geometry as the output of unit 0
  const T* resT;
  T* outK;
  T* outT;
  T* out2K;  // REPI_SUB with a side conv: the shortcut output
  T* out2T;
  const T* qm1;  // [n][n_act][S][S][2]
  const T* qm2;
  long long P;
  int n2, self, tri;
  int cws;      // column stride at this resolution
  int act_in;   // apply ABRelu (q-map act_id[0]) to the input rows of unit 0
  int n_units;  // 1 or 2 chained convs (unit 1 always has an ABRelu, q-map act_id[1], in front)
  int side;     // shortcut conv on the raw input rows
  int res;      // RES_*
  int epi;      // REPI_*
  long long qm_stride;   // V2 entries per sample in qm1/qm2
  long long act_off[2];  // V2 offset (inside a sample) of the q-map in front of unit 0 / unit 1
  T in_scale;    // FROM_X: alpha/C of unit 0
  T raw_scale;   // alpha of unit 0 when it has no activation in front and the input is LOADed
  T side_scale;  // alpha of the side conv
  T side_bias;
  T epi_scale;
  FLayer<T> lp[2];  // ABRelu constants in front of unit u (pre-scaled by unit u's alpha); bias of unit u
};

// fp32: 8 consecutive w per thread; fp64: 4 -- the three rings of 8 doubles for K and T alone are 192 registers, which
// spilled ~1 KB per thread in round 1 (k_res<double>: 680-712 byte stack frames); with 4 the kernel fits and twice as
// many warps cover the DFMA latency.
template <int S, int W = 8>
struct ResGeom {
  static constexpr int WPT = W;
  static constexpr int TPP = S * S / WPT;
  static constexpr int NT = TPP < 128 ? 128 : TPP;
  static constexpr int GROUPS = NT / TPP;
  static constexpr int LPG = TPP < 32 ? TPP : 32;
  static constexpr int NWB = S / WPT;
  static constexpr int LW = LPG / NWB;
};

// ERF: the network contains Erf activations (a runtime branch per activation row picks the closed
// form); pure-ABRelu networks run the ERF = false instantiation, which carries no Erf code.
// fp32: 8 w per thread for ABRelu networks, 4 for networks with Erf (measured on B200 with the input ring, WideResNet
// block 96 x 96: Relu 436.6 k entries/s at 8 w vs 376.1 k at 4 w; Erf 320.6 k at 8 w vs 349.3 k at 4 w -- the Erf rows carry
// more live values per element, at 8 w the kernel sits at 255 registers).  fp64: 4 w; 2 w (16 warps per SM, NTK_RES_WPT_F64=2) measured
// slower for ABRelu (168.8 k vs 186.1 k) and equal for Erf (154.8 k vs 153.5 k).
template <typename T, bool ERF>
struct ResWpt {
#ifndef NTK_RES_WPT_F64
#define NTK_RES_WPT_F64 4
#endif
  static constexpr int value = sizeof(T) == 8 ? NTK_RES_WPT_F64 : (ERF ? 4 : 8);
};

// Input rows of a LOADing kernel travel through a per-thread ring in shared memory, filled with cp.async one marched row
// ahead (round 2: the direct __ldg rows were consumed at once -- 43 % of the warp stalls were long-scoreboard waits on
// them); the ring keeps the last kResRing rows, so the identity shortcut (row t - 2) is read from it instead of from L2.
constexpr int kResRing = 4;

template <int BYTES>
__device__ __forceinline__ void res_cp_async(void* dst_smem, const void* src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(src), "n"(BYTES) : "memory");
}

template <typename T, int S, int IN, bool NTK, int CIN, bool ERF>
__global__ void __launch_bounds__((ResGeom<S, ResWpt<T, ERF>::value>::NT),
                                  (sizeof(T) == 4 && ResWpt<T, ERF>::value == 4 ? 512 / ResGeom<S, ResWpt<T, ERF>::value>::NT : 1))
k_res(const ResArgs<T> a) {
  using G = ResGeom<S, ResWpt<T, ERF>::value>;
  using V2 = typename Vec2<T>::type;
  constexpr int WPT = G::WPT, TPP = G::TPP, LPG = G::LPG, NWB = G::NWB, LW = G::LW;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int XS1 = IN == IN_FROM_X ? S * S * CIN : 0;
  constexpr int XS2 = IN == IN_FROM_X ? S * S * 4 : 0;
  constexpr int QM = 2 * S * S * 2;  // two activation layers
  constexpr int RG = IN == IN_LOAD ? kResRing * 2 * WPT * TPP : 0;  // input-row ring [slot][K|T][i][thread]
  constexpr int PER_GROUP = XS1 + XS2 + 2 * QM + RG;
  const int tid = threadIdx.x;
  const int grp = tid / TPP, tg = tid % TPP;
  T* sm = reinterpret_cast<T*>(smem_raw) + (size_t)grp * PER_GROUP;
  T* x1s = sm;
  T* x2s = x1s + XS1;
  V2* q1m = reinterpret_cast<V2*>(x2s + XS2);
  V2* q2m = q1m + 2 * S * S;
  T* ring = reinterpret_cast<T*>(q2m + 2 * S * S);

  const int lig = tg % LPG, wig = tg / 32;
  const int wblk = lig / LW, cwsub = lig % LW;
  const int c = wig * LW + cwsub;  // lane column index in [0, S)
  const int cws = a.cws;
  const int ncw = S / cws;         // columns kept per axis
  const int kw = c % ncw, chs = c / ncw;
  const int cw = kw * cws;
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

  if (IN == IN_FROM_X) {
    const T* g1 = a.x1 + (long long)si * S * S * CIN;
    const T* g2 = a.x2 + (long long)sj * S * S * CIN;
    for (int e = tg; e < S * S * CIN; e += TPP) x1s[e] = mul_rn(g1[e], a.in_scale);
    for (int e = tg; e < S * S; e += TPP) {
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) x2s[e * 4 + ci] = ci < CIN ? g2[e * CIN + ci] : (T)0;
    }
  }
  {
    const V2* g1 = reinterpret_cast<const V2*>(a.qm1) + (long long)si * a.qm_stride;
    const V2* g2 = reinterpret_cast<const V2*>(a.qm2) + (long long)sj * a.qm_stride;
    for (int u = 0; u < 2; ++u) {
      const bool used = u == 0 ? a.act_in != 0 : a.n_units == 2;
      if (!used) continue;
      for (int e = tg; e < S * S; e += TPP) {
        q1m[u * S * S + e] = g1[a.act_off[u] + e];
        q2m[u * S * S + e] = g2[a.act_off[u] + e];
      }
    }
  }
  if (TPP > 32)
    __syncthreads();
  else
    __syncwarp();

  T lk[WPT + 1];
  int off2[WPT];
#pragma unroll
  for (int i = 0; i <= WPT; ++i) {
    const int wl = w0 + i - 1, wr = w0 + i;
    lk[i] = (wl >= 0 && wr <= S - 1 && ((wl + cw) % S) != S - 1) ? (T)1 : (T)0;
  }
#pragma unroll
  for (int i = 0; i < WPT; ++i) off2[i] = ((w0 + i + cw) % S) * (int)sizeof(V2);
  const unsigned q2base = (unsigned)__cvta_generic_to_shared(q2m);

  // rings: two main units + the side conv
  T RK[3][2][WPT], RT[3][2][WPT];
#pragma unroll
  for (int u = 0; u < 3; ++u)
#pragma unroll
    for (int sl = 0; sl < 2; ++sl)
#pragma unroll
      for (int i = 0; i < WPT; ++i) {
        RK[u][sl][i] = (T)0;
        RT[u][sl][i] = (T)0;
      }
  T SBK[WPT], SBT[WPT];  // side-conv output of the previous step (RES_SIDE)
#pragma unroll
  for (int i = 0; i < WPT; ++i) {
    SBK[i] = (T)0;
    SBT[i] = (T)0;
  }

  const int ngroups = ncw / cws;  // column groups marched (cws columns each)
  const int nrows = ngroups * S;
  // compact tensor geometry: [p][kh][h][w][kw]
  const long long in_pair = (long long)ncw * S * S * ncw;
  // offsets inside one pair's tensor fit 32 bits (<= S^4 elements): int index arithmetic, one 64-bit add per access
  auto row_off = [&](int r, int& h, int& ch) -> int {
    // marched row r -> (column group, h); this thread's column index kh and ch
    const int cg = r / S;
    h = r % S;
    const int kh = cg * cws + chs;
    ch = kh * cws;
    return (kh * S + h) * S * ncw;
  };
  const T* inK = IN == IN_LOAD ? a.inK + p * in_pair : nullptr;
  const T* inT = (IN == IN_LOAD && NTK) ? a.inT + p * in_pair : nullptr;
  const T* rsK = a.res == RES_STREAM ? a.resK + p * in_pair : nullptr;
  const T* rsT = (a.res == RES_STREAM && NTK) ? a.resT + p * in_pair : nullptr;

  T gap_k = (T)0, gap_t = (T)0;
  const int n_units = a.n_units;

  auto load_row = [&](const T* bk, const T* bt, int r, T* dk, T* dt) {
    int h, ch;
    const int rc = r < 0 ? 0 : (r > nrows - 1 ? nrows - 1 : r);
    const int ro = row_off(rc, h, ch) + w0 * ncw + kw;
#pragma unroll
    for (int i = 0; i < WPT; ++i) {
      dk[i] = __ldg(bk + (ro + i * ncw));
      dt[i] = NTK ? __ldg(bt + (ro + i * ncw)) : (T)0;
    }
  };

  auto ring_at = [&](int r, int kt, int i) -> T* {
    return ring + (((r & (kResRing - 1)) * 2 + kt) * WPT + i) * TPP + tg;
  };
  // start the copy of input row r into its ring slot (one commit group per row)
  auto issue_row = [&](int r) {
    if (IN == IN_LOAD) {
      int h, ch;
      const int rc = r < 0 ? 0 : (r > nrows - 1 ? nrows - 1 : r);
      const int ro = row_off(rc, h, ch) + w0 * ncw + kw;
#pragma unroll
      for (int i = 0; i < WPT; ++i) {
        res_cp_async<sizeof(T)>(ring_at(r, 0, i), inK + (ro + i * ncw));
        if (NTK) res_cp_async<sizeof(T)>(ring_at(r, 1, i), inT + (ro + i * ncw));
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
  };
  auto ring_row = [&](int r, T* dk, T* dt) {
#pragma unroll
    for (int i = 0; i < WPT; ++i) {
      dk[i] = *ring_at(r, 0, i);
      dt[i] = NTK ? *ring_at(r, 1, i) : (T)0;
    }
  };

  // one conv unit: vertical taps of the old rows, horizontal taps of the new row (in place),
  // finish; returns the conv row (pre-bias) in ok/ot
  auto conv_unit = [&](const int u, const int slot, const T* pk, const T* pt, const bool has_t,
                       const T vU, const T vD, T* ok, T* ot) {
    T tk[WPT], tt[WPT];
#pragma unroll
    for (int i = 0; i < WPT; ++i) {
      tk[i] = fma_t(vU, RK[u][slot][i], RK[u][slot ^ 1][i]);
      tt[i] = has_t ? fma_t(vU, RT[u][slot][i], RT[u][slot ^ 1][i]) : (T)0;
    }
    T left = (T)0, right = (T)0, leftT = (T)0, rightT = (T)0;
    if (NWB > 1) {
      left = __shfl_up_sync(0xffffffffu, pk[WPT - 1], LW);
      right = __shfl_down_sync(0xffffffffu, pk[0], LW);
      if (has_t) {
        leftT = __shfl_up_sync(0xffffffffu, pt[WPT - 1], LW);
        rightT = __shfl_down_sync(0xffffffffu, pt[0], LW);
      }
    }
#pragma unroll
    for (int i = 0; i < WPT; ++i) {
      RK[u][slot][i] = hsum3<T>(i == 0 ? left : pk[i == 0 ? 0 : i - 1], pk[i],
                                i == WPT - 1 ? right : pk[i == WPT - 1 ? i : i + 1], lk[i], lk[i + 1]);
      if (has_t)
        RT[u][slot][i] = hsum3<T>(i == 0 ? leftT : pt[i == 0 ? 0 : i - 1], pt[i],
                                  i == WPT - 1 ? rightT : pt[i == WPT - 1 ? i : i + 1], lk[i], lk[i + 1]);
    }
#pragma unroll
    for (int i = 0; i < WPT; ++i) {
      ok[i] = fma_t(vD, RK[u][slot][i], tk[i]);
      ot[i] = has_t ? fma_t(vD, RT[u][slot][i], tt[i]) : (T)0;
    }
  };

  // ABRelu on a row (K, T) located at marched row r, with q-map layer u
  auto act_row = [&](const int u, int r, T* k, T* t) {
    int h, ch;
    const int rc = r < 0 ? 0 : (r > nrows - 1 ? nrows - 1 : r);
    row_off(rc, h, ch);
    const int h2 = (h + ch) % S;
    const V2* q1r = q1m + (u * S + h) * S + w0;
    const unsigned q2row = q2base + (unsigned)((u * S + h2) * S * (int)sizeof(V2));
    const T coef = a.lp[u].coef, half_ab = a.lp[u].half_ab, hab2 = a.lp[u].hab2;
#pragma unroll
    for (int i = 0; i < WPT; ++i) {
      const V2 qa = q1r[i];
      const V2 qb = lds_v2<T>(q2row + off2[i]);
      T ko, to;
      if (ERF && a.lp[u].kind == ACT_ERF)
        erf_act_point<T>(k[i], t[i], qa.x, qa.y, qb.x, qb.y, a.lp[u].e_in, a.lp[u].eA, a.lp[u].eT, a.lp[u].eC, ko, to);
      else
        act_point(k[i], t[i], qa.x, qa.y, qb.x, qb.y, coef, half_ab, hab2, ko, to);
      k[i] = ko;
      t[i] = to;
    }
  };

  auto vmasks = [&](int r, T& vU, T& vD) {
    int h, ch;
    const int rc = r < 0 ? 0 : (r > nrows - 1 ? nrows - 1 : r);
    row_off(rc, h, ch);
    const int h2 = (h + ch) % S;
    vU = (h > 0 && h2 != 0) ? (T)1 : (T)0;
    vD = (h < S - 1 && h2 != S - 1) ? (T)1 : (T)0;
  };

  auto step = [&](const int t, auto par_c) {
    constexpr int par = decltype(par_c)::value;
    T XK[WPT], XT[WPT];
    bool x_has_t = NTK && IN == IN_LOAD;
    // ---- input row r = t ----------------------------------------------------------------
    if (IN == IN_FROM_X) {
      int h, ch;
      row_off(t < nrows ? t : nrows - 1, h, ch);
      const int h2 = (h + ch) % S;
      const T* xa = x1s + (h * S + w0) * CIN;
      const char* xb = reinterpret_cast<const char*>(x2s + h2 * S * 4);
#pragma unroll
      for (int i = 0; i < WPT; ++i) {
        const T* b4 = reinterpret_cast<const T*>(xb + off2[i] * 2);
        T acc = mul_rn(xa[i * CIN], b4[0]);
#pragma unroll
        for (int ci = 1; ci < CIN; ++ci) acc = fma_t(xa[i * CIN + ci], b4[ci], acc);
        XK[i] = acc;
        XT[i] = (T)0;
      }
    } else {
      issue_row(t + 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");  // row t has landed (row t + 1 may be in flight)
      ring_row(t, XK, XT);
    }
    // ---- side (shortcut) conv on the raw rows: emits row t - 1 -----------------------------
    T SK[WPT], ST[WPT];
    T vU, vD;
    vmasks(t - 1, vU, vD);
    if (a.side) {
      conv_unit(2, par, XK, XT, x_has_t, vU, vD, SK, ST);
#pragma unroll
      for (int i = 0; i < WPT; ++i) {
        SK[i] = fma_t(SK[i], a.side_scale, a.side_bias);
        ST[i] = NTK ? (x_has_t ? fma_t(ST[i], a.side_scale, SK[i]) : SK[i]) : (T)0;
      }
    }
    // ---- unit 0 (optionally ABRelu in front): emits row t - 1 ------------------------------
    if (a.act_in) {
      act_row(0, t, XK, XT);
    } else if (IN == IN_LOAD) {
#pragma unroll
      for (int i = 0; i < WPT; ++i) {
        XK[i] = mul_rn(XK[i], a.raw_scale);
        XT[i] = mul_rn(XT[i], a.raw_scale);
      }
    }
    T YK[WPT], YT[WPT];
    conv_unit(0, par, XK, XT, x_has_t, vU, vD, YK, YT);
#pragma unroll
    for (int i = 0; i < WPT; ++i) {
      YK[i] = add_rn(YK[i], a.lp[0].bias);
      YT[i] = NTK ? (x_has_t ? add_rn(YT[i], YK[i]) : YK[i]) : (T)0;  // linear.py:1396-1398
    }
    int r_out = t - 1;
    if (n_units == 2) {
      // ---- unit 1: ABRelu on row t - 1, conv emits row t - 2 -------------------------------
      act_row(1, t - 1, YK, YT);
      T ZK[WPT], ZT[WPT];
      T vU1, vD1;
      vmasks(t - 2, vU1, vD1);
      conv_unit(1, par ^ 1, YK, YT, NTK, vU1, vD1, ZK, ZT);
#pragma unroll
      for (int i = 0; i < WPT; ++i) {
        YK[i] = add_rn(ZK[i], a.lp[1].bias);
        YT[i] = NTK ? add_rn(ZT[i], YK[i]) : (T)0;
      }
      r_out = t - 2;
    }
    // ---- residual (FanInSum, branching.py:87-93) -------------------------------------------
    if (a.res == RES_INPUT && IN == IN_LOAD) {
      T RKr[WPT], RTr[WPT];
      ring_row(r_out, RKr, RTr);  // the block input, one or two rows back: still in the ring
#pragma unroll
      for (int i = 0; i < WPT; ++i) {
        YK[i] = add_rn(YK[i], RKr[i]);
        if (NTK) YT[i] = add_rn(YT[i], RTr[i]);
      }
    } else if (a.res == RES_STREAM) {
      T RKr[WPT], RTr[WPT];
      load_row(rsK, rsT, r_out, RKr, RTr);
#pragma unroll
      for (int i = 0; i < WPT; ++i) {
        YK[i] = add_rn(YK[i], RKr[i]);
        if (NTK) YT[i] = add_rn(YT[i], RTr[i]);
      }
    } else if (a.res == RES_SIDE) {
#pragma unroll
      for (int i = 0; i < WPT; ++i) {
        YK[i] = add_rn(YK[i], SBK[i]);
        if (NTK) YT[i] = add_rn(YT[i], SBT[i]);
        SBK[i] = SK[i];
        SBT[i] = ST[i];
      }
    }
    // ---- epilogue on row r_out ----------------------------------------------------------------
    if (r_out >= 0 && r_out < nrows && live) {
      int h, ch;
      const int ro = row_off(r_out, h, ch);
      if (a.epi == REPI_STORE) {
        T* oK = a.outK + p * in_pair;
        T* oT = (NTK && a.outT) ? a.outT + p * in_pair : nullptr;  // outT == nullptr: the ntk equals the nngp (stem)
        const int base = ro + w0 * ncw + kw;
#pragma unroll
        for (int i = 0; i < WPT; ++i) {
          oK[base + i * ncw] = YK[i];
          if (oT) oT[base + i * ncw] = YT[i];
        }
      } else if (a.epi == REPI_SUB) {
        // stride-2 SAME conv == stride-1 box filter sampled at odd (h, w); output at S/2 keeps the
        // column indices (ch/2 = (cws/2) kh)
        if (h & 1) {
          constexpr int SO = S / 2;
          const long long out_pair = (long long)ncw * SO * SO * ncw;
          const int kh = ch / cws;
          const long long base = p * out_pair + (((long long)kh * SO + (h >> 1)) * SO) * ncw + kw;
#pragma unroll
          for (int i = 1; i < WPT; i += 2) {
            const long long o = base + (long long)((w0 + i) >> 1) * ncw;
            a.outK[o] = YK[i];
            if (NTK) a.outT[o] = YT[i];
            if (a.side) {
              a.out2K[o] = SK[i];
              if (NTK) a.out2T[o] = ST[i];
            }
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < WPT; ++i) {
          gap_k = add_rn(gap_k, YK[i]);
          if (NTK) gap_t = add_rn(gap_t, YT[i]);
        }
      }
    }
  };

  issue_row(0);
  const int nsteps0 = nrows + n_units;
  const int nsteps = nsteps0 + (nsteps0 & 1);
  for (int t0 = 0; t0 < nsteps; t0 += 2) {
    step(t0, std::integral_constant<int, 0>{});
    step(t0 + 1, std::integral_constant<int, 1>{});
  }

  if (a.epi == REPI_GAP) {
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
// Diagonal program: the network evaluated on the diagonal column (ch = cw = 0) of one sample,
// recording (q, 1/sqrt q) in front of every activation.  One CTA per sample.
// ---------------------------------------------------------------------------------------
enum { Q_CONV = 0, Q_ACT = 1, Q_COPY = 2, Q_ADD = 3, Q_INPUT = 4 };
constexpr int kQValid = 3;  // QOp::stride code of a VALID conv (linear.py:3341-3378 with no padding: every tap is inside)
constexpr int kQCirc = 4;   // 3x3 / stride 1 / CIRCULAR (linear.py:3158-3165: wrap-pad by the SAME amounts, then VALID): taps mod S
__host__ __device__ __forceinline__ int qconv_out_size(int S, int st) {
  return st == kQValid ? S - 2 : (st == kQCirc ? S : (S + st - 1) / st);
}
// A 3x3 / stride-2 / SAME conv on a size-S axis (lax.padtype_to_pads): out = ceil(S/2), total padding
// (out-1)*2 + 3 - S = 1 (S even: lo = 0, window centred at 2a+1) or 2 (S odd: lo = 1, centred at 2a).
__host__ __device__ __forceinline__ int strided_center_offset(int S) { return (S & 1) ? 0 : 1; }
struct QOp {
  int kind;
  int dst, src;  // image buffers 0..2
  int stride;    // Q_CONV: 1 or 2 (SAME), kQValid = 3x3 / stride 1 / VALID (out = S - 2, window centred at a + 1)
  int act_id;    // Q_ACT: q-map layer written
  double alpha, bias, kd0;
  int akind = ACT_ABRELU;  // Q_ACT: ABRelu(a = alpha, b = bias) or Erf(a = alpha, b = bias, c = erf_c)
  double erf_c = 0.0;
};
constexpr int kMaxQOps = 96;
template <typename T>
struct QProg {
  int n;
  int kind[kMaxQOps];
  signed char dst[kMaxQOps], src[kMaxQOps], stride[kMaxQOps];
  short act_id[kMaxQOps];
  T alpha[kMaxQOps], bias[kMaxQOps], kd0[kMaxQOps];
  T coef[kMaxQOps], hab2[kMaxQOps];  // Q_ACT on cross pairs (k_diagnet)
  signed char akind[kMaxQOps];       // Q_ACT: ACT_ABRELU | ACT_ERF
  T e_in[kMaxQOps], eA[kMaxQOps], eT[kMaxQOps], eC[kMaxQOps];  // Erf constants (FLayer)
};

template <typename T>
__device__ __forceinline__ FLayer<T> qprog_erf_layer(const QProg<T>& prog, int op) {
  FLayer<T> f;
  f.coef = f.half_ab = f.hab2 = f.bias = (T)0;
  f.kind = ACT_ERF;
  f.e_in = prog.e_in[op];
  f.eA = prog.eA[op];
  f.eT = prog.eT[op];
  f.eC = prog.eC[op];
  return f;
}

template <typename T>
__global__ void k_qprog(const T* __restrict__ x, int S0, int C, T in_scale, const QProg<T>* __restrict__ prog_g,
                        long long qm_stride, const long long* __restrict__ act_off, T* __restrict__ qm) {
  // three S0 x S0 image buffers + one scratch row-sum buffer
  extern __shared__ __align__(16) unsigned char qsm[];
  T* img = reinterpret_cast<T*>(qsm);
  T* scratch = img + 3 * S0 * S0;
  __shared__ int cur_S[3];
  const int n = blockIdx.x;
  const QProg<T>& prog = *prog_g;
  for (int op = 0; op < prog.n; ++op) {
    const int kind = prog.kind[op], d = prog.dst[op], sidx = prog.src[op];
    if (kind == Q_INPUT) {
      for (int e = threadIdx.x; e < S0 * S0; e += blockDim.x) {
        const T* xp = x + ((long long)n * S0 * S0 + e) * C;
        T v = mul_rn(mul_rn(xp[0], in_scale), xp[0]);
        for (int ci = 1; ci < C; ++ci) v = fma_t(mul_rn(xp[ci], in_scale), xp[ci], v);
        img[d * S0 * S0 + e] = v;
      }
      if (threadIdx.x == 0) cur_S[d] = S0;
    } else if (kind == Q_CONV) {
      const int S = cur_S[sidx];
      const T* P = img + sidx * S0 * S0;
      for (int e = threadIdx.x; e < S * S; e += blockDim.x) {
        const int h = e / S, w = e % S;
        const bool circ = prog.stride[op] == kQCirc;
        const T mL = (circ || w > 0) ? (T)1 : (T)0, mR = (circ || w < S - 1) ? (T)1 : (T)0;
        scratch[e] = hsum3<T>(P[h * S + (w > 0 ? w - 1 : (circ ? S - 1 : w))], P[e],
                              P[h * S + (w < S - 1 ? w + 1 : (circ ? 0 : w))], mL, mR);
      }
      __syncthreads();
      const int st = prog.stride[op];
      const int So = qconv_out_size(S, st);       // SAME: ceil(S / stride); VALID: S - 2
      const int o2 = strided_center_offset(S);    // window centre of a stride-2 conv: 2a+1 (S even), 2a (S odd)
      T* D = img + d * S0 * S0;
      for (int e = threadIdx.x; e < So * So; e += blockDim.x) {
        const int a_ = e / So, b_ = e % So;
        const int h = st == 2 ? 2 * a_ + o2 : (st == kQValid ? a_ + 1 : a_);
        const int w = st == 2 ? 2 * b_ + o2 : (st == kQValid ? b_ + 1 : b_);
        const bool circ = st == kQCirc;
        const T vU = (circ || h > 0) ? (T)1 : (T)0, vD = (circ || h < S - 1) ? (T)1 : (T)0;
        const T box = fma_t(vD, scratch[(h < S - 1 ? h + 1 : (circ ? 0 : h)) * S + w],
                            fma_t(vU, scratch[(h > 0 ? h - 1 : (circ ? S - 1 : h)) * S + w], scratch[h * S + w]));
        D[e] = fma_t(box, prog.alpha[op], prog.bias[op]);
      }
      __syncthreads();
      if (threadIdx.x == 0) cur_S[d] = So;
    } else if (kind == Q_ACT) {
      const int S = cur_S[d];
      typename Vec2<T>::type* out = reinterpret_cast<typename Vec2<T>::type*>(qm) +
                                    (long long)n * qm_stride + act_off[prog.act_id[op]];
      T* D = img + d * S0 * S0;
      for (int e = threadIdx.x; e < S * S; e += blockDim.x) {
        const T q = D[e];
        typename Vec2<T>::type o;
        if (prog.akind[op] == ACT_ERF) {
          const FLayer<T> f = qprog_erf_layer(prog, op);
          o.x = fma_t(f.e_in, q, (T)1);
          o.y = rsqrt_t(o.x);
          T ko, to;
          erf_act_point<T>(q, (T)0, o.x, o.y, o.x, o.y, f.e_in, f.eA, f.eT, f.eC, ko, to);
          D[e] = ko;
        } else {
          o.x = q;
          o.y = q > (T)0 ? rsqrt_t(q) : (T)0;
          D[e] = mul_rn(prog.kd0[op], q);
        }
        out[e] = o;
      }
    } else if (kind == Q_COPY) {
      const int S = cur_S[sidx];
      for (int e = threadIdx.x; e < S * S; e += blockDim.x) img[d * S0 * S0 + e] = img[sidx * S0 * S0 + e];
      if (threadIdx.x == 0) cur_S[d] = S;
    } else if (kind == Q_ADD) {
      const int S = cur_S[d];
      for (int e = threadIdx.x; e < S * S; e += blockDim.x)
        img[d * S0 * S0 + e] = add_rn(img[d * S0 * S0 + e], img[sidx * S0 * S0 + e]);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
// host side: recognise  Conv (ABRelu Conv[/s] ABRelu Conv [+ Conv[/s] shortcut] FanInSum)+
//            (GAP | AvgPool(SxS) Flatten)  Dense*
// ---------------------------------------------------------------------------------------
struct ResBlock {
  double w1, b1, w2, b2, ws, bs;  // W^2 and b^2 of conv1, conv2, shortcut conv
  double a1, c1, a2, c2;          // ABRelu (a, b) [Erf (a, b)] in front of conv1 / conv2
  int k1 = ACT_ABRELU, k2 = ACT_ABRELU;
  double g1 = 0, g2 = 0;          // Erf c
  int stride;                     // of conv1 (and of the shortcut conv)
  bool conv_shortcut;
};
struct ResPlan {
  bool ok = false;
  double w0 = 0, b0 = 0;  // stem conv
  std::vector<ResBlock> blocks;
  std::vector<ntk_op_t> dense_tail;
  int n_strided = 0;
};

inline ResPlan plan_resnet(const std::vector<ntk_op_t>& ops, int out_slot) {
  ResPlan plan;
  const int n = (int)ops.size();
  auto conv3 = [](const ntk_op_t& o, int stride) {
    return o.kind == NTK_OP_CONV && o.i[0] == 3 && o.i[1] == 3 && o.i[2] == stride && o.i[3] == stride &&
           o.i[4] == NTK_PAD_SAME;
  };
  auto is_act = [](const ntk_op_t& o) { return (o.kind == NTK_OP_ABRELU && o.i[0] == 0) || o.kind == NTK_OP_ERF; };
  if (n < 2 || !conv3(ops[0], 1) || ops[0].src != 0) return plan;
  plan.w0 = ops[0].f[0];
  plan.b0 = ops[0].i[5] ? ops[0].f[1] : 0.0;
  int z = ops[0].dst;
  int k = 1;
  while (k + 4 < n && is_act(ops[k]) && ops[k].src == z) {
    ResBlock b{};
    const ntk_op_t &r1 = ops[k], &c1 = ops[k + 1], &r2 = ops[k + 2], &c2 = ops[k + 3];
    const int stride = c1.kind == NTK_OP_CONV ? c1.i[2] : 0;
    if (!(stride == 1 || stride == 2) || !conv3(c1, stride) || c1.src != r1.dst) return ResPlan();
    if (!is_act(r2) || r2.src != c1.dst || !conv3(c2, 1) || c2.src != r2.dst) return ResPlan();
    b.a1 = r1.f[0];
    b.c1 = r1.f[1];
    b.a2 = r2.f[0];
    b.c2 = r2.f[1];
    if (r1.kind == NTK_OP_ERF) {
      b.k1 = ACT_ERF;
      b.g1 = r1.f[2];
    }
    if (r2.kind == NTK_OP_ERF) {
      b.k2 = ACT_ERF;
      b.g2 = r2.f[2];
    }
    b.w1 = c1.f[0];
    b.b1 = c1.i[5] ? c1.f[1] : 0.0;
    b.w2 = c2.f[0];
    b.b2 = c2.i[5] ? c2.f[1] : 0.0;
    b.stride = stride;
    int kk = k + 4;
    int sc_slot = z;
    if (ops[kk].kind == NTK_OP_CONV) {  // conv shortcut on the block input
      if (!conv3(ops[kk], stride) || ops[kk].src != z) return ResPlan();
      b.conv_shortcut = true;
      b.ws = ops[kk].f[0];
      b.bs = ops[kk].i[5] ? ops[kk].f[1] : 0.0;
      sc_slot = ops[kk].dst;
      ++kk;
    } else if (stride != 1) {
      return ResPlan();
    }
    if (kk >= n || ops[kk].kind != NTK_OP_FANINSUM || ops[kk].src != c2.dst || ops[kk].src2 != sc_slot)
      return ResPlan();
    z = ops[kk].dst;
    k = kk + 1;
    if (stride == 2) plan.n_strided++;
    plan.blocks.push_back(b);
  }
  if (plan.blocks.empty() || k >= n) return ResPlan();
  // tail: GAP, or AvgPool covering the whole (square) map followed by Flatten -- checked at run time
  if (ops[k].kind == NTK_OP_GAP && ops[k].i[0] == 0 && ops[k].src == z) {
    z = ops[k].dst;
    ++k;
  } else if (ops[k].kind == NTK_OP_AVGPOOL && ops[k].src == z && k + 1 < n && ops[k + 1].kind == NTK_OP_FLATTEN &&
             ops[k + 1].src == ops[k].dst && ops[k].i[4] == NTK_PAD_VALID && ops[k].i[0] == ops[k].i[1] &&
             (ops[k].i[5] & 2) == 0) {
    plan.n_strided |= ops[k].i[0] << 8;  // remember the pool window to validate against S
    z = ops[k + 1].dst;
    k += 2;
  } else {
    return ResPlan();
  }
  for (; k < n; ++k) {
    if (ops[k].kind != NTK_OP_DENSE || ops[k].src != z) return ResPlan();
    z = ops[k].dst;
    plan.dense_tail.push_back(ops[k]);
  }
  if (z != out_slot) return ResPlan();
  plan.ok = true;
  return plan;
}

template <typename T>
bool res_supported(const ResPlan& plan, int H, int W, int C) {
  if (!plan.ok || H != W || C != 3) return false;
  if (H != 32 && H != 16 && H != 8) return false;
  const int ns = plan.n_strided & 0xff, win = plan.n_strided >> 8;
  int S = H, cws = 1 << ns;
  for (int s = 0; s <= ns; ++s) {
    if (S < 8) return false;
    const int ncw = S / cws;
    if (ncw < 1 || ncw % cws != 0) return false;
    S /= 2;
    cws /= 2;
  }
  const int S_last = H >> ns;
  if (win && win != S_last) return false;  // AvgPool + Flatten must be a global mean
  return true;
}

template <typename T, bool NTK, bool ERF>
int launch_res(cudaStream_t stream, int64_t* launches, int S, bool from_x, const ResArgs<T>& a) {
  (*launches)++;
  auto go = [&](auto kern, int nt, int groups, size_t smem) -> int {
    NTK_TRY(ensure_dynamic_smem((const void*)kern, smem));
    const long long blocks = (a.P + groups - 1) / groups;
    kern<<<(unsigned)blocks, nt, smem, stream>>>(a);
    NTK_CUDA(cudaGetLastError());
    return NTK_OK;
  };
  auto smem_for = [&](int s, bool fx) {
    const int tpp = s * s / ResWpt<T, ERF>::value, nt = tpp < 128 ? 128 : tpp;
    const size_t per = (fx ? (size_t)s * s * 3 + (size_t)s * s * 4 : (size_t)kResRing * 2 * ResWpt<T, ERF>::value * tpp) +
                       2 * (size_t)(2 * s * s * 2);
    return per * (nt / tpp) * sizeof(T);
  };
#define NTK_RES_CASE(SS)                                                                              \
  if (S == SS) {                                                                                      \
    using G = ResGeom<SS, ResWpt<T, ERF>::value>;                                                          \
    if (from_x) return go(k_res<T, SS, IN_FROM_X, NTK, 3, ERF>, G::NT, G::GROUPS, smem_for(SS, true)); \
    return go(k_res<T, SS, IN_LOAD, NTK, 1, ERF>, G::NT, G::GROUPS, smem_for(SS, false));             \
  }
  NTK_RES_CASE(32)
  NTK_RES_CASE(16)
  NTK_RES_CASE(8)
#undef NTK_RES_CASE
  return fail(NTK_EUNSUPPORTED, "no residual kernel for S = %d", S);
}

template <typename T>
FLayer<T> res_act_consts(double a, double b, double alpha_next, double bias, int kind = ACT_ABRELU,
                         double erf_c = 0.0) {
  const double pi = 3.14159265358979323846, two_pi = 2.0 * pi;
  const double d = a - b;
  const double coef = alpha_next * d * d / two_pi, half_ab = alpha_next * (a * a + b * b) / 2.0;
  FLayer<T> f;
  f.coef = (T)coef;
  f.half_ab = (T)half_ab;
  f.hab2 = (T)(half_ab - coef * 1.57079632679489661923);
  f.bias = (T)bias;
  f.kind = kind;
  f.e_in = (T)(2.0 * b * b);
  f.eA = (T)(alpha_next * a * a * 2.0 / pi);
  f.eT = (T)(alpha_next * a * a * b * b * 4.0 / pi);
  f.eC = (T)(alpha_next * erf_c * erf_c);
  return f;
}

inline float host_kd0(float coef, float hab2) { return fmaf(coef, 1.57079632679489661923f, hab2); }
inline double host_kd0(double coef, double hab2) { return fma(coef, 1.5707963267948966, hab2); }

// Whole Gram block of a residual network.
template <typename T>
int res_gram(const ResPlan& plan, Arena& arena, cudaStream_t stream, int64_t* launches, const T* x1,
             int n1, const T* x2, int n2, bool symmetric, int S0, int C, bool want_ntk, T* out_nngp,
             T* out_ntk, long long ld, bool full_square) {
  const bool triangular = symmetric && !full_square && n1 == n2 && n1 > 1;
  const int ns = plan.n_strided & 0xff;
  const size_t nb = plan.blocks.size();
  bool any_erf = false;
  for (const ResBlock& B : plan.blocks) any_erf = any_erf || B.k1 == ACT_ERF || B.k2 == ACT_ERF;
  // ---- q-program --------------------------------------------------------------------------
  std::vector<long long> act_off;  // V2 offset of every activation's q-map inside a sample
  std::vector<int> act_S;
  QProg<T> qp{};
  auto push = [&](int kind, int dst, int src, int stride, int act_id, double alpha, double bias, T kd0,
                  const FLayer<T>* f = nullptr) {
    const int i = qp.n++;
    qp.coef[i] = (T)0;
    qp.hab2[i] = (T)0;
    qp.akind[i] = (signed char)(f ? f->kind : ACT_ABRELU);
    qp.e_in[i] = f ? f->e_in : (T)0;
    qp.eA[i] = f ? f->eA : (T)0;
    qp.eT[i] = f ? f->eT : (T)0;
    qp.eC[i] = f ? f->eC : (T)0;
    qp.kind[i] = kind;
    qp.dst[i] = (signed char)dst;
    qp.src[i] = (signed char)src;
    qp.stride[i] = (signed char)stride;
    qp.act_id[i] = (short)act_id;
    qp.alpha[i] = (T)alpha;
    qp.bias[i] = (T)bias;
    qp.kd0[i] = kd0;
  };
  if (5 + 7 * nb > (size_t)kMaxQOps) return fail(NTK_EUNSUPPORTED, "network too deep for the q-program");
  push(Q_INPUT, 0, 0, 1, 0, 1.0, 0.0, (T)0);
  push(Q_CONV, 0, 0, 1, 0, 1.0, plan.b0, (T)0);
  {
    int S = S0;
    long long off = 0;
    for (size_t b = 0; b < nb; ++b) {
      const ResBlock& B = plan.blocks[b];
      const FLayer<T> f1 = res_act_consts<T>(B.a1, B.c1, B.w1 / 9.0, B.b1, B.k1, B.g1);
      const FLayer<T> f2 = res_act_consts<T>(B.a2, B.c2, B.w2 / 9.0, B.b2, B.k2, B.g2);
      push(Q_COPY, 1, 0, 1, 0, 1.0, 0.0, (T)0);
      act_off.push_back(off);
      act_S.push_back(S);
      off += (long long)S * S;
      push(Q_ACT, 0, 0, 1, (int)act_off.size() - 1, 1.0, 0.0, host_kd0(f1.coef, f1.hab2), &f1);
      push(Q_CONV, 0, 0, B.stride, 0, 1.0, B.b1, (T)0);
      S /= B.stride;
      act_off.push_back(off);
      act_S.push_back(S);
      off += (long long)S * S;
      push(Q_ACT, 0, 0, 1, (int)act_off.size() - 1, 1.0, 0.0, host_kd0(f2.coef, f2.hab2), &f2);
      push(Q_CONV, 0, 0, 1, 0, 1.0, B.b2, (T)0);
      if (B.conv_shortcut) push(Q_CONV, 1, 1, B.stride, 0, B.ws / 9.0, B.bs, (T)0);
      push(Q_ADD, 0, 1, 1, 0, 1.0, 0.0, (T)0);
    }
    act_off.push_back(off);  // total
  }
  const long long qm_stride = act_off.back();
  QProg<T>* qp_d = (QProg<T>*)arena.alloc(sizeof(QProg<T>));
  long long* off_d = (long long*)arena.alloc(act_off.size() * sizeof(long long));
  T* qm1 = (T*)arena.alloc((size_t)n1 * qm_stride * 2 * sizeof(T));
  T* qm2 = symmetric ? qm1 : (T*)arena.alloc((size_t)n2 * qm_stride * 2 * sizeof(T));
  if (!qp_d || !off_d || !qm1 || !qm2) return fail(NTK_ENOMEM, "workspace too small for the q-maps");
  NTK_CUDA(cudaMemcpyAsync(qp_d, &qp, sizeof(qp), cudaMemcpyHostToDevice, stream));
  NTK_CUDA(cudaMemcpyAsync(off_d, act_off.data(), act_off.size() * sizeof(long long), cudaMemcpyHostToDevice, stream));
  NTK_CUDA(cudaStreamSynchronize(stream));  // qp / act_off live on the host stack
  const T in_scale = (T)(plan.w0 / 9.0 / (double)C);
  NTK_TRY(ensure_dynamic_smem((const void*)k_qprog<T>, (size_t)4 * S0 * S0 * sizeof(T)));
  for (int set = 0; set < (symmetric ? 1 : 2); ++set) {
    (*launches)++;
    k_qprog<T><<<set == 0 ? n1 : n2, 256, (size_t)4 * S0 * S0 * sizeof(T), stream>>>(
        set == 0 ? x1 : x2, S0, C, in_scale, qp_d, qm_stride, off_d, set == 0 ? qm1 : qm2);
    NTK_CUDA(cudaGetLastError());
  }

  // ---- tiles ---------------------------------------------------------------------------------
  const int cws0 = 1 << ns;
  const size_t per_pair = (size_t)(S0 / cws0) * (S0 / cws0) * S0 * S0 * (want_ntk ? 2 : 1) * sizeof(T);
  long long tile_pairs = (long long)n1 * n2;
  T* buf[3] = {nullptr, nullptr, nullptr};
  for (;;) {
    bool okb = true;
    for (int i = 0; i < 3; ++i) {
      buf[i] = (T*)arena.alloc((size_t)tile_pairs * per_pair);
      okb = okb && buf[i];
    }
    if (okb) break;
    for (int i = 0; i < 3; ++i) {
      if (buf[i]) arena.release(buf[i]);
      buf[i] = nullptr;
    }
    if (tile_pairs <= 1) return fail(NTK_ENOMEM, "workspace too small for one pair");
    tile_pairs = (tile_pairs + 1) / 2;
  }
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

  for (int r0 = 0; r0 < n1; r0 += t1) {
    const int a1 = std::min(t1, n1 - r0);
    const bool tri_tile = triangular && t2 == n2;
    for (int c0 = triangular ? r0 : 0; c0 < n2; c0 += t2) {
      const int a2 = tri_tile ? n2 - r0 : std::min(t2, n2 - c0);
      const long long P = tri_tile ? tri_prefix(a1, a2) : (long long)a1 * a2;
      int S = S0, cws = cws0;
      auto base_args = [&]() {
        ResArgs<T> a{};
        a.x1 = x1 + (size_t)r0 * S0 * S0 * C;
        a.x2 = x2 + (size_t)c0 * S0 * S0 * C;
        a.qm1 = qm1 + (size_t)r0 * qm_stride * 2;
        a.qm2 = qm2 + (size_t)c0 * qm_stride * 2;
        a.qm_stride = qm_stride;
        a.P = P;
        a.n2 = a2;
        a.self = 0;
        a.tri = tri_tile ? 1 : 0;
        a.in_scale = in_scale;
        a.side_scale = (T)1;
        a.raw_scale = (T)1;
        return a;
      };
      auto tsize = [&](int s, int cw_) { return (size_t)(s / cw_) * (s / cw_) * s * s; };
      auto run = [&](int Sx, bool from_x, const ResArgs<T>& a) -> int {
        if (any_erf) {
          if (want_ntk) return launch_res<T, true, true>(stream, launches, Sx, from_x, a);
          return launch_res<T, false, true>(stream, launches, Sx, from_x, a);
        }
        if (want_ntk) return launch_res<T, true, false>(stream, launches, Sx, from_x, a);
        return launch_res<T, false, false>(stream, launches, Sx, from_x, a);
      };
      // stem
      int cur = 0;  // index of the buffer holding the current block input Z
      {
        ResArgs<T> a = base_args();
        a.cws = cws;
        a.act_in = 0;
        a.n_units = 1;
        a.side = 0;
        a.res = RES_NONE;
        a.epi = REPI_STORE;
        a.lp[0].bias = (T)plan.b0;
        a.outK = buf[cur];
        // the ntk behind the first conv IS the nngp (linear.py:1396-1398 with ntk = 0): it is stored once and the first
        // block reads both of its inputs from the same tensor (half the stem's writes, half the first block's DRAM reads)
        a.outT = nullptr;
        NTK_TRY(run(S, true, a));
      }
      int act = 0;
      for (size_t b = 0; b < nb; ++b) {
        const ResBlock& B = plan.blocks[b];
        const bool last = b + 1 == nb;
        const FLayer<T> f1 = res_act_consts<T>(B.a1, B.c1, B.w1 / 9.0, B.b1, B.k1, B.g1);
        const FLayer<T> f2 = res_act_consts<T>(B.a2, B.c2, B.w2 / 9.0, B.b2, B.k2, B.g2);
        const int o1 = (cur + 1) % 3, o2 = (cur + 2) % 3;
        if (B.stride == 1) {
          ResArgs<T> a = base_args();
          a.cws = cws;
          a.inK = buf[cur];
          a.inT = want_ntk ? (b == 0 ? buf[cur] : buf[cur] + (size_t)P * tsize(S, cws)) : nullptr;
          a.act_in = 1;
          a.n_units = 2;
          a.side = B.conv_shortcut ? 1 : 0;
          a.res = B.conv_shortcut ? RES_SIDE : RES_INPUT;
          a.side_scale = (T)(B.ws / 9.0);
          a.side_bias = (T)B.bs;
          a.lp[0] = f1;
          a.lp[1] = f2;
          a.act_off[0] = act_off[act];
          a.act_off[1] = act_off[act + 1];
          if (last) {
            a.epi = REPI_GAP;
            a.outK = resK;
            a.outT = resT;
            a.epi_scale = (T)(1.0 / ((double)S * S * S * S));
          } else {
            a.epi = REPI_STORE;
            a.outK = buf[o1];
            a.outT = want_ntk ? buf[o1] + (size_t)P * tsize(S, cws) : nullptr;
          }
          NTK_TRY(run(S, false, a));
          cur = o1;
        } else {
          // half A at S: relu -> conv1/2 and the shortcut conv/2, both subsampled to S/2
          ResArgs<T> a = base_args();
          a.cws = cws;
          a.inK = buf[cur];
          a.inT = want_ntk ? (b == 0 ? buf[cur] : buf[cur] + (size_t)P * tsize(S, cws)) : nullptr;
          a.act_in = 1;
          a.n_units = 1;
          a.side = 1;
          a.res = RES_NONE;
          a.epi = REPI_SUB;
          a.side_scale = (T)(B.ws / 9.0);
          a.side_bias = (T)B.bs;
          a.lp[0] = f1;
          a.act_off[0] = act_off[act];
          const int S2 = S / 2, cws2 = cws / 2;
          a.outK = buf[o1];
          a.outT = want_ntk ? buf[o1] + (size_t)P * tsize(S2, cws2) : nullptr;
          a.out2K = buf[o2];
          a.out2T = want_ntk ? buf[o2] + (size_t)P * tsize(S2, cws2) : nullptr;
          NTK_TRY(run(S, false, a));
          // half B at S/2: relu -> conv2 -> + shortcut
          S = S2;
          cws = cws2;
          ResArgs<T> h = base_args();
          h.cws = cws;
          h.inK = buf[o1];
          h.inT = want_ntk ? buf[o1] + (size_t)P * tsize(S, cws) : nullptr;
          h.resK = buf[o2];
          h.resT = want_ntk ? buf[o2] + (size_t)P * tsize(S, cws) : nullptr;
          h.act_in = 1;
          h.n_units = 1;
          h.side = 0;
          h.res = RES_STREAM;
          h.lp[0] = f2;
          h.act_off[0] = act_off[act + 1];
          if (last) {
            h.epi = REPI_GAP;
            h.outK = resK;
            h.outT = resT;
            h.epi_scale = (T)(1.0 / ((double)S * S * S * S));
          } else {
            h.epi = REPI_STORE;
            h.outK = buf[cur];  // the old block input is dead now
            h.outT = want_ntk ? buf[cur] + (size_t)P * tsize(S, cws) : nullptr;
          }
          NTK_TRY(run(S, false, h));
          // cur stays: Z' was written into buf[cur]
        }
        act += 2;
      }
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
      if (tri_tile) break;
    }
  }
  if (triangular) {
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

// ---------------------------------------------------------------------------------------
// Pool-free networks ending in Flatten (the reference's `diagonal_spatial` fast path,
// `_src/stax/linear.py:3381-3437`, README.md:399-416): Flatten only reads the diagonal column
// (ch = cw = 0), and conv / activation / FanInSum never mix columns, so the whole cross-pair
// computation is the q-program run on K[h,h,w,w] (and T) of every pair.  One CTA per pair.
// ---------------------------------------------------------------------------------------
template <typename T, bool NTK>
__global__ void __launch_bounds__(128)
k_diagnet(const T* __restrict__ x1, const T* __restrict__ x2, int S0, int C, T in_scale,
          const QProg<T>* __restrict__ prog_g, long long qm_stride, const long long* __restrict__ act_off,
          const T* __restrict__ qm1, const T* __restrict__ qm2, long long P, int n2, int tri,
          T* __restrict__ outK, T* __restrict__ outT) {
  using V2 = typename Vec2<T>::type;
  extern __shared__ __align__(16) unsigned char dsm[];
  const int SS = S0 * S0;
  T* imgK = reinterpret_cast<T*>(dsm);       // 3 buffers
  T* imgT = imgK + 3 * SS;                   // 3 buffers
  T* scK = imgT + 3 * SS;
  T* scT = scK + SS;
  __shared__ int cur_S[3];
  __shared__ int has_t[3];
  __shared__ T red[2][4];
  const QProg<T>& prog = *prog_g;
  for (long long p = blockIdx.x; p < P; p += gridDim.x) {
    int si, sj;
    if (tri) {
      int off;
      tri_unrank(p, n2, si, off);
      sj = si + off;
    } else {
      si = (int)(p / n2);
      sj = (int)(p % n2);
    }
    const V2* q1 = reinterpret_cast<const V2*>(qm1) + (long long)si * qm_stride;
    const V2* q2 = reinterpret_cast<const V2*>(qm2) + (long long)sj * qm_stride;
    for (int op = 0; op < prog.n; ++op) {
      const int kind = prog.kind[op], d = prog.dst[op], sidx = prog.src[op];
      T* DK = imgK + d * SS;
      T* DT = imgT + d * SS;
      if (kind == Q_INPUT) {
        for (int e = threadIdx.x; e < SS; e += blockDim.x) {
          const T* xa = x1 + ((long long)si * SS + e) * C;
          const T* xb = x2 + ((long long)sj * SS + e) * C;
          T v = mul_rn(mul_rn(xa[0], in_scale), xb[0]);
          for (int ci = 1; ci < C; ++ci) v = fma_t(mul_rn(xa[ci], in_scale), xb[ci], v);
          DK[e] = v;
        }
        if (threadIdx.x == 0) {
          cur_S[d] = S0;
          has_t[d] = 0;
        }
      } else if (kind == Q_CONV) {
        const int S = cur_S[sidx];
        const bool ht = NTK && has_t[sidx];
        const T* PK = imgK + sidx * SS;
        const T* PT = imgT + sidx * SS;
        for (int e = threadIdx.x; e < S * S; e += blockDim.x) {
          const int h = e / S, w = e % S;
          const bool circ = prog.stride[op] == kQCirc;
          const T mL = (circ || w > 0) ? (T)1 : (T)0, mR = (circ || w < S - 1) ? (T)1 : (T)0;
          const int el = h * S + (w > 0 ? w - 1 : (circ ? S - 1 : w)), er = h * S + (w < S - 1 ? w + 1 : (circ ? 0 : w));
          scK[e] = hsum3<T>(PK[el], PK[e], PK[er], mL, mR);
          if (ht) scT[e] = hsum3<T>(PT[el], PT[e], PT[er], mL, mR);
        }
        __syncthreads();
        const int st = prog.stride[op];
        const int So = qconv_out_size(S, st);
        const int o2 = strided_center_offset(S);
        for (int e = threadIdx.x; e < So * So; e += blockDim.x) {
          const int a_ = e / So, b_ = e % So;
          const int h = st == 2 ? 2 * a_ + o2 : (st == kQValid ? a_ + 1 : a_);
          const int w = st == 2 ? 2 * b_ + o2 : (st == kQValid ? b_ + 1 : b_);
          const bool circ = st == kQCirc;
          const T vU = (circ || h > 0) ? (T)1 : (T)0, vD = (circ || h < S - 1) ? (T)1 : (T)0;
          const int eu = (h > 0 ? h - 1 : (circ ? S - 1 : h)) * S + w, ed = (h < S - 1 ? h + 1 : (circ ? 0 : h)) * S + w;
          const T k = fma_t(fma_t(vD, scK[ed], fma_t(vU, scK[eu], scK[h * S + w])), prog.alpha[op],
                            prog.bias[op]);
          DK[e] = k;
          if (NTK)
            DT[e] = ht ? fma_t(fma_t(vD, scT[ed], fma_t(vU, scT[eu], scT[h * S + w])), prog.alpha[op], k) : k;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
          cur_S[d] = So;
          has_t[d] = 1;
        }
      } else if (kind == Q_ACT) {
        const int S = cur_S[d];
        const V2* a1 = q1 + act_off[prog.act_id[op]];
        const V2* a2 = q2 + act_off[prog.act_id[op]];
        const bool ht = NTK && has_t[d];
        for (int e = threadIdx.x; e < S * S; e += blockDim.x) {
          const V2 qa = a1[e], qb = a2[e];
          T ko, to;
          if (prog.akind[op] == ACT_ERF)
            erf_act_point<T>(DK[e], ht ? DT[e] : (T)0, qa.x, qa.y, qb.x, qb.y, prog.e_in[op], prog.eA[op],
                             prog.eT[op], prog.eC[op], ko, to);
          else
            act_point(DK[e], ht ? DT[e] : (T)0, qa.x, qa.y, qb.x, qb.y, prog.coef[op], (T)0, prog.hab2[op], ko, to);
          DK[e] = ko;
          if (ht) DT[e] = to;
        }
      } else if (kind == Q_COPY) {
        const int S = cur_S[sidx];
        for (int e = threadIdx.x; e < S * S; e += blockDim.x) {
          DK[e] = imgK[sidx * SS + e];
          if (NTK) DT[e] = imgT[sidx * SS + e];
        }
        if (threadIdx.x == 0) {
          cur_S[d] = S;
          has_t[d] = has_t[sidx];
        }
      } else if (kind == Q_ADD) {
        const int S = cur_S[d];
        const bool hd = NTK && has_t[d], hs = NTK && has_t[sidx];
        for (int e = threadIdx.x; e < S * S; e += blockDim.x) {
          DK[e] = add_rn(DK[e], imgK[sidx * SS + e]);
          if (hd && hs)
            DT[e] = add_rn(DT[e], imgT[sidx * SS + e]);
          else if (hs)
            DT[e] = imgT[sidx * SS + e];
        }
        __syncthreads();
        if (threadIdx.x == 0) has_t[d] = has_t[d] | has_t[sidx];
      }
      __syncthreads();
    }
    // Flatten: mean over the diagonal (linear.py:1880-1882) of buffer 0
    {
      const int S = cur_S[0];
      const bool ht = NTK && has_t[0];
      T sk = (T)0, st = (T)0;
      for (int e = threadIdx.x; e < S * S; e += blockDim.x) {
        sk = add_rn(sk, imgK[e]);
        if (ht) st = add_rn(st, imgT[e]);
      }
      for (int o = 16; o > 0; o >>= 1) {
        sk = add_rn(sk, __shfl_down_sync(0xffffffffu, sk, o));
        st = add_rn(st, __shfl_down_sync(0xffffffffu, st, o));
      }
      if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = sk;
        red[1][threadIdx.x >> 5] = st;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        T a_ = (T)0, b_ = (T)0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
          a_ = add_rn(a_, red[0][w]);
          b_ = add_rn(b_, red[1][w]);
        }
        const T inv = (T)1 / (T)(S * S);
        outK[p] = a_ * inv;
        if (NTK) outT[p] = b_ * inv;
      }
      __syncthreads();
    }
  }
}

// Generic q-program compiler for programs made of Conv3x3(/1|/2, SAME), ABRelu, FanInSum and
// Identity ending in Flatten + Dense*.  Slots are mapped onto three image buffers.
struct DiagPlan {
  bool ok = false;
  std::vector<QOp> ops;
  int n_act = 0;
  std::vector<ntk_op_t> dense_tail;
  double w_first = 0;  // alpha of the first conv is folded into the input scale
};

inline DiagPlan plan_diag(const std::vector<ntk_op_t>& ops, const std::vector<int>& last_use, int out_slot) {
  DiagPlan plan;
  const int n = (int)ops.size();
  std::vector<int> buf_of(last_use.size(), -1);
  bool busy[3] = {false, false, false};
  auto grab = [&]() {
    for (int b = 0; b < 3; ++b)
      if (!busy[b]) {
        busy[b] = true;
        return b;
      }
    return -1;
  };
  buf_of[0] = grab();
  plan.ops.push_back(QOp{Q_INPUT, buf_of[0], buf_of[0], 1, 0, 1.0, 0.0, 0.0});
  int k = 0;
  for (; k < n; ++k) {
    const ntk_op_t& o = ops[k];
    if (o.kind == NTK_OP_FLATTEN) break;
    if (o.src < 0 || buf_of[o.src] < 0) return DiagPlan();
    const bool src_dies = last_use[o.src] == k;
    auto dst_buf = [&](int src_slot, bool dies) {
      if (dies) return buf_of[src_slot];  // in place
      const int b = grab();
      if (b >= 0) plan.ops.push_back(QOp{Q_COPY, b, buf_of[src_slot], 1, 0, 1.0, 0.0, 0.0});
      return b;
    };
    if (o.kind == NTK_OP_CONV) {
      const bool same = o.i[4] == NTK_PAD_SAME && (o.i[2] == 1 || o.i[2] == 2);
      const bool valid = o.i[4] == NTK_PAD_VALID && o.i[2] == 1;  // round 2: VALID 3x3 / 1 on the diagonal column
      const bool circ = o.i[4] == NTK_PAD_CIRCULAR && o.i[2] == 1;  // and CIRCULAR 3x3 / 1 (taps mod S)
      if (!(o.i[0] == 3 && o.i[1] == 3 && o.i[2] == o.i[3] && (same || valid || circ))) return DiagPlan();
      const int b = dst_buf(o.src, src_dies);
      if (b < 0) return DiagPlan();
      plan.ops.push_back(
          QOp{Q_CONV, b, b, valid ? kQValid : (circ ? kQCirc : o.i[2]), 0, o.f[0] / 9.0, o.i[5] ? o.f[1] : 0.0, 0.0});
      buf_of[o.dst] = b;
    } else if (o.kind == NTK_OP_ABRELU || o.kind == NTK_OP_ERF) {
      if (o.kind == NTK_OP_ABRELU && o.i[0]) return DiagPlan();
      const int b = dst_buf(o.src, src_dies);
      if (b < 0) return DiagPlan();
      QOp q{Q_ACT, b, b, 1, plan.n_act++, o.f[0], o.f[1], 0.0};  // alpha/bias carry (a, b) of the activation
      if (o.kind == NTK_OP_ERF) {
        q.akind = ACT_ERF;
        q.erf_c = o.f[2];
      }
      plan.ops.push_back(q);
      buf_of[o.dst] = b;
    } else if (o.kind == NTK_OP_FANINSUM) {
      if (buf_of[o.src2] < 0) return DiagPlan();
      const int b = dst_buf(o.src, src_dies);
      if (b < 0) return DiagPlan();
      plan.ops.push_back(QOp{Q_ADD, b, buf_of[o.src2], 1, 0, 1.0, 0.0, 0.0});
      if (last_use[o.src2] == k && buf_of[o.src2] != b) busy[buf_of[o.src2]] = false;
      buf_of[o.dst] = b;
    } else if (o.kind == NTK_OP_IDENTITY) {
      buf_of[o.dst] = buf_of[o.src];
      continue;
    } else {
      return DiagPlan();
    }
    if (src_dies && buf_of[o.src] != buf_of[o.dst]) busy[buf_of[o.src]] = false;
  }
  if (k >= n || ops[k].kind != NTK_OP_FLATTEN || buf_of[ops[k].src] < 0) return DiagPlan();
  if (buf_of[ops[k].src] != 0)
    plan.ops.push_back(QOp{Q_COPY, 0, buf_of[ops[k].src], 1, 0, 1.0, 0.0, 0.0});
  int z = ops[k].dst;
  for (++k; k < n; ++k) {
    if (ops[k].kind != NTK_OP_DENSE || ops[k].src != z) return DiagPlan();
    z = ops[k].dst;
    plan.dense_tail.push_back(ops[k]);
  }
  if (z != out_slot || plan.ops.size() > (size_t)kMaxQOps || plan.n_act == 0) return DiagPlan();
  plan.ok = true;
  return plan;
}

// The diagonal-column kernels keep 8 (k_diagnet) / 4 (k_qprog) S0 x S0 images in shared memory.
template <typename T>
bool diag_supported(const DiagPlan& plan, int H, int W) {
  return plan.ok && H > 0 && H == W && (size_t)8 * H * H * sizeof(T) <= (size_t)200 * 1024;
}

template <typename T>
int diag_gram(const DiagPlan& plan, Arena& arena, cudaStream_t stream, int64_t* launches, const T* x1,
              int n1, const T* x2, int n2, bool symmetric, int S0, int C, bool want_ntk, T* out_nngp,
              T* out_ntk, long long ld, bool full_square) {
  const bool triangular = symmetric && !full_square && n1 == n2 && n1 > 1;
  QProg<T> qp{};
  std::vector<long long> act_off;
  {
    // track the resolution to size the q-maps
    int Sb[3] = {S0, S0, S0};
    long long off = 0;
    for (const QOp& o : plan.ops) {
      const int i = qp.n++;
      qp.kind[i] = o.kind;
      qp.dst[i] = (signed char)o.dst;
      qp.src[i] = (signed char)o.src;
      qp.stride[i] = (signed char)o.stride;
      qp.act_id[i] = (short)o.act_id;
      qp.alpha[i] = (T)o.alpha;
      qp.bias[i] = (T)o.bias;
      qp.kd0[i] = qp.coef[i] = qp.hab2[i] = (T)0;
      qp.akind[i] = ACT_ABRELU;
      qp.e_in[i] = qp.eA[i] = qp.eT[i] = qp.eC[i] = (T)0;
      if (o.kind == Q_CONV) {
        Sb[o.dst] = qconv_out_size(Sb[o.src], o.stride);  // SAME: ceil; VALID: S - 2
        if (Sb[o.dst] < 1) return fail(NTK_EINVAL, "Conv output would be empty");
      } else if (o.kind == Q_COPY) {
        Sb[o.dst] = Sb[o.src];
      } else if (o.kind == Q_ACT) {
        const FLayer<T> f = res_act_consts<T>(o.alpha, o.bias, 1.0, 0.0, o.akind, o.erf_c);
        qp.akind[i] = (signed char)f.kind;
        qp.e_in[i] = f.e_in;
        qp.eA[i] = f.eA;
        qp.eT[i] = f.eT;
        qp.eC[i] = f.eC;
        qp.coef[i] = f.coef;
        qp.hab2[i] = f.hab2;
        qp.kd0[i] = host_kd0(f.coef, f.hab2);
        qp.alpha[i] = qp.bias[i] = (T)0;
        act_off.push_back(off);
        off += (long long)Sb[o.dst] * Sb[o.dst];
      }
    }
    act_off.push_back(off);
  }
  const long long qm_stride = act_off.back();
  QProg<T>* qp_d = (QProg<T>*)arena.alloc(sizeof(QProg<T>));
  long long* off_d = (long long*)arena.alloc(act_off.size() * sizeof(long long));
  T* qm1 = (T*)arena.alloc((size_t)n1 * qm_stride * 2 * sizeof(T));
  T* qm2 = symmetric ? qm1 : (T*)arena.alloc((size_t)n2 * qm_stride * 2 * sizeof(T));
  const long long Pmax = (long long)n1 * n2;
  T* resK = (T*)arena.alloc((size_t)Pmax * sizeof(T));
  T* resT = want_ntk ? (T*)arena.alloc((size_t)Pmax * sizeof(T)) : nullptr;
  if (!qp_d || !off_d || !qm1 || !qm2 || !resK || (want_ntk && !resT))
    return fail(NTK_ENOMEM, "workspace too small for the diagonal path");
  NTK_CUDA(cudaMemcpyAsync(qp_d, &qp, sizeof(qp), cudaMemcpyHostToDevice, stream));
  NTK_CUDA(cudaMemcpyAsync(off_d, act_off.data(), act_off.size() * sizeof(long long), cudaMemcpyHostToDevice, stream));
  NTK_CUDA(cudaStreamSynchronize(stream));
  const T in_scale = (T)(1.0 / (double)C);
  NTK_TRY(ensure_dynamic_smem((const void*)k_qprog<T>, (size_t)4 * S0 * S0 * sizeof(T)));
  for (int set = 0; set < (symmetric ? 1 : 2); ++set) {
    (*launches)++;
    k_qprog<T><<<set == 0 ? n1 : n2, 256, (size_t)4 * S0 * S0 * sizeof(T), stream>>>(
        set == 0 ? x1 : x2, S0, C, in_scale, qp_d, qm_stride, off_d, set == 0 ? qm1 : qm2);
    NTK_CUDA(cudaGetLastError());
  }
  const long long P = triangular ? tri_prefix(n1, n2) : Pmax;
  const size_t smem = (size_t)8 * S0 * S0 * sizeof(T);
  const int grid = (int)std::min<long long>(P, (long long)kNumSMs * 32);
  (*launches)++;
  if (want_ntk) {
    NTK_TRY(ensure_dynamic_smem((const void*)k_diagnet<T, true>, smem));
    k_diagnet<T, true><<<grid, 128, smem, stream>>>(x1, x2, S0, C, in_scale, qp_d, qm_stride, off_d, qm1, qm2,
                                                   P, n2, triangular ? 1 : 0, resK, resT);
  } else {
    NTK_TRY(ensure_dynamic_smem((const void*)k_diagnet<T, false>, smem));
    k_diagnet<T, false><<<grid, 128, smem, stream>>>(x1, x2, S0, C, in_scale, qp_d, qm_stride, off_d, qm1, qm2,
                                                    P, n2, triangular ? 1 : 0, resK, resT);
  }
  NTK_CUDA(cudaGetLastError());
  for (const ntk_op_t& d : plan.dense_tail) {
    (*launches)++;
    k_dense<T><<<grid_for(P), kThreads, 0, stream>>>(resK, resT, P, (T)d.f[0], (T)(d.i[0] ? d.f[1] : 0.0), 0);
    NTK_CUDA(cudaGetLastError());
  }
  (*launches)++;
  if (triangular)
    k_scatter_tri<T><<<grid_for(P), kThreads, 0, stream>>>(resK, out_nngp, P, n2, ld, 0);
  else
    k_scatter<T><<<grid_for(P), kThreads, 0, stream>>>(resK, out_nngp, n1, n2, 1LL, ld, 0, 0);
  NTK_CUDA(cudaGetLastError());
  if (want_ntk) {
    (*launches)++;
    if (triangular)
      k_scatter_tri<T><<<grid_for(P), kThreads, 0, stream>>>(resT, out_ntk, P, n2, ld, 0);
    else
      k_scatter<T><<<grid_for(P), kThreads, 0, stream>>>(resT, out_ntk, n1, n2, 1LL, ld, 0, 0);
    NTK_CUDA(cudaGetLastError());
  }
  if (triangular) {
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

}  // namespace ntk
