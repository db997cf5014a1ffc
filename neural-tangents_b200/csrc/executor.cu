// Program interpreter + C-ABI entry points of libntk_b200.so.
//
// A program is the slot form of a `stax` layer tree (include/ntk_b200.h).  The
// executor tiles the n1 x n2 pair grid (the job of nt.batch's serial loop,
// _src/batching.py:314-502), builds the input kernel of each tile
// (_src/stax/requirements.py:641-830), runs the layer rules on the GPU and
// scatters the tile into the result matrices.  There is no CPU path.
#include <algorithm>
#include <memory>
#include <type_traits>

#include "common.cuh"
#include "generic_kernels.cuh"
#include "instantiate.cuh"

namespace ntk {
NTK_FUSED_INSTANCES(extern, float)
NTK_FUSED_INSTANCES(extern, double)
NTK_FUSED_ERF_INSTANCES(extern, float)
NTK_FUSED_ERF_INSTANCES(extern, double)
NTK_FUSED_EMB_INSTANCES(extern, float)
NTK_FUSED_EMB_INSTANCES(extern, double)
NTK_FUSED_EMB_GEN_INSTANCES(extern, float)
NTK_FUSED_EMB_GEN_INSTANCES(extern, double)
NTK_FUSED_GEN_INSTANCES(extern, float)
NTK_FUSED_GEN_INSTANCES(extern, double)
NTK_RES_ERF_INSTANCES(extern, float)
NTK_RES_ERF_INSTANCES(extern, double)
NTK_RES_INSTANCES(extern, float)
NTK_RES_INSTANCES(extern, double)
}  // namespace ntk

using namespace ntk;

// ------------------------------------------------------------------------------
struct ntk_program {
  std::vector<ntk_op_t> ops;
  int n_slots = 0;
  int out_slot = 0;
  std::vector<int> last_use;  // per slot: index of the last op reading it (or n_ops if output)
  FcnProg fcn{};              // Dense/ABRelu/Erf chain on [N, d] inputs (k_fcn_chain); n == 0 if not matched
  FusedPlan fused;            // fast-path plan (fused_kernels.cuh); empty if not matched
  FusedPlan per_layer;        // same kernels, one layer per launch (NTK_FLAG_PER_LAYER)
  ResPlan res;                // residual-network plan (res_kernels.cuh)
  DiagPlan diag;              // pool-free nets ending in Flatten: diagonal column only
};

struct ntk_context {
  int device = 0;
  cudaStream_t stream = nullptr;      // stream of the current call: own_stream unless a *_on_stream entry is running
  cudaStream_t own_stream = nullptr;  // created (and destroyed) by the context
  // Workspace reuse across calls on different streams is ordered on the device: every call records
  // `order_ev` on the stream it used; a call on another stream makes that stream wait for it first.
  cudaEvent_t order_ev = nullptr;
  cudaStream_t last_stream = nullptr;
  bool order_valid = false;
  // pinned staging ring of ntk_gram_host (pageable caller arrays): 2 slots of kPinSlot bytes
  char* pin = nullptr;
  cudaEvent_t pin_ev[2] = {nullptr, nullptr};
  bool pin_busy[2] = {false, false};
  char* ws = nullptr;
  size_t ws_bytes = 0;
  int64_t launches = 0;
  Arena arena;
  // grow-only device IO buffers for the *_host entry points
  void* io[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t io_bytes[6] = {0, 0, 0, 0, 0, 0};
  StageProfile prof;
};

namespace {

struct Buf {
  void* p = nullptr;
  int refs = 0;
};

struct TState {
  Buf* nngp = nullptr;
  Buf* ntk = nullptr;
  Buf* cov1 = nullptr;
  Buf* cov2 = nullptr;
  int H = 0, W = 0;
  int ntk_mode = NTK_NTK_NONE;
  bool gaussian = false;
  bool valid = false;
};

// Execution environment of one tile (or one dry run).
struct Env {
  ntk_context* ctx = nullptr;  // nullptr in dry runs
  Arena* arena = nullptr;
  cudaStream_t stream = nullptr;
  bool dry = false;
  int64_t launches = 0;
  std::vector<std::unique_ptr<Buf>> bufs;

  Buf* alloc(size_t bytes, int* status) {
    void* p = arena->alloc(bytes);
    if (!p) {
      *status = fail(NTK_ENOMEM, "workspace exhausted allocating %zu bytes", bytes);
      return nullptr;
    }
    bufs.emplace_back(new Buf());
    bufs.back()->p = p;
    bufs.back()->refs = 1;
    return bufs.back().get();
  }
  void unref(Buf*& b) {
    if (b && --b->refs == 0) {
      arena->release(b->p);
      b->p = nullptr;
    }
    b = nullptr;
  }
  void release(TState& s) {
    unref(s.nngp);
    unref(s.ntk);
    unref(s.cov1);
    unref(s.cov2);
    s.valid = false;
  }
};

#define LAUNCH(env, kernel, grid, block, smem, ...)                                   \
  do {                                                                                \
    (env).launches++;                                                                 \
    if (!(env).dry) {                                                                 \
      kernel<<<(grid), (block), (smem), (env).stream>>>(__VA_ARGS__);                 \
      NTK_CUDA(cudaGetLastError());                                                   \
    }                                                                                 \
  } while (0)

inline long long per_of(int H, int W) { return H > 0 ? (long long)H * H * W * W : 1; }

template <typename T>
int copy_buf(Env& env, Buf* src, size_t bytes, Buf** out) {
  int st = NTK_OK;
  Buf* b = env.alloc(bytes, &st);
  if (!b) return st;
  if (!env.dry) NTK_CUDA(cudaMemcpyAsync(b->p, src->p, bytes, cudaMemcpyDeviceToDevice, env.stream));
  env.launches++;
  *out = b;
  return NTK_OK;
}

// Make `b` exclusively owned (copy-on-write) so an in-place kernel may run on it.
template <typename T>
int make_unique_buf(Env& env, Buf*& b, size_t bytes) {
  if (!b || b->refs == 1) return NTK_OK;
  Buf* nb = nullptr;
  NTK_TRY(copy_buf<T>(env, b, bytes, &nb));
  env.unref(b);
  b = nb;
  return NTK_OK;
}

// ---- layer rules on a tile --------------------------------------------------------
template <typename T>
int op_conv(Env& env, const ntk_op_t& op, const TState& in, TState& out, int t1, int t2) {
  if (in.H <= 0) return fail(NTK_EINVAL, "Conv needs spatial inputs");
  const int kh = op.i[0], kw = op.i[1], sh = op.i[2], sw = op.i[3], pad = op.i[4];
  AxisGeom gh = axis_geom(in.H, kh, sh, pad), gw = axis_geom(in.W, kw, sw, pad);
  if (gh.out <= 0 || gw.out <= 0) return fail(NTK_EINVAL, "Conv output would be empty");
  ConvGeom g{in.H, in.W, gh.out, gw.out, kh, kw, sh, sw, gh.lo, gw.lo, pad == NTK_PAD_CIRCULAR};
  const T scale = (T)(op.f[0] / (double)(kh * kw));
  const T shift = (T)(op.i[5] ? op.f[1] : 0.0);
  const long long per_o = per_of(g.Ho, g.Wo);
  int st = NTK_OK;
  auto run = [&](Buf* src, Buf* addend, long long P, T sc, T shf, Buf** dst) -> int {
    Buf* o = env.alloc((size_t)(P * per_o) * sizeof(T), &st);
    if (!o) return st;
    LAUNCH(env, k_conv<T>, grid_for(P * per_o), kThreads, 0, (const T*)src->p,
           addend ? (const T*)addend->p : (const T*)nullptr, (T*)o->p, P, g, sc, shf);
    *dst = o;
    return NTK_OK;
  };
  const long long P = (long long)t1 * t2;
  const bool k3 = kh == 3 && kw == 3 && pad != NTK_PAD_CIRCULAR;
  auto run3 = [&](Buf* srcK, Buf* srcT, long long np, Buf** dstK, Buf** dstT) -> int {
    Buf* ok = env.alloc((size_t)(np * per_o) * sizeof(T), &st);
    if (!ok) return st;
    Buf* ot = nullptr;
    if (srcT) {
      ot = env.alloc((size_t)(np * per_o) * sizeof(T), &st);
      if (!ot) return st;
    }
    LAUNCH(env, k_conv3<T>, grid_for(np * per_o), kThreads, 0, (const T*)srcK->p, srcT ? (const T*)srcT->p : (const T*)nullptr,
           (T*)ok->p, ot ? (T*)ot->p : (T*)nullptr, np, g, scale, shift);
    *dstK = ok;
    if (dstT) *dstT = ot;
    return NTK_OK;
  };
  out.ntk_mode = in.ntk_mode;
  if (k3) {
    NTK_TRY(run3(in.nngp, in.ntk_mode == NTK_NTK_TENSOR ? in.ntk : nullptr, P, &out.nngp, &out.ntk));
  } else {
    NTK_TRY(run(in.nngp, nullptr, P, scale, shift, &out.nngp));
    if (in.ntk_mode == NTK_NTK_TENSOR) NTK_TRY(run(in.ntk, out.nngp, P, scale, (T)0, &out.ntk));  // linear.py:1396-1398
  }
  if (in.ntk_mode == NTK_NTK_ZERO) {
    NTK_TRY(copy_buf<T>(env, out.nngp, (size_t)(P * per_o) * sizeof(T), &out.ntk));
    out.ntk_mode = NTK_NTK_TENSOR;
  }
  if (k3) {
    NTK_TRY(run3(in.cov1, nullptr, t1, &out.cov1, nullptr));
    NTK_TRY(run3(in.cov2, nullptr, t2, &out.cov2, nullptr));
    out.H = g.Ho;
    out.W = g.Wo;
    out.gaussian = true;
    out.valid = true;
    return NTK_OK;
  }
  NTK_TRY(run(in.cov1, nullptr, t1, scale, shift, &out.cov1));
  NTK_TRY(run(in.cov2, nullptr, t2, scale, shift, &out.cov2));
  out.H = g.Ho;
  out.W = g.Wo;
  out.gaussian = true;
  out.valid = true;
  return NTK_OK;
}

template <typename T>
int op_pool(Env& env, const ntk_op_t& op, const TState& in, TState& out, int t1, int t2) {
  if (in.H <= 0) return fail(NTK_EINVAL, "AvgPool needs spatial inputs");
  const int wh = op.i[0], ww = op.i[1], sh = op.i[2], sw = op.i[3], pad = op.i[4];
  AxisGeom gh = axis_geom(in.H, wh, sh, pad), gw = axis_geom(in.W, ww, sw, pad);
  if (gh.out <= 0 || gw.out <= 0) return fail(NTK_EINVAL, "AvgPool output would be empty");
  PoolGeom g{in.H, in.W, gh.out, gw.out, wh, ww, sh, sw, gh.lo, gw.lo, pad == NTK_PAD_CIRCULAR,
             ((op.i[5] & 1) && pad == NTK_PAD_SAME) ? 1 : 0, (op.i[5] & 2) ? 1 : 0};
  const long long per_o = per_of(g.Ho, g.Wo);
  int st = NTK_OK;
  auto run = [&](Buf* src, long long P, Buf** dst) -> int {
    Buf* o = env.alloc((size_t)(P * per_o) * sizeof(T), &st);
    if (!o) return st;
    LAUNCH(env, k_pool<T>, grid_for(P * per_o), kThreads, 0, (const T*)src->p, (T*)o->p, P, g);
    *dst = o;
    return NTK_OK;
  };
  const long long P = (long long)t1 * t2;
  NTK_TRY(run(in.nngp, P, &out.nngp));
  out.ntk_mode = in.ntk_mode;
  if (in.ntk_mode == NTK_NTK_TENSOR) NTK_TRY(run(in.ntk, P, &out.ntk));
  NTK_TRY(run(in.cov1, t1, &out.cov1));
  NTK_TRY(run(in.cov2, t2, &out.cov2));
  out.H = g.Ho;
  out.W = g.Wo;
  out.gaussian = in.gaussian;
  out.valid = true;
  return NTK_OK;
}

template <typename T>
int op_reduce(Env& env, const ntk_op_t& op, const TState& in, TState& out, int t1, int t2) {
  const bool flatten = op.kind == NTK_OP_FLATTEN;
  if (in.H <= 0) {
    if (!flatten) return fail(NTK_EINVAL, "GlobalAvgPool needs spatial inputs");
    out = in;  // Flatten of [N,d] is the identity (shares buffers)
    for (Buf* b : {out.nngp, out.ntk, out.cov1, out.cov2})
      if (b) b->refs++;
    out.gaussian = false;
    return NTK_OK;
  }
  int st = NTK_OK;
  auto run = [&](Buf* src, long long P, Buf** dst) -> int {
    Buf* o = env.alloc((size_t)P * sizeof(T), &st);
    if (!o) return st;
    int grid = (int)std::min<long long>(P, (long long)kNumSMs * 16);
    if (flatten)
      LAUNCH(env, (k_reduce_spatial<T, true>), grid, kThreads, 0, (const T*)src->p, (T*)o->p, P,
             in.H, in.W);
    else
      LAUNCH(env, (k_reduce_spatial<T, false>), grid, kThreads, 0, (const T*)src->p, (T*)o->p, P,
             in.H, in.W, op.i[0] == 1 ? 1 : 0);
    *dst = o;
    return NTK_OK;
  };
  const long long P = (long long)t1 * t2;
  NTK_TRY(run(in.nngp, P, &out.nngp));
  out.ntk_mode = in.ntk_mode;
  if (in.ntk_mode == NTK_NTK_TENSOR) NTK_TRY(run(in.ntk, P, &out.ntk));
  NTK_TRY(run(in.cov1, t1, &out.cov1));
  NTK_TRY(run(in.cov2, t2, &out.cov2));
  out.H = out.W = 0;
  out.gaussian = flatten ? false : in.gaussian;
  out.valid = true;
  return NTK_OK;
}

// Moves (steal == true) or shares the buffers of `in` into `out`, then makes them
// exclusively owned so in-place kernels can run.
template <typename T>
int take_for_inplace(Env& env, TState& in, TState& out, bool steal, int t1, int t2) {
  out = in;
  if (steal) {
    in.nngp = in.ntk = in.cov1 = in.cov2 = nullptr;
    in.valid = false;
  } else {
    for (Buf* b : {out.nngp, out.ntk, out.cov1, out.cov2})
      if (b) b->refs++;
  }
  const long long per = per_of(out.H, out.W);
  const long long P = (long long)t1 * t2;
  NTK_TRY(make_unique_buf<T>(env, out.nngp, (size_t)(P * per) * sizeof(T)));
  NTK_TRY(make_unique_buf<T>(env, out.ntk, (size_t)(P * per) * sizeof(T)));
  NTK_TRY(make_unique_buf<T>(env, out.cov1, (size_t)(t1 * per) * sizeof(T)));
  NTK_TRY(make_unique_buf<T>(env, out.cov2, (size_t)(t2 * per) * sizeof(T)));
  return NTK_OK;
}

template <typename T>
int op_act(Env& env, const ntk_op_t& op, TState& in, TState& out, bool steal, int t1, int t2) {
  if (!in.gaussian && op.kind == NTK_OP_LAYERNORM)
    return fail(NTK_EUNSUPPORTED, "LayerNorm only implemented for Gaussian inputs.");   // linear.py:2519-2521
  if (!in.gaussian)
    return fail(NTK_ENOTGAUSSIAN,
                "The input to the activation function must be Gaussian, i.e. a random affine "
                "transform is required before the activation function.");
  NTK_TRY((take_for_inplace<T>(env, in, out, steal, t1, t2)));
  const int H = out.H > 0 ? out.H : 1, W = out.W > 0 ? out.W : 1;
  const long long per = per_of(out.H, out.W);
  const long long P = (long long)t1 * t2;
  int st = NTK_OK;
  Buf* q1 = env.alloc((size_t)t1 * H * W * sizeof(T), &st);
  if (!q1) return st;
  Buf* q2 = env.alloc((size_t)t2 * H * W * sizeof(T), &st);
  if (!q2) return st;
  LAUNCH(env, k_diag<T>, grid_for((long long)t1 * H * W), kThreads, 0, (const T*)out.cov1->p,
         (T*)q1->p, (long long)t1, H, W);
  LAUNCH(env, k_diag<T>, grid_for((long long)t2 * H * W), kThreads, 0, (const T*)out.cov2->p,
         (T*)q2->p, (long long)t2, H, W);
  ActParams ap{op.kind, op.f[0], op.f[1], op.f[2]};
  Buf* stab = nullptr;
  if (op.kind == NTK_OP_ABRELU && op.i[0]) {  // do_stabilize: elementwise.py:430-436
    stab = env.alloc(256, &st);
    if (!stab) return st;
    if (!env.dry) NTK_CUDA(cudaMemsetAsync(stab->p, 0, 256, env.stream));
    LAUNCH(env, k_absmax<T>, grid_for(P * per), kThreads, 0, (const T*)out.nngp->p, P * per,
           (T*)stab->p);
  }
  const T* sp = stab ? (const T*)stab->p : (const T*)nullptr;
  T* tt = out.ntk_mode == NTK_NTK_TENSOR ? (T*)out.ntk->p : (T*)nullptr;
  auto act = [&](T* kk, T* ttp, const T* qa, const T* qb, long long np, PairMap pm) -> int {
    if (W % 4 == 0)  // 4 consecutive w' per lane
      LAUNCH(env, (k_act<T, 4>), grid_for(np * per / 4), kThreads, 0, kk, ttp, qa, qb, np, pm, H, W, ap, sp);
    else
      LAUNCH(env, (k_act<T, 1>), grid_for(np * per), kThreads, 0, kk, ttp, qa, qb, np, pm, H, W, ap, sp);
    return NTK_OK;
  };
  NTK_TRY(act((T*)out.nngp->p, tt, (const T*)q1->p, (const T*)q2->p, P, PairMap{t2, 0}));
  NTK_TRY(act((T*)out.cov1->p, (T*)nullptr, (const T*)q1->p, (const T*)q1->p, (long long)t1, PairMap{1, 1}));
  NTK_TRY(act((T*)out.cov2->p, (T*)nullptr, (const T*)q2->p, (const T*)q2->p, (long long)t2, PairMap{1, 1}));
  env.unref(q1);
  env.unref(q2);
  if (stab) env.unref(stab);
  out.gaussian = op.kind == NTK_OP_LAYERNORM;  // LayerNorm keeps a Gaussian input Gaussian; activations do not
  out.valid = true;
  return NTK_OK;
}

template <typename T>
int op_dense(Env& env, const ntk_op_t& op, TState& in, TState& out, bool steal, int t1, int t2) {
  NTK_TRY((take_for_inplace<T>(env, in, out, steal, t1, t2)));
  const long long per = per_of(out.H, out.W);
  const long long P = (long long)t1 * t2;
  const T w2 = (T)op.f[0], b2 = (T)(op.i[0] ? op.f[1] : 0.0);
  int st = NTK_OK;
  int t_zero = 0;
  if (out.ntk_mode == NTK_NTK_ZERO) {
    out.ntk = env.alloc((size_t)(P * per) * sizeof(T), &st);
    if (!out.ntk) return st;
    out.ntk_mode = NTK_NTK_TENSOR;
    t_zero = 1;
  }
  T* tt = out.ntk_mode == NTK_NTK_TENSOR ? (T*)out.ntk->p : (T*)nullptr;
  LAUNCH(env, k_dense<T>, grid_for(P * per), kThreads, 0, (T*)out.nngp->p, tt, P * per, w2, b2,
         t_zero);
  LAUNCH(env, k_dense<T>, grid_for(t1 * per), kThreads, 0, (T*)out.cov1->p, (T*)nullptr, t1 * per,
         w2, b2, 0);
  LAUNCH(env, k_dense<T>, grid_for(t2 * per), kThreads, 0, (T*)out.cov2->p, (T*)nullptr, t2 * per,
         w2, b2, 0);
  out.gaussian = true;
  out.valid = true;
  return NTK_OK;
}

template <typename T>
int op_faninsum(Env& env, const TState& a, const TState& b, TState& out, int t1, int t2) {
  if (a.H != b.H || a.W != b.W)
    return fail(NTK_ESHAPE, "All shapes should be equal in `FanInSum`, got %dx%d and %dx%d", a.H,
                a.W, b.H, b.W);
  if (!(a.gaussian && b.gaussian))
    return fail(NTK_EUNSUPPORTED,
                "`FanInSum` is only implemented for the case where all input layers are "
                "guaranteed to be mean-zero Gaussian (is_gaussian == True).");
  const long long per = per_of(a.H, a.W);
  const long long P = (long long)t1 * t2;
  int st = NTK_OK;
  auto add = [&](Buf* x, Buf* y, long long n, Buf** dst) -> int {
    Buf* o = env.alloc((size_t)n * sizeof(T), &st);
    if (!o) return st;
    LAUNCH(env, k_add<T>, grid_for(n), kThreads, 0, (const T*)x->p, (const T*)y->p, (T*)o->p, n);
    *dst = o;
    return NTK_OK;
  };
  NTK_TRY(add(a.nngp, b.nngp, P * per, &out.nngp));
  NTK_TRY(add(a.cov1, b.cov1, t1 * per, &out.cov1));
  NTK_TRY(add(a.cov2, b.cov2, t2 * per, &out.cov2));
  if (a.ntk_mode == NTK_NTK_TENSOR && b.ntk_mode == NTK_NTK_TENSOR) {
    NTK_TRY(add(a.ntk, b.ntk, P * per, &out.ntk));
    out.ntk_mode = NTK_NTK_TENSOR;
  } else if (a.ntk_mode == NTK_NTK_TENSOR || b.ntk_mode == NTK_NTK_TENSOR) {
    out.ntk = a.ntk_mode == NTK_NTK_TENSOR ? a.ntk : b.ntk;
    out.ntk->refs++;
    out.ntk_mode = NTK_NTK_TENSOR;
  } else {
    out.ntk_mode = a.ntk_mode;
  }
  out.H = a.H;
  out.W = a.W;
  out.gaussian = true;
  out.valid = true;
  return NTK_OK;
}

// Runs ops [first, n_ops) of the program on slots; slot `ops[first].src` etc. must be valid.
template <typename T>
int run_ops(Env& env, const ntk_program& prog, std::vector<TState>& slots, int first, int t1,
            int t2) {
  const int n_ops = (int)prog.ops.size();
  for (int k = first; k < n_ops; ++k) {
    const ntk_op_t& op = prog.ops[k];
    TState& in = slots[op.src];
    if (!in.valid) return fail(NTK_EINVAL, "op %d reads empty slot %d", k, op.src);
    TState out;
    const bool steal = prog.last_use[op.src] == k;
    switch (op.kind) {
      case NTK_OP_CONV:
        NTK_TRY(op_conv<T>(env, op, in, out, t1, t2));
        break;
      case NTK_OP_AVGPOOL:
        NTK_TRY(op_pool<T>(env, op, in, out, t1, t2));
        break;
      case NTK_OP_GAP:
      case NTK_OP_FLATTEN:
        NTK_TRY(op_reduce<T>(env, op, in, out, t1, t2));
        out.valid = true;
        break;
      case NTK_OP_ABRELU:
      case NTK_OP_ERF:
      case NTK_OP_GELU:
      case NTK_OP_SIN:
      case NTK_OP_RBF:
      case NTK_OP_LAYERNORM:
        NTK_TRY(op_act<T>(env, op, in, out, steal, t1, t2));
        break;
      case NTK_OP_DENSE:
        NTK_TRY(op_dense<T>(env, op, in, out, steal, t1, t2));
        break;
      case NTK_OP_FANINSUM: {
        TState& in2 = slots[op.src2];
        if (!in2.valid) return fail(NTK_EINVAL, "op %d reads empty slot %d", k, op.src2);
        NTK_TRY(op_faninsum<T>(env, in, in2, out, t1, t2));
        break;
      }
      case NTK_OP_IDENTITY:
        out = in;
        for (Buf* b : {out.nngp, out.ntk, out.cov1, out.cov2})
          if (b) b->refs++;
        break;
      default:
        return fail(NTK_EINVAL, "unknown op kind %d", op.kind);
    }
    // retire dead inputs, then publish the result
    if (prog.last_use[op.src] == k && slots[op.src].valid) env.release(slots[op.src]);
    if (op.src2 >= 0 && prog.last_use[op.src2] == k && slots[op.src2].valid)
      env.release(slots[op.src2]);
    if (slots[op.dst].valid) env.release(slots[op.dst]);
    slots[op.dst] = out;
  }
  return NTK_OK;
}

// Builds the input kernel of a tile: requirements.py:641-830.
template <typename T>
int build_input_state(Env& env, const T* x1, int t1, const T* x2, int t2, int H, int W, int C,
                      bool want_ntk, TState& s) {
  const long long P = (long long)t1 * t2;
  int st = NTK_OK;
  const T inv_c = (T)(1.0 / (double)C);
  if (H > 0) {
    const long long per = per_of(H, W);
    s.nngp = env.alloc((size_t)(P * per) * sizeof(T), &st);
    if (!s.nngp) return st;
    s.cov1 = env.alloc((size_t)(t1 * per) * sizeof(T), &st);
    if (!s.cov1) return st;
    s.cov2 = env.alloc((size_t)(t2 * per) * sizeof(T), &st);
    if (!s.cov2) return st;
    LAUNCH(env, k_input_cov<T>, grid_for(P * per), kThreads, 0, x1, x2, (T*)s.nngp->p, P,
           PairMap{t2, 0}, H, W, C, inv_c);
    LAUNCH(env, k_input_cov<T>, grid_for(t1 * per), kThreads, 0, x1, x1, (T*)s.cov1->p,
           (long long)t1, PairMap{1, 1}, H, W, C, inv_c);
    LAUNCH(env, k_input_cov<T>, grid_for(t2 * per), kThreads, 0, x2, x2, (T*)s.cov2->p,
           (long long)t2, PairMap{1, 1}, H, W, C, inv_c);
  } else {
    s.nngp = env.alloc((size_t)P * sizeof(T), &st);
    if (!s.nngp) return st;
    s.cov1 = env.alloc((size_t)t1 * sizeof(T), &st);
    if (!s.cov1) return st;
    s.cov2 = env.alloc((size_t)t2 * sizeof(T), &st);
    if (!s.cov2) return st;
    NTK_TRY((fcn_input_gram<T>(env.dry, env.stream, &env.launches, x1, t1, x2, t2, C,
                               (T*)s.nngp->p)));
    LAUNCH(env, k_rowdot<T>, grid_for((long long)t1 * 32), kThreads, 0, x1, x1, (T*)s.cov1->p,
           (long long)t1, PairMap{1, 1}, C, inv_c);
    LAUNCH(env, k_rowdot<T>, grid_for((long long)t2 * 32), kThreads, 0, x2, x2, (T*)s.cov2->p,
           (long long)t2, PairMap{1, 1}, C, inv_c);
  }
  s.H = H;
  s.W = W;
  s.ntk_mode = want_ntk ? NTK_NTK_ZERO : NTK_NTK_NONE;
  s.gaussian = false;
  s.valid = true;
  return NTK_OK;
}

struct OutPtrs {
  void* nngp;
  void* ntk;
  void* cov1;
  void* cov2;
  long long ld;
};

template <typename T>
int scatter_tile(Env& env, const TState& s, int t1, int t2, int r0, int c0, int n2, const OutPtrs& o) {
  const long long per = per_of(s.H, s.W);
  const long long ld = s.H > 0 ? n2 : o.ld;
  const long long P = (long long)t1 * t2;
  if (o.nngp)
    LAUNCH(env, k_scatter<T>, grid_for(P * per), kThreads, 0, (const T*)s.nngp->p, (T*)o.nngp, t1,
           t2, per, ld, r0, c0);
  if (o.ntk) {
    if (s.ntk_mode == NTK_NTK_TENSOR) {
      LAUNCH(env, k_scatter<T>, grid_for(P * per), kThreads, 0, (const T*)s.ntk->p, (T*)o.ntk, t1,
             t2, per, ld, r0, c0);
    } else {
      // the 0-d zero of requirements.py:807 never met a Dense/Conv: materialise zeros.
      int st = NTK_OK;
      Buf* z = env.alloc((size_t)(P * per) * sizeof(T), &st);
      if (!z) return st;
      LAUNCH(env, k_fill<T>, grid_for(P * per), kThreads, 0, (T*)z->p, P * per, (T)0);
      LAUNCH(env, k_scatter<T>, grid_for(P * per), kThreads, 0, (const T*)z->p, (T*)o.ntk, t1, t2,
             per, ld, r0, c0);
      env.unref(z);
    }
  }
  if (o.cov1 && c0 == 0)
    LAUNCH(env, k_scatter<T>, grid_for(t1 * per), kThreads, 0, (const T*)s.cov1->p, (T*)o.cov1, t1,
           1, per, 1LL, r0, 0);
  if (o.cov2 && r0 == 0)
    LAUNCH(env, k_scatter<T>, grid_for(t2 * per), kThreads, 0, (const T*)s.cov2->p, (T*)o.cov2, t2,
           1, per, 1LL, c0, 0);
  return NTK_OK;
}

// One tile through the general per-layer path.
template <typename T>
int run_tile_generic(Env& env, const ntk_program& prog, const T* x1, int t1, const T* x2, int t2,
                     int H, int W, int C, bool want_ntk, int r0, int c0, int n2,
                     const OutPtrs& out) {
  std::vector<TState> slots(prog.n_slots);
  NTK_TRY(build_input_state<T>(env, x1, t1, x2, t2, H, W, C, want_ntk, slots[0]));
  NTK_TRY(run_ops<T>(env, prog, slots, 0, t1, t2));
  TState& fin = slots[prog.out_slot];
  if (!fin.valid) return fail(NTK_EINVAL, "program left its output slot empty");
  NTK_TRY(scatter_tile<T>(env, fin, t1, t2, r0, c0, n2, out));
  for (auto& s : slots)
    if (s.valid) env.release(s);
  return NTK_OK;
}

template <typename T>
int dry_peak_generic(const ntk_program& prog, int t1, int t2, int H, int W, int C, bool want_ntk,
                     size_t* peak) {
  Arena a;
  a.reset(nullptr, 0, true);
  Env env;
  env.arena = &a;
  env.dry = true;
  OutPtrs o{(void*)1, want_ntk ? (void*)1 : nullptr, nullptr, nullptr, t2};
  int st = run_tile_generic<T>(env, prog, (const T*)nullptr, t1, (const T*)nullptr, t2, H, W, C,
                               want_ntk, 0, 0, t2, o);
  *peak = a.peak();
  return st;
}

template <typename T>
int choose_tile(const ntk_program& prog, size_t ws, int n1, int n2, int H, int W, int C,
                bool want_ntk, int* t1o, int* t2o) {
  auto fits = [&](int a, int b, bool* ok) -> int {
    size_t peak = 0;
    int st = dry_peak_generic<T>(prog, a, b, H, W, C, want_ntk, &peak);
    if (st != NTK_OK) return st;
    *ok = peak <= ws;
    return NTK_OK;
  };
  bool ok = false;
  NTK_TRY(fits(1, 1, &ok));
  if (!ok) return fail(NTK_ENOMEM, "workspace (%zu bytes) cannot hold a single pair", ws);
  int t1 = 1, t2 = 1;
  // grow alternately while it fits
  for (;;) {
    bool grew = false;
    if (t2 < n2) {
      int c = std::min(n2, t2 * 2);
      NTK_TRY(fits(t1, c, &ok));
      if (ok) {
        t2 = c;
        grew = true;
      }
    }
    if (t1 < n1) {
      int c = std::min(n1, t1 * 2);
      NTK_TRY(fits(c, t2, &ok));
      if (ok) {
        t1 = c;
        grew = true;
      }
    }
    if (!grew) break;
    if ((long long)t1 * t2 >= (1LL << 22)) break;
  }
  *t1o = t1;
  *t2o = t2;
  return NTK_OK;
}

int validate_program(ntk_program& p) {
  const int n = (int)p.ops.size();
  if (p.n_slots < 1 || p.out_slot < 0 || p.out_slot >= p.n_slots)
    return fail(NTK_EINVAL, "bad slot configuration");
  p.last_use.assign(p.n_slots, -1);
  for (int k = 0; k < n; ++k) {
    const ntk_op_t& op = p.ops[k];
    if (op.src < 0 || op.src >= p.n_slots || op.dst < 0 || op.dst >= p.n_slots)
      return fail(NTK_EINVAL, "op %d: slot out of range", k);
    if (op.kind == NTK_OP_FANINSUM && (op.src2 < 0 || op.src2 >= p.n_slots))
      return fail(NTK_EINVAL, "op %d: FanInSum needs src2", k);
    p.last_use[op.src] = k;
    if (op.kind == NTK_OP_FANINSUM) p.last_use[op.src2] = k;
    switch (op.kind) {
      case NTK_OP_CONV:
      case NTK_OP_AVGPOOL:
        if (op.i[0] < 1 || op.i[1] < 1 || op.i[2] < 1 || op.i[3] < 1 || op.i[4] < 0 || op.i[4] > 2)
          return fail(NTK_EINVAL, "op %d: bad window/stride/padding", k);
        break;
      case NTK_OP_DENSE:
      case NTK_OP_ABRELU:
      case NTK_OP_ERF:
      case NTK_OP_GELU:
      case NTK_OP_SIN:
      case NTK_OP_RBF:
      case NTK_OP_LAYERNORM:
      case NTK_OP_GAP:
      case NTK_OP_FLATTEN:
      case NTK_OP_FANINSUM:
      case NTK_OP_IDENTITY:
        break;
      default:
        return fail(NTK_EINVAL, "op %d: unknown kind %d", k, op.kind);
    }
  }
  p.last_use[p.out_slot] = n;  // never retired by an op
  return NTK_OK;
}

int ensure_io(ntk_context* ctx, int which, size_t bytes) {
  if (ctx->io_bytes[which] >= bytes) return NTK_OK;
  if (ctx->io[which]) NTK_CUDA(cudaFree(ctx->io[which]));
  ctx->io[which] = nullptr;
  ctx->io_bytes[which] = 0;
  size_t want = std::max(bytes, (size_t)1 << 20);
  NTK_CUDA(cudaMalloc(&ctx->io[which], want));
  ctx->io_bytes[which] = want;
  return NTK_OK;
}

// Recognises a serial chain of Dense / ABRelu (no do_stabilize) / Erf ops from slot 0 to the output.
FcnProg plan_fcn(const std::vector<ntk_op_t>& ops, int out_slot) {
  FcnProg f{};
  const int n = (int)ops.size();
  if (n == 0 || n > kMaxFcnOps || ops[0].kind != NTK_OP_DENSE) return FcnProg{};
  bool gaussian = false;
  for (int k = 0; k < n; ++k) {
    const ntk_op_t& o = ops[k];
    if (o.src != (k == 0 ? 0 : ops[k - 1].dst)) return FcnProg{};
    if (o.kind == NTK_OP_DENSE) {
      gaussian = true;
    } else if ((o.kind == NTK_OP_ABRELU && o.i[0] == 0) || o.kind == NTK_OP_ERF) {
      if (!gaussian) return FcnProg{};  // the per-op path reports the reference's error
      gaussian = false;
    } else {
      return FcnProg{};
    }
    f.kind[k] = o.kind;
    f.has_bias[k] = o.kind == NTK_OP_DENSE ? o.i[0] : 0;
    f.f0[k] = o.f[0];
    f.f1[k] = o.f[1];
    f.f2[k] = o.f[2];
  }
  if (ops[n - 1].dst != out_slot) return FcnProg{};
  f.n = n;
  return f;
}

// [N, d] inputs + a Dense/ABRelu/Erf chain: input Gram on the tensor cores, then ONE elementwise launch.
template <typename T>
int fcn_gram(ntk_context* ctx, Env& env, const FcnProg& fp, const T* x1, int n1, const T* x2, int n2,
             bool symmetric, int C, bool want_ntk, T* out_nngp, T* out_ntk, long long ld) {
  Arena& arena = ctx->arena;
  T* c1 = (T*)arena.alloc((size_t)n1 * sizeof(T));
  T* c2 = symmetric ? c1 : (T*)arena.alloc((size_t)n2 * sizeof(T));
  if (!c1 || !c2) return fail(NTK_ENOMEM, "workspace too small");
  const T inv_c = (T)(1.0 / (double)C);
  LAUNCH(env, k_rowdot<T>, grid_for((long long)n1 * 32), kThreads, 0, x1, x1, c1, (long long)n1, PairMap{1, 1}, C, inv_c);
  if (!symmetric)
    LAUNCH(env, k_rowdot<T>, grid_for((long long)n2 * 32), kThreads, 0, x2, x2, c2, (long long)n2, PairMap{1, 1}, C, inv_c);
  // per-sample variance chains (value in front of every activation): [n_act][n]
  int n_act = 0;
  for (int o = 0; o < fp.n; ++o) n_act += fp.kind[o] != NTK_OP_DENSE;
  T* qs1 = (T*)arena.alloc((size_t)std::max(1, n_act) * n1 * sizeof(T));
  T* qs2 = symmetric ? qs1 : (T*)arena.alloc((size_t)std::max(1, n_act) * n2 * sizeof(T));
  if (!qs1 || !qs2) return fail(NTK_ENOMEM, "workspace too small");
  LAUNCH(env, k_fcn_qchain<T>, grid_for(n1), kThreads, 0, (const T*)c1, n1, fp, qs1);
  if (!symmetric) LAUNCH(env, k_fcn_qchain<T>, grid_for(n2), kThreads, 0, (const T*)c2, n2, fp, qs2);
  // fp32: split x into TF32 hi / lo parts once, then the pipelined tcgen05 GEMM (gemm_kernels.cuh)
  static const bool no_tc = getenv("NTK_B200_NO_TC") != nullptr;
  static const bool no_pipe = getenv("NTK_B200_NO_GEMM_PIPE") != nullptr;
  // TMA-fed warp-specialised GEMM (gemm_tma.cu) is the default: 8192 x 8192 x 784 Gram + chain 3.39 -> 1.88 ms, results
  // bit-identical to the cp.async pipeline (profiles/check_gemm_tma.py); NTK_B200_GEMM_CPASYNC=1 switches back for A/B runs.
  static const bool use_tma = getenv("NTK_B200_GEMM_CPASYNC") == nullptr;
  float *a_hi = nullptr, *a_lo = nullptr, *b_hi = nullptr, *b_lo = nullptr;
  const int d_pad = gram_pad_k(C);
  bool pipe = false;
  if constexpr (std::is_same<T, float>::value) {
    if (C >= 16 && !no_tc && !no_pipe) {
      b_hi = (float*)arena.alloc((size_t)n2 * d_pad * sizeof(float));
      b_lo = (float*)arena.alloc((size_t)n2 * d_pad * sizeof(float));
      a_hi = symmetric ? b_hi : (float*)arena.alloc((size_t)n1 * d_pad * sizeof(float));
      a_lo = symmetric ? b_lo : (float*)arena.alloc((size_t)n1 * d_pad * sizeof(float));
      if (b_hi && b_lo && a_hi && a_lo) {
        pipe = true;
        env.launches++;
        NTK_TRY(launch_split_tf32(env.stream, x2, n2, C, b_hi, b_lo));
        if (!symmetric) {
          env.launches++;
          NTK_TRY(launch_split_tf32(env.stream, x1, n1, C, a_hi, a_lo));
        }
      }
    }
  }
  int t1 = n1;
  T* K0 = nullptr;
  for (;;) {
    K0 = (T*)arena.alloc((size_t)t1 * n2 * sizeof(T));
    if (K0) break;
    if (t1 <= 1) return fail(NTK_ENOMEM, "workspace too small for one row of the input Gram");
    t1 = (t1 + 1) / 2;
  }
  for (int r0 = 0; r0 < n1; r0 += t1) {
    const int a1 = std::min(t1, n1 - r0);
    if constexpr (std::is_same<T, float>::value) {
      if (pipe) {
        env.launches++;
        if (use_tma)
          NTK_TRY(launch_gram_tc_tma(env.stream, a_hi + (size_t)r0 * d_pad, a_lo + (size_t)r0 * d_pad, a1, b_hi, b_lo,
                                     n2, C, K0, (long long)n2));
        else
          NTK_TRY(launch_gram_tc_pipe(env.stream, a_hi + (size_t)r0 * d_pad, a_lo + (size_t)r0 * d_pad, a1, b_hi, b_lo,
                                      n2, C, K0, (long long)n2));
      }
    }
    if (!pipe)
      NTK_TRY((fcn_input_gram<T>(false, env.stream, &env.launches, x1 + (size_t)r0 * C, a1, x2, n2, C, K0)));
    LAUNCH(env, k_fcn_chain<T>, grid_for((long long)a1 * n2), kThreads, 0, (const T*)K0, (const T*)(qs1 + r0),
           (const T*)qs2, a1, n2, (long long)n1, (long long)n2, fp, out_nngp + (size_t)r0 * ld,
           want_ntk ? out_ntk + (size_t)r0 * ld : (T*)nullptr, ld);
  }
  return NTK_OK;
}

template <typename T>
int gram_device_t(ntk_context* ctx, const ntk_program* prog, const T* x1, int n1, const T* x2,
                  int n2, int H, int W, int C, uint32_t flags, const OutPtrs& out) {
  const bool want_ntk = (flags & NTK_FLAG_NTK) != 0;
  const bool symmetric = x2 == nullptr;
  if (symmetric) {
    x2 = x1;
    n2 = n1;
  }
  NTK_CUDA(cudaSetDevice(ctx->device));
  Env env;
  env.ctx = ctx;
  env.arena = &ctx->arena;
  env.stream = ctx->stream;

  // Fully-connected networks: tensor-core input Gram + one elementwise launch for the whole chain.
  if (!(flags & NTK_FLAG_NO_FUSION) && prog->fcn.n > 0 && H == 0 && !out.cov1 && !out.cov2) {
    ctx->arena.reset(ctx->ws, ctx->ws_bytes, false);
    int st = fcn_gram<T>(ctx, env, prog->fcn, x1, n1, x2, n2, symmetric, C, want_ntk, (T*)out.nngp,
                         (T*)out.ntk, out.ld);
    ctx->launches += env.launches;
    return st;
  }

  // Fast path: fused diagonal-marching kernels (fused_kernels.cuh).
  if (!(flags & NTK_FLAG_NO_FUSION) && prog->fused.ok && H > 0 &&
      fused_supported<T>(prog->fused, H, W, C) && !out.cov1 && !out.cov2) {
    ctx->arena.reset(ctx->ws, ctx->ws_bytes, false);
    const FusedPlan& plan = (flags & NTK_FLAG_PER_LAYER) ? prog->per_layer : prog->fused;
    int st = fused_gram<T>(plan, ctx->arena, ctx->stream, &env.launches, &ctx->prof, x1, n1, x2, n2,
                           symmetric, H, W, C, want_ntk, (T*)out.nngp, (T*)out.ntk, out.ld,
                           (flags & NTK_FLAG_FULL_SQUARE) != 0, (flags & NTK_FLAG_UPPER_ONLY) != 0);
    ctx->launches += env.launches;
    return st;
  }

  // Pool-free networks ending in Flatten: only the diagonal column matters (res_kernels.cuh).
  if (!(flags & NTK_FLAG_NO_FUSION) && diag_supported<T>(prog->diag, H, W) && !out.cov1 && !out.cov2) {
    ctx->arena.reset(ctx->ws, ctx->ws_bytes, false);
    int st = diag_gram<T>(prog->diag, ctx->arena, ctx->stream, &env.launches, x1, n1, x2, n2, symmetric,
                          H, C, want_ntk, (T*)out.nngp, (T*)out.ntk, out.ld,
                          (flags & NTK_FLAG_FULL_SQUARE) != 0);
    ctx->launches += env.launches;
    return st;
  }

  // Residual networks (WideResNet): fused column-sparse kernels (res_kernels.cuh).
  if (!(flags & NTK_FLAG_NO_FUSION) && prog->res.ok && H > 0 && res_supported<T>(prog->res, H, W, C) &&
      !out.cov1 && !out.cov2) {
    ctx->arena.reset(ctx->ws, ctx->ws_bytes, false);
    int st = res_gram<T>(prog->res, ctx->arena, ctx->stream, &env.launches, x1, n1, x2, n2, symmetric,
                         H, C, want_ntk, (T*)out.nngp, (T*)out.ntk, out.ld,
                         (flags & NTK_FLAG_FULL_SQUARE) != 0);
    ctx->launches += env.launches;
    return st;
  }

  int t1 = 0, t2 = 0;
  NTK_TRY(choose_tile<T>(*prog, ctx->ws_bytes, n1, n2, H, W, C, want_ntk, &t1, &t2));
  const size_t row = (size_t)(H > 0 ? (size_t)H * W * C : (size_t)C);
  for (int r0 = 0; r0 < n1; r0 += t1) {
    const int a = std::min(t1, n1 - r0);
    for (int c0 = 0; c0 < n2; c0 += t2) {
      const int b = std::min(t2, n2 - c0);
      ctx->arena.reset(ctx->ws, ctx->ws_bytes, false);
      env.bufs.clear();
      int st = run_tile_generic<T>(env, *prog, x1 + (size_t)r0 * row, a, x2 + (size_t)c0 * row, b,
                                   H, W, C, want_ntk, r0, c0, n2, out);
      if (st != NTK_OK) {
        ctx->launches += env.launches;
        return st;
      }
    }
  }
  ctx->launches += env.launches;
  return NTK_OK;
}

}  // namespace

// Kernel-in / Kernel-out (requirements.py:935-937).  `host`: `in`/`out` hold HOST pointers (copied through the
// workspace, the call synchronises); otherwise DEVICE pointers (wrapped without a copy, asynchronous).
template <typename T>
static int apply_t(ntk_context* ctx, const ntk_program* prog, const ntk_state_t* in, ntk_state_t* out, bool host) {
  const int t1 = in->n1, t2 = in->n2;
  const long long per_i = per_of(in->H, in->W);
  const long long P = (long long)t1 * t2;
  ctx->arena.reset(ctx->ws, ctx->ws_bytes, false);
  Env env;
  env.ctx = ctx;
  env.arena = &ctx->arena;
  env.stream = ctx->stream;
  std::vector<TState> slots(prog->n_slots);
  TState& s = slots[0];
  int st = NTK_OK;
  auto up = [&](const void* src, long long n, Buf** dst) -> int {
    if (host) {
      Buf* b = env.alloc((size_t)n * sizeof(T), &st);
      if (!b) return st;
      NTK_CUDA(cudaMemcpyAsync(b->p, src, (size_t)n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
      *dst = b;
    } else {
      // caller-owned device memory: two references, so in-place layer rules copy first (copy-on-write) and the
      // arena never sees the pointer
      env.bufs.emplace_back(new Buf());
      env.bufs.back()->p = const_cast<void*>(src);
      env.bufs.back()->refs = 2;
      *dst = env.bufs.back().get();
    }
    return NTK_OK;
  };
  NTK_TRY(up(in->nngp, P * per_i, &s.nngp));
  NTK_TRY(up(in->cov1, t1 * per_i, &s.cov1));
  NTK_TRY(up(in->cov2 ? in->cov2 : in->cov1, t2 * per_i, &s.cov2));
  s.ntk_mode = in->ntk_mode;
  if (in->ntk_mode == NTK_NTK_TENSOR) NTK_TRY(up(in->ntk, P * per_i, &s.ntk));
  s.H = in->H;
  s.W = in->W;
  s.gaussian = in->is_gaussian != 0;
  s.valid = true;
  NTK_TRY(run_ops<T>(env, *prog, slots, 0, t1, t2));
  TState& f = slots[prog->out_slot];
  if (!f.valid) return fail(NTK_EINVAL, "program left its output slot empty");
  const long long per_o = per_of(f.H, f.W);
  auto down = [&](void* dst, Buf* b, long long n) -> int {
    if (dst && b && dst != b->p)
      NTK_CUDA(cudaMemcpyAsync(dst, b->p, (size_t)n * sizeof(T),
                               host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, ctx->stream));
    return NTK_OK;
  };
  NTK_TRY(down(out->nngp, f.nngp, P * per_o));
  if (f.ntk_mode == NTK_NTK_TENSOR) NTK_TRY(down(out->ntk, f.ntk, P * per_o));
  NTK_TRY(down(out->cov1, f.cov1, t1 * per_o));
  if (in->cov2) NTK_TRY(down(out->cov2, f.cov2, t2 * per_o));
  out->n1 = t1;
  out->n2 = t2;
  out->H = f.H;
  out->W = f.W;
  out->ntk_mode = f.ntk_mode;
  out->is_gaussian = f.gaussian ? 1 : 0;
  if (host) NTK_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->launches += env.launches;
  return NTK_OK;
}

// Runs `body` with ctx->stream = `s` (NULL: the context's own stream).  No host synchronisation: workspace reuse
// between consecutive calls on different streams is ordered with an event the new stream waits on.
template <typename F>
static int with_stream(ntk_context* ctx, void* cuda_stream, F body) {
  NTK_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  if (ctx->order_valid && ctx->last_stream != s) NTK_CUDA(cudaStreamWaitEvent(s, ctx->order_ev, 0));
  ctx->stream = s;
  const int st = body();
  ctx->stream = ctx->own_stream;
  // even a failed call may have enqueued work that touches the workspace
  cudaError_t e = cudaEventRecord(ctx->order_ev, s);
  ctx->last_stream = s;
  ctx->order_valid = e == cudaSuccess;
  if (st == NTK_OK && e != cudaSuccess) return fail(NTK_ECUDA, "cudaEventRecord -> %s", cudaGetErrorString(e));
  return st;
}

// out[i, j] = slabs[row_of[i], j] (j >= i) or slabs[row_of[j], i] (j < i)
template <typename T>
__global__ void k_sym_assemble(const T* __restrict__ slabs, long long ld_s, const int* __restrict__ row_of, int n,
                               T* __restrict__ out, long long ld_o) {
  const long long total = (long long)n * n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx / n), j = (int)(idx % n);
    const int a = j >= i ? i : j, b = j >= i ? j : i;
    out[(long long)i * ld_o + j] = slabs[(long long)row_of[a] * ld_s + b];
  }
}

constexpr size_t kPinSlot = (size_t)8 << 20;

static bool is_pinned_host(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

static int ensure_pin(ntk_context* ctx) {
  if (ctx->pin) return NTK_OK;
  NTK_CUDA(cudaMallocHost((void**)&ctx->pin, 2 * kPinSlot));
  for (int k = 0; k < 2; ++k) NTK_CUDA(cudaEventCreateWithFlags(&ctx->pin_ev[k], cudaEventDisableTiming));
  return NTK_OK;
}

// Host -> device.  Page-locked sources are DMA'd directly; pageable ones go through the pinned ring in
// kPinSlot chunks, so the CPU copy of chunk k+1 overlaps the DMA of chunk k.
static int h2d_staged(ntk_context* ctx, void* dst, const void* src, size_t bytes) {
  if (bytes == 0) return NTK_OK;
  if (is_pinned_host(src)) {
    NTK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return NTK_OK;
  }
  NTK_TRY(ensure_pin(ctx));
  int slot = 0;
  for (size_t off = 0; off < bytes; off += kPinSlot, slot ^= 1) {
    const size_t len = std::min(kPinSlot, bytes - off);
    if (ctx->pin_busy[slot]) NTK_CUDA(cudaEventSynchronize(ctx->pin_ev[slot]));
    memcpy(ctx->pin + slot * kPinSlot, (const char*)src + off, len);
    NTK_CUDA(cudaMemcpyAsync((char*)dst + off, ctx->pin + slot * kPinSlot, len, cudaMemcpyHostToDevice, ctx->stream));
    NTK_CUDA(cudaEventRecord(ctx->pin_ev[slot], ctx->stream));
    ctx->pin_busy[slot] = true;
  }
  return NTK_OK;
}

// Device -> host, same scheme (two chunks in flight).  Returns with the data in `dst` for pageable destinations;
// for page-locked ones the copy is only enqueued (the caller synchronises the stream).
static int d2h_staged(ntk_context* ctx, void* dst, const void* src, size_t bytes) {
  if (bytes == 0) return NTK_OK;
  if (is_pinned_host(dst)) {
    NTK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return NTK_OK;
  }
  NTK_TRY(ensure_pin(ctx));
  for (int k = 0; k < 2; ++k)
    if (ctx->pin_busy[k]) {
      NTK_CUDA(cudaEventSynchronize(ctx->pin_ev[k]));
      ctx->pin_busy[k] = false;
    }
  const size_t n_chunks = (bytes + kPinSlot - 1) / kPinSlot;
  auto issue = [&](size_t k) -> int {
    const size_t off = k * kPinSlot, len = std::min(kPinSlot, bytes - off);
    NTK_CUDA(cudaMemcpyAsync(ctx->pin + (k & 1) * kPinSlot, (const char*)src + off, len, cudaMemcpyDeviceToHost,
                             ctx->stream));
    NTK_CUDA(cudaEventRecord(ctx->pin_ev[k & 1], ctx->stream));
    return NTK_OK;
  };
  NTK_TRY(issue(0));
  if (n_chunks > 1) NTK_TRY(issue(1));
  for (size_t k = 0; k < n_chunks; ++k) {
    const size_t off = k * kPinSlot, len = std::min(kPinSlot, bytes - off);
    NTK_CUDA(cudaEventSynchronize(ctx->pin_ev[k & 1]));
    memcpy((char*)dst + off, ctx->pin + (k & 1) * kPinSlot, len);
    if (k + 2 < n_chunks) NTK_TRY(issue(k + 2));
  }
  return NTK_OK;
}

// =================================== C ABI ===================================
extern "C" {

int ntk_abi_version(void) { return NTK_B200_ABI_VERSION; }

const char* ntk_last_error(void) { return last_error().c_str(); }

int ntk_device_count(int* count) {
  if (!count) return fail(NTK_EINVAL, "count is NULL");
  NTK_CUDA(cudaGetDeviceCount(count));
  return NTK_OK;
}

int ntk_program_create(const ntk_op_t* ops, int32_t n_ops, int32_t n_slots, int32_t out_slot,
                       ntk_program_t** out) {
  if (!out || (n_ops > 0 && !ops) || n_ops < 0) return fail(NTK_EINVAL, "bad arguments");
  std::unique_ptr<ntk_program> p(new ntk_program());
  p->ops.assign(ops, ops + n_ops);
  p->n_slots = n_slots;
  p->out_slot = out_slot;
  NTK_TRY(validate_program(*p));
  p->fcn = plan_fcn(p->ops, p->out_slot);
  p->fused = plan_fused(p->ops, p->n_slots, p->out_slot);
  p->per_layer = plan_fused(p->ops, p->n_slots, p->out_slot, 1);
  p->res = plan_resnet(p->ops, p->out_slot);
  p->diag = plan_diag(p->ops, p->last_use, p->out_slot);
  *out = p.release();
  return NTK_OK;
}

void ntk_program_destroy(ntk_program_t* prog) { delete prog; }

int ntk_program_output_shape(const ntk_program_t* prog, int32_t H, int32_t W,
                             int32_t in_is_gaussian, int32_t* out_H, int32_t* out_W,
                             int32_t* out_is_gaussian) {
  if (!prog) return fail(NTK_EINVAL, "prog is NULL");
  struct S {
    int H, W;
    bool g, valid;
  };
  std::vector<S> slots(prog->n_slots, S{0, 0, false, false});
  slots[0] = S{H, W, in_is_gaussian != 0, true};
  for (size_t k = 0; k < prog->ops.size(); ++k) {
    const ntk_op_t& op = prog->ops[k];
    S in = slots[op.src];
    if (!in.valid) return fail(NTK_EINVAL, "op %zu reads empty slot", k);
    S o = in;
    switch (op.kind) {
      case NTK_OP_CONV:
      case NTK_OP_AVGPOOL: {
        if (in.H <= 0) return fail(NTK_EINVAL, "spatial op on non-spatial input");
        AxisGeom gh = axis_geom(in.H, op.i[0], op.i[2], op.i[4]);
        AxisGeom gw = axis_geom(in.W, op.i[1], op.i[3], op.i[4]);
        if (gh.out <= 0 || gw.out <= 0) return fail(NTK_EINVAL, "empty output");
        o.H = gh.out;
        o.W = gw.out;
        if (op.kind == NTK_OP_CONV) o.g = true;
        break;
      }
      case NTK_OP_DENSE:
        o.g = true;
        break;
      case NTK_OP_ABRELU:
      case NTK_OP_ERF:
      case NTK_OP_GELU:
      case NTK_OP_SIN:
      case NTK_OP_RBF:
        if (!in.g) return fail(NTK_ENOTGAUSSIAN, "The input to the activation function must be Gaussian");
        o.g = false;
        break;
      case NTK_OP_LAYERNORM:
        if (!in.g) return fail(NTK_EUNSUPPORTED, "LayerNorm only implemented for Gaussian inputs.");
        break;
      case NTK_OP_GAP:
        if (in.H <= 0) return fail(NTK_EINVAL, "GlobalAvgPool needs spatial inputs");
        o.H = o.W = 0;
        break;
      case NTK_OP_FLATTEN:
        o.H = o.W = 0;
        o.g = false;
        break;
      case NTK_OP_FANINSUM: {
        S b = slots[op.src2];
        if (b.H != in.H || b.W != in.W) return fail(NTK_ESHAPE, "All shapes should be equal in `FanInSum`");
        if (!(in.g && b.g)) return fail(NTK_EUNSUPPORTED, "`FanInSum` needs Gaussian inputs");
        o.g = true;
        break;
      }
      default:
        break;
    }
    slots[op.dst] = o;
  }
  S f = slots[prog->out_slot];
  if (out_H) *out_H = f.H;
  if (out_W) *out_W = f.W;
  if (out_is_gaussian) *out_is_gaussian = f.g ? 1 : 0;
  return NTK_OK;
}

// Mirrors the dispatch order of gram_device_t (above); keep the two in step.
int ntk_program_path(const ntk_program_t* prog, int32_t dtype, int32_t H, int32_t W, int32_t C,
                     uint32_t flags, int32_t* path) {
  if (!prog || !path) return fail(NTK_EINVAL, "bad arguments");
  if (dtype != NTK_F32 && dtype != NTK_F64) return fail(NTK_EINVAL, "unknown dtype %d", dtype);
  const bool fusable = !(flags & NTK_FLAG_NO_FUSION) && !(flags & NTK_FLAG_WANT_COV);
  *path = NTK_PATH_GENERIC;
  if (!fusable) return NTK_OK;
  if (prog->fcn.n > 0 && H == 0) {
    *path = NTK_PATH_FCN;
  } else if (prog->fused.ok && H > 0 && fused_supported<float>(prog->fused, H, W, C)) {
    *path = NTK_PATH_FUSED;
  } else if (dtype == NTK_F32 ? diag_supported<float>(prog->diag, H, W) : diag_supported<double>(prog->diag, H, W)) {
    *path = NTK_PATH_DIAG;
  } else if (prog->res.ok && H > 0 && res_supported<float>(prog->res, H, W, C)) {
    *path = NTK_PATH_RES;
  }
  return NTK_OK;
}

int ntk_context_create(int32_t device, size_t workspace_bytes, ntk_context_t** out) {
  if (!out) return fail(NTK_EINVAL, "out is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(NTK_ECUDA, "no CUDA device available (%s); ntk_b200 has no CPU fallback",
                cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(NTK_EINVAL, "device %d out of range [0,%d)", device, n);
  NTK_CUDA(cudaSetDevice(device));
  std::unique_ptr<ntk_context> c(new ntk_context());
  c->device = device;
  NTK_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  c->stream = c->own_stream;
  NTK_CUDA(cudaEventCreateWithFlags(&c->order_ev, cudaEventDisableTiming));
  if (workspace_bytes == 0) {
    size_t free_b = 0, total_b = 0;
    NTK_CUDA(cudaMemGetInfo(&free_b, &total_b));
    // default: 70 % of the free HBM, at most 128 GB (a B200 has 180 GB; big tiles amortise the launch tails)
    workspace_bytes = std::min<size_t>((size_t)(free_b * 0.7), (size_t)128 << 30);
  }
  NTK_CUDA(cudaMalloc((void**)&c->ws, workspace_bytes));
  c->ws_bytes = workspace_bytes;
  NTK_TRY((fused_configure_device<float, false>()));
  NTK_TRY((fused_configure_device<double, false>()));
  NTK_TRY((fused_configure_device<float, true>()));
  NTK_TRY((fused_configure_device<double, true>()));
  NTK_TRY((fused_configure_device<float, 2>()));
  NTK_TRY((fused_configure_device<double, 2>()));
  NTK_TRY(stage_packed_configure());
  NTK_TRY(stage_packed_erf_configure());
  *out = c.release();
  return NTK_OK;
}

void ntk_context_destroy(ntk_context_t* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->own_stream);
  if (ctx->order_valid) cudaEventSynchronize(ctx->order_ev);
  for (int k = 0; k < 6; ++k)
    if (ctx->io[k]) cudaFree(ctx->io[k]);
  if (ctx->ws) cudaFree(ctx->ws);
  if (ctx->pin) cudaFreeHost(ctx->pin);
  for (int k = 0; k < 2; ++k)
    if (ctx->pin_ev[k]) cudaEventDestroy(ctx->pin_ev[k]);
  if (ctx->order_ev) cudaEventDestroy(ctx->order_ev);
  cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

int ntk_context_synchronize(ntk_context_t* ctx) {
  if (!ctx) return fail(NTK_EINVAL, "ctx is NULL");
  NTK_CUDA(cudaSetDevice(ctx->device));
  NTK_CUDA(cudaStreamSynchronize(ctx->stream));
  return NTK_OK;
}

int ntk_context_set_profiling(ntk_context_t* ctx, int32_t enabled) {
  if (!ctx) return fail(NTK_EINVAL, "ctx is NULL");
  NTK_CUDA(cudaSetDevice(ctx->device));
  NTK_CUDA(cudaStreamSynchronize(ctx->stream));
  StageProfile& p = ctx->prof;
  for (int s = 0; s < StageProfile::kMaxStages; ++s) {
    for (auto& ev : p.pending[s]) {
      cudaEventDestroy(ev.first);
      cudaEventDestroy(ev.second);
    }
    p.pending[s].clear();
    p.pending_pairs[s].clear();
    p.total_ms[s] = 0;
    p.launches[s] = 0;
    p.pairs[s] = 0;
  }
  p.enabled = enabled != 0;
  return NTK_OK;
}

int ntk_context_profile(ntk_context_t* ctx, int32_t stage, double* total_ms, int64_t* launches,
                        int64_t* pairs) {
  if (!ctx || stage < 0 || stage >= StageProfile::kMaxStages) return fail(NTK_EINVAL, "bad arguments");
  NTK_CUDA(cudaSetDevice(ctx->device));
  NTK_CUDA(cudaStreamSynchronize(ctx->stream));
  StageProfile& p = ctx->prof;
  for (size_t k = 0; k < p.pending[stage].size(); ++k) {
    float ms = 0.f;
    NTK_CUDA(cudaEventElapsedTime(&ms, p.pending[stage][k].first, p.pending[stage][k].second));
    p.total_ms[stage] += ms;
    p.launches[stage] += 1;
    p.pairs[stage] += p.pending_pairs[stage][k];
    cudaEventDestroy(p.pending[stage][k].first);
    cudaEventDestroy(p.pending[stage][k].second);
  }
  p.pending[stage].clear();
  p.pending_pairs[stage].clear();
  if (total_ms) *total_ms = p.total_ms[stage];
  if (launches) *launches = p.launches[stage];
  if (pairs) *pairs = p.pairs[stage];
  return NTK_OK;
}

void* ntk_context_stream(ntk_context_t* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int64_t ntk_context_launch_count(const ntk_context_t* ctx) { return ctx ? ctx->launches : 0; }

int ntk_gram_device_on_stream(ntk_context_t* ctx, const ntk_program_t* prog, int32_t dtype, const void* x1,
                              int32_t n1, const void* x2, int32_t n2, int32_t H, int32_t W, int32_t C,
                              uint32_t flags, void* nngp, void* ntk, int64_t ld, void* cov1, void* cov2,
                              void* cuda_stream) {
  if (!ctx || !prog || !x1 || n1 <= 0 || C <= 0 || (x2 && n2 <= 0))
    return fail(NTK_EINVAL, "bad arguments");
  if ((H > 0) != (W > 0) || H < 0) return fail(NTK_EINVAL, "bad spatial shape");
  if ((flags & NTK_FLAG_NTK) && !ntk) return fail(NTK_EINVAL, "NTK requested but ntk is NULL");
  if (dtype != NTK_F32 && dtype != NTK_F64) return fail(NTK_EINVAL, "unknown dtype %d", dtype);
  OutPtrs o{nngp, (flags & NTK_FLAG_NTK) ? ntk : nullptr,
            (flags & NTK_FLAG_WANT_COV) ? cov1 : nullptr,
            (flags & NTK_FLAG_WANT_COV) ? cov2 : nullptr, ld};
  return with_stream(ctx, cuda_stream, [&]() -> int {
    if (dtype == NTK_F32)
      return gram_device_t<float>(ctx, prog, (const float*)x1, n1, (const float*)x2, n2, H, W, C, flags, o);
    return gram_device_t<double>(ctx, prog, (const double*)x1, n1, (const double*)x2, n2, H, W, C, flags, o);
  });
}

int ntk_gram_device(ntk_context_t* ctx, const ntk_program_t* prog, int32_t dtype, const void* x1,
                    int32_t n1, const void* x2, int32_t n2, int32_t H, int32_t W, int32_t C,
                    uint32_t flags, void* nngp, void* ntk, int64_t ld, void* cov1, void* cov2) {
  return ntk_gram_device_on_stream(ctx, prog, dtype, x1, n1, x2, n2, H, W, C, flags, nngp, ntk, ld, cov1, cov2,
                                   nullptr);
}

int ntk_gram_host(ntk_context_t* ctx, const ntk_program_t* prog, int32_t dtype, const void* x1,
                  int32_t n1, const void* x2, int32_t n2, int32_t H, int32_t W, int32_t C,
                  uint32_t flags, void* nngp, void* ntk, int64_t ld, void* cov1, void* cov2) {
  if (!ctx || !prog || !x1 || n1 <= 0 || C <= 0) return fail(NTK_EINVAL, "bad arguments");
  if (dtype != NTK_F32 && dtype != NTK_F64) return fail(NTK_EINVAL, "unknown dtype %d", dtype);
  NTK_CUDA(cudaSetDevice(ctx->device));
  const size_t sz = dtype == NTK_F32 ? 4 : 8;
  const size_t row = (size_t)(H > 0 ? (size_t)H * W * C : (size_t)C);
  const int m2 = x2 ? n2 : n1;
  int oh = 0, ow = 0, og = 0;
  NTK_TRY(ntk_program_output_shape(prog, H, W, 0, &oh, &ow, &og));
  const size_t per = (size_t)per_of(oh, ow);
  const size_t ldd = oh > 0 ? (size_t)m2 : (size_t)m2;  // device results are dense [n1, m2]
  const bool want_ntk = (flags & NTK_FLAG_NTK) != 0;
  const bool want_cov = (flags & NTK_FLAG_WANT_COV) != 0;
  NTK_TRY(ensure_io(ctx, 0, (size_t)n1 * row * sz));
  if (x2) NTK_TRY(ensure_io(ctx, 1, (size_t)n2 * row * sz));
  NTK_TRY(ensure_io(ctx, 2, (size_t)n1 * ldd * per * sz));
  if (want_ntk) NTK_TRY(ensure_io(ctx, 3, (size_t)n1 * ldd * per * sz));
  if (want_cov) {
    NTK_TRY(ensure_io(ctx, 4, (size_t)n1 * per * sz));
    if (x2) NTK_TRY(ensure_io(ctx, 5, (size_t)n2 * per * sz));
  }
  if (ctx->order_valid && ctx->last_stream != ctx->own_stream) {
    NTK_CUDA(cudaStreamWaitEvent(ctx->own_stream, ctx->order_ev, 0));
    ctx->last_stream = ctx->own_stream;
  }
  NTK_TRY(h2d_staged(ctx, ctx->io[0], x1, (size_t)n1 * row * sz));
  if (x2) NTK_TRY(h2d_staged(ctx, ctx->io[1], x2, (size_t)n2 * row * sz));
  NTK_TRY(ntk_gram_device(ctx, prog, dtype, ctx->io[0], n1, x2 ? ctx->io[1] : nullptr, n2, H, W, C,
                          flags, ctx->io[2], want_ntk ? ctx->io[3] : nullptr, (int64_t)ldd,
                          want_cov ? ctx->io[4] : nullptr, (want_cov && x2) ? ctx->io[5] : nullptr));
  auto d2h = [&](void* dst, const void* src) -> int {
    if (oh > 0 || (size_t)ld == ldd) {
      NTK_TRY(d2h_staged(ctx, dst, src, (size_t)n1 * ldd * per * sz));
    } else {
      NTK_CUDA(cudaMemcpy2DAsync(dst, (size_t)ld * sz, src, ldd * sz, (size_t)m2 * sz, (size_t)n1,
                                 cudaMemcpyDeviceToHost, ctx->stream));
    }
    return NTK_OK;
  };
  if (nngp) NTK_TRY(d2h(nngp, ctx->io[2]));
  if (want_ntk && ntk) NTK_TRY(d2h(ntk, ctx->io[3]));
  if (want_cov && cov1)
    NTK_CUDA(cudaMemcpyAsync(cov1, ctx->io[4], (size_t)n1 * per * sz, cudaMemcpyDeviceToHost, ctx->stream));
  if (want_cov && cov2 && x2)
    NTK_CUDA(cudaMemcpyAsync(cov2, ctx->io[5], (size_t)n2 * per * sz, cudaMemcpyDeviceToHost, ctx->stream));
  NTK_CUDA(cudaStreamSynchronize(ctx->stream));
  return NTK_OK;
}

int ntk_apply_host(ntk_context_t* ctx, const ntk_program_t* prog, int32_t dtype,
                   const ntk_state_t* in, ntk_state_t* out) {
  if (!ctx || !prog || !in || !out || !in->nngp || !in->cov1 || in->n1 <= 0 || in->n2 <= 0)
    return fail(NTK_EINVAL, "bad arguments");
  if (dtype != NTK_F32 && dtype != NTK_F64) return fail(NTK_EINVAL, "unknown dtype %d", dtype);
  return with_stream(ctx, nullptr, [&]() -> int {
    return dtype == NTK_F32 ? apply_t<float>(ctx, prog, in, out, true) : apply_t<double>(ctx, prog, in, out, true);
  });
}

int ntk_apply_device(ntk_context_t* ctx, const ntk_program_t* prog, int32_t dtype, const ntk_state_t* in,
                     ntk_state_t* out, void* cuda_stream) {
  if (!ctx || !prog || !in || !out || !in->nngp || !in->cov1 || in->n1 <= 0 || in->n2 <= 0)
    return fail(NTK_EINVAL, "bad arguments");
  if (dtype != NTK_F32 && dtype != NTK_F64) return fail(NTK_EINVAL, "unknown dtype %d", dtype);
  return with_stream(ctx, cuda_stream, [&]() -> int {
    return dtype == NTK_F32 ? apply_t<float>(ctx, prog, in, out, false) : apply_t<double>(ctx, prog, in, out, false);
  });
}

int ntk_sym_assemble(ntk_context_t* ctx, int32_t dtype, const void* slabs, int64_t ld_slabs,
                     const int32_t* row_of, int32_t n, void* out, int64_t ld_out) {
  if (!ctx || !slabs || !row_of || !out || n <= 0) return fail(NTK_EINVAL, "bad arguments");
  if (dtype != NTK_F32 && dtype != NTK_F64) return fail(NTK_EINVAL, "unknown dtype %d", dtype);
  return with_stream(ctx, nullptr, [&]() -> int {
    ctx->launches++;
    if (dtype == NTK_F32)
      k_sym_assemble<float><<<grid_for((long long)n * n), kThreads, 0, ctx->stream>>>(
          (const float*)slabs, ld_slabs, row_of, n, (float*)out, ld_out);
    else
      k_sym_assemble<double><<<grid_for((long long)n * n), kThreads, 0, ctx->stream>>>(
          (const double*)slabs, ld_slabs, row_of, n, (double*)out, ld_out);
    NTK_CUDA(cudaGetLastError());
    return NTK_OK;
  });
}

int ntk_workspace_bytes(const ntk_program_t* prog, int32_t dtype, int32_t t1, int32_t t2, int32_t H,
                        int32_t W, int32_t C, uint32_t flags, size_t* bytes) {
  if (!prog || !bytes || t1 <= 0 || t2 <= 0) return fail(NTK_EINVAL, "bad arguments");
  const bool want_ntk = (flags & NTK_FLAG_NTK) != 0;
  if (dtype == NTK_F32) return dry_peak_generic<float>(*prog, t1, t2, H, W, C, want_ntk, bytes);
  if (dtype == NTK_F64) return dry_peak_generic<double>(*prog, t1, t2, H, W, C, want_ntk, bytes);
  return fail(NTK_EINVAL, "unknown dtype %d", dtype);
}

int ntk_device_malloc(int32_t device, size_t bytes, void** ptr) {
  if (!ptr) return fail(NTK_EINVAL, "ptr is NULL");
  NTK_CUDA(cudaSetDevice(device));
  NTK_CUDA(cudaMalloc(ptr, bytes));
  return NTK_OK;
}

int ntk_device_free(int32_t device, void* ptr) {
  NTK_CUDA(cudaSetDevice(device));
  NTK_CUDA(cudaFree(ptr));
  return NTK_OK;
}

int ntk_memcpy_h2d(ntk_context_t* ctx, void* dst_dev, const void* src_host, size_t bytes) {
  if (!ctx) return fail(NTK_EINVAL, "ctx is NULL");
  NTK_CUDA(cudaSetDevice(ctx->device));
  NTK_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return NTK_OK;
}

int ntk_memcpy_d2h(ntk_context_t* ctx, void* dst_host, const void* src_dev, size_t bytes) {
  if (!ctx) return fail(NTK_EINVAL, "ctx is NULL");
  NTK_CUDA(cudaSetDevice(ctx->device));
  NTK_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  NTK_CUDA(cudaStreamSynchronize(ctx->stream));
  return NTK_OK;
}

int ntk_memset_async(ntk_context_t* ctx, void* dst_dev, int32_t value, size_t bytes) {
  if (!ctx) return fail(NTK_EINVAL, "ctx is NULL");
  NTK_CUDA(cudaSetDevice(ctx->device));
  NTK_CUDA(cudaMemsetAsync(dst_dev, value, bytes, ctx->stream));
  return NTK_OK;
}

int ntk_context_device(const ntk_context_t* ctx) { return ctx ? ctx->device : -1; }

int ntk_host_alloc(size_t bytes, void** ptr) {
  if (!ptr) return fail(NTK_EINVAL, "ptr is NULL");
  NTK_CUDA(cudaMallocHost(ptr, bytes));
  return NTK_OK;
}

int ntk_host_free(void* ptr) {
  NTK_CUDA(cudaFreeHost(ptr));
  return NTK_OK;
}

int ntk_event_create(void** event) {
  if (!event) return fail(NTK_EINVAL, "event is NULL");
  cudaEvent_t e;
  NTK_CUDA(cudaEventCreate(&e));
  *event = (void*)e;
  return NTK_OK;
}

int ntk_event_record(ntk_context_t* ctx, void* event) {
  if (!ctx || !event) return fail(NTK_EINVAL, "bad arguments");
  NTK_CUDA(cudaSetDevice(ctx->device));
  NTK_CUDA(cudaEventRecord((cudaEvent_t)event, ctx->stream));
  return NTK_OK;
}

int ntk_event_elapsed_ms(void* start, void* stop, float* ms) {
  if (!start || !stop || !ms) return fail(NTK_EINVAL, "bad arguments");
  NTK_CUDA(cudaEventSynchronize((cudaEvent_t)stop));
  NTK_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
  return NTK_OK;
}

void ntk_event_destroy(void* event) {
  if (event) cudaEventDestroy((cudaEvent_t)event);
}

int ntk_stream_create(int32_t device, void** cuda_stream) {
  if (!cuda_stream) return fail(NTK_EINVAL, "cuda_stream is NULL");
  NTK_CUDA(cudaSetDevice(device));
  cudaStream_t s;
  NTK_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  *cuda_stream = (void*)s;
  return NTK_OK;
}

int ntk_stream_synchronize(void* cuda_stream) {
  NTK_CUDA(cudaStreamSynchronize((cudaStream_t)cuda_stream));
  return NTK_OK;
}

int ntk_stream_query(void* cuda_stream, int32_t* done) {
  if (!done) return fail(NTK_EINVAL, "done is NULL");
  cudaError_t e = cudaStreamQuery((cudaStream_t)cuda_stream);
  if (e != cudaSuccess && e != cudaErrorNotReady) NTK_CUDA(e);
  *done = e == cudaSuccess ? 1 : 0;
  return NTK_OK;
}

void ntk_stream_destroy(void* cuda_stream) {
  if (cuda_stream) cudaStreamDestroy((cudaStream_t)cuda_stream);
}

}  // extern "C"
