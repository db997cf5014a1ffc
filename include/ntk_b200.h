/*
 * ntk_b200.h — C-ABI of the B200-native analytic NNGP/NTK kernel path.
 *
 * Drop-in boundary for google/neural-tangents' `stax` `kernel_fn(x1, x2, get)`
 * and the `nt.batch` Gram tiling.  The reference has no FFI of its own (it is
 * pure Python on JAX/XLA); every entry point below cites the reference
 * interface it replaces (paths relative to /root/reference/neural_tangents/).
 * INTEGRATION.md shows the reference-side binding (ctypes today, an XLA-FFI
 * handler over the same symbols when jaxlib headers are available).
 *
 * Conventions
 *  - extern "C", plain pointers and sizes only; no C++/torch/JAX types.
 *  - Every function returns 0 on success or a negative NTK_E* code and never
 *    throws; `ntk_last_error()` returns a thread-local message.
 *  - dtype: NTK_F32 or NTK_F64 is both the storage and the arithmetic type
 *    (`_src/stax/requirements.py:794`: float32 unless jax_enable_x64).
 *  - Spatial tensors use the canonical zipped layout [n1, n2, h, h', w, w']
 *    (`_src/utils/kernel.py:32-36` with is_reversed == False); the Python
 *    front end presents the reference's per-Conv axis reversal lazily.
 *  - There is NO CPU fallback: every compute entry point needs a CUDA device.
 */
#ifndef NTK_B200_H_
#define NTK_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NTK_B200_ABI_VERSION 2

/* ---- status codes ------------------------------------------------------ */
#define NTK_OK 0
#define NTK_EINVAL (-1)        /* bad argument / malformed program            */
#define NTK_ECUDA (-2)         /* CUDA runtime error (see ntk_last_error)     */
#define NTK_ENOMEM (-3)        /* workspace too small / allocation failed     */
#define NTK_ENOTGAUSSIAN (-4)  /* activation on a non-Gaussian input
                                  (`_src/stax/elementwise.py:1267-1270`)     */
#define NTK_EUNSUPPORTED (-5)  /* valid in the reference, outside this path   */
#define NTK_ESHAPE (-6)        /* FanInSum shape mismatch
                                  (`_src/stax/branching.py:71-75`)           */

/* ---- dtypes ------------------------------------------------------------ */
#define NTK_F32 0
#define NTK_F64 1

/* ---- op kinds (one per in-scope reference layer rule) -------------------- */
enum {
  NTK_OP_DENSE = 1,    /* `_src/stax/linear.py:899-926`   f[0]=W_std^2 f[1]=b_std^2 (i[0]=has_bias)      */
  NTK_OP_CONV = 2,     /* `_src/stax/linear.py:1321-1424` i = {kh,kw,sh,sw,padding,has_bias}
                                                          f[0]=W_std^2 f[1]=b_std^2                      */
  NTK_OP_ABRELU = 3,   /* `_src/stax/elementwise.py:423-477` f[0]=a f[1]=b, i[0]=do_stabilize           */
  NTK_OP_ERF = 4,      /* `_src/stax/elementwise.py:67-112`  f[0]=a f[1]=b f[2]=c                        */
  NTK_OP_AVGPOOL = 5,  /* `_src/stax/linear.py:1631-1664,3499-3572`
                                                          i = {wh,ww,sh,sw,padding,flags}: flags bit 0 =
                                                          normalize_edges, bit 1 = SumPool (`:1503`, no division) */
  NTK_OP_GAP = 6,      /* `_src/stax/linear.py:1771-1801`  i[0] = 1: GlobalSumPool (`:1674`)              */
  NTK_OP_FLATTEN = 7,  /* `_src/stax/linear.py:1865-1899`                                               */
  NTK_OP_FANINSUM = 8, /* `_src/stax/branching.py:55-117`  dst = src + src2                              */
  NTK_OP_IDENTITY = 9, /* `_src/stax/linear.py:107-119`    dst = src (FanOut is expressed by slot reuse) */
  NTK_OP_GELU = 10,    /* `_src/stax/elementwise.py:195-263`                                             */
  NTK_OP_SIN = 11,     /* `_src/stax/elementwise.py:266-341` a sin(b x + c): f = {a, b, c} (Cos: c + pi/2) */
  NTK_OP_RBF = 12,     /* `_src/stax/elementwise.py:344-400` f[0] = gamma                                */
  NTK_OP_LAYERNORM = 13 /* `_src/stax/linear.py:2476-2590` over the channel axis: f[0] = eps              */
};

/* padding modes (`_src/stax/linear.py:54-69`) */
#define NTK_PAD_VALID 0
#define NTK_PAD_SAME 1
#define NTK_PAD_CIRCULAR 2

/*
 * One layer application in slot (SSA-like) form: dst = op(src [, src2]).
 * `serial` (`_src/stax/combinators.py:40-68`) becomes consecutive ops;
 * `FanOut`/`parallel`/`FanInSum` (`_src/stax/branching.py:36-117`,
 * `_src/stax/combinators.py:166-198`) become slot reuse + NTK_OP_FANINSUM.
 * Slot 0 holds the input kernel (built from x1/x2 or supplied by the caller).
 */
typedef struct ntk_op {
  int32_t kind;
  int32_t src;
  int32_t src2; /* -1 unless NTK_OP_FANINSUM */
  int32_t dst;
  int32_t i[6];
  double f[4];
} ntk_op_t;

/* A kernel state crossing the ABI (Kernel-in / Kernel-out composition,
 * `_src/stax/requirements.py:935-937,1039-1044`; fields of
 * `_src/utils/kernel.py:124-145`).  Pointers are HOST pointers for the *_host
 * entry points and DEVICE pointers for the *_device ones.  Spatial tensors are
 * [n1,n2,H,H,W,W] (nngp, ntk) and [n,H,H,W,W] (cov); H == W == 0 means the
 * spatial axes have been pooled away ([n1,n2] and [n]). */
typedef struct ntk_state {
  void* nngp;
  void* ntk;  /* NULL when ntk_mode != NTK_NTK_TENSOR */
  void* cov1;
  void* cov2; /* NULL <=> x2 is None (cov2 == cov1) */
  int32_t n1, n2;
  int32_t H, W;
  int32_t ntk_mode;    /* NTK_NTK_* below */
  int32_t is_gaussian; /* `_src/utils/kernel.py:131` */
} ntk_state_t;

#define NTK_NTK_NONE 0   /* ntk not requested (`requirements.py:943`)         */
#define NTK_NTK_ZERO 1   /* the 0-d zero of `requirements.py:807`             */
#define NTK_NTK_TENSOR 2

/* flags for the gram entry points */
#define NTK_FLAG_NTK 1u          /* compute ntk as well as nngp                */
#define NTK_FLAG_NO_FUSION 2u    /* force the general one-kernel-per-op path   */
#define NTK_FLAG_FULL_SQUARE 16u /* x2 == NULL: compute all n x n entries like the
                                    reference (`_src/batching.py:370`) instead of
                                    the upper triangle + mirror                  */
#define NTK_FLAG_PER_LAYER 8u    /* stencil kernels, but ONE Conv+ABRelu layer per
                                    launch: every layer makes one HBM round trip
                                    (the traffic model of the roofline)          */
#define NTK_FLAG_WANT_COV 4u     /* also return cov1 / cov2                    */
#define NTK_FLAG_UPPER_ONLY 32u  /* x2 != NULL and x1 is the SAME samples as x2[0:n1] (a row slab of a symmetric
                                    Gram, columns starting at the slab's first row): entries (i, j < i) are not
                                    needed and may be left unwritten.  The fused path enumerates only the upper
                                    trapezoid; the other paths ignore the hint.  This is the building block of the
                                    triangular multi-GPU schedule the reference leaves as a TODO
                                    (`_src/batching.py:370`). */

typedef struct ntk_program ntk_program_t;
typedef struct ntk_context ntk_context_t;

/* ---- library ------------------------------------------------------------ */
int ntk_abi_version(void);
const char* ntk_last_error(void);
int ntk_device_count(int* count);

/* ---- programs (replace the closure tree built by `stax.serial(...)`) ---- */
int ntk_program_create(const ntk_op_t* ops, int32_t n_ops, int32_t n_slots, int32_t out_slot,
                       ntk_program_t** out);
void ntk_program_destroy(ntk_program_t* prog);

/* Output geometry of `prog` for inputs of spatial size H x W (0,0 for [N,d]
 * inputs): replaces the `eval_shape` walk of `requirements.py:833-879`.     */
int ntk_program_output_shape(const ntk_program_t* prog, int32_t H, int32_t W,
                             int32_t in_is_gaussian, int32_t* out_H, int32_t* out_W,
                             int32_t* out_is_gaussian);

/* Which kernel family ntk_gram_* runs `prog` on for inputs [n, H, W, C] (H == W == 0: [n, C]) under `flags`
 * (INTEGRATION.md §5).  Pure host logic -- no GPU needed -- so a caller (or a test) can see
 * that a network is on a fused path rather than on the one-kernel-per-op path. */
#define NTK_PATH_GENERIC 0 /* one kernel per op on the canonical layout            */
#define NTK_PATH_FUSED 1   /* fused stage kernels (k_stage_p / k_stage)            */
#define NTK_PATH_RES 2     /* column-sparse residual kernels (k_res)               */
#define NTK_PATH_DIAG 3    /* diagonal-column kernel (k_diagnet)                   */
#define NTK_PATH_FCN 4     /* tensor-core input Gram + k_fcn_chain                 */
int ntk_program_path(const ntk_program_t* prog, int32_t dtype, int32_t H, int32_t W, int32_t C,
                     uint32_t flags, int32_t* path);

/* ---- contexts: one per (host thread, GPU) -------------------------------
 * Own a stream, a device workspace and pinned staging buffers; replace the
 * PjRt client/executable cache behind `jit`/`pmap` in `_src/batching.py:689-787`. */
int ntk_context_create(int32_t device, size_t workspace_bytes, ntk_context_t** out);
void ntk_context_destroy(ntk_context_t* ctx);
int ntk_context_synchronize(ntk_context_t* ctx);
/* cudaStream_t of the context as an opaque pointer (for event timing by the host). */
void* ntk_context_stream(ntk_context_t* ctx);
/* Number of CUDA kernels this context has launched so far (bench `gpu_launches`). */
int64_t ntk_context_launch_count(const ntk_context_t* ctx);

/* Per-kernel device timing for bench.py's roofline line: when enabled, the fused path
 * brackets every stage kernel of the cross-pair pass with CUDA events on the context
 * stream.  `ntk_context_profile` returns, for stage `stage` (0 = first fused stage), the
 * accumulated device time, the number of launches and the number of sample pairs they
 * processed since profiling was (re-)enabled; it synchronises the stream first. */
int ntk_context_set_profiling(ntk_context_t* ctx, int32_t enabled);
int ntk_context_profile(ntk_context_t* ctx, int32_t stage, double* total_ms, int64_t* launches,
                        int64_t* pairs);

/* ---- the hot path --------------------------------------------------------
 * kernel_fn(x1, x2, get) on raw inputs (`requirements.py:939-953`):
 *   x1: [n1, H, W, C] (or [n1, C] with H == W == 0), x2 likewise or NULL (== x1).
 * Results: nngp/ntk [n1, n2] row-major with leading dimension `ld` (elements),
 * or [n1, n2, Ho, Ho, Wo, Wo] when the program keeps spatial axes (then `ld`
 * is ignored).  cov1/cov2 ([n1..], [n2..]) only with NTK_FLAG_WANT_COV.
 * The call tiles the n1 x n2 pair grid internally to fit the workspace
 * (the role of `nt.batch`'s serial loop, `batching.py:314-502`).
 */

/* HOST buffers in, HOST buffers out; copies run on the context stream and the
 * call returns after the results have landed. */
int ntk_gram_host(ntk_context_t* ctx, const ntk_program_t* prog, int32_t dtype, const void* x1,
                  int32_t n1, const void* x2, int32_t n2, int32_t H, int32_t W, int32_t C,
                  uint32_t flags, void* nngp, void* ntk, int64_t ld, void* cov1, void* cov2);

/* DEVICE buffers in/out, asynchronous on the context stream. */
int ntk_gram_device(ntk_context_t* ctx, const ntk_program_t* prog, int32_t dtype, const void* x1,
                    int32_t n1, const void* x2, int32_t n2, int32_t H, int32_t W, int32_t C,
                    uint32_t flags, void* nngp, void* ntk, int64_t ld, void* cov1, void* cov2);

/* Kernel-in / Kernel-out: apply `prog` to a caller-supplied kernel state
 * (`requirements.py:935-937`).  `out` pointers must be preallocated to the
 * geometry reported by ntk_program_output_shape.  HOST pointers. */
int ntk_apply_host(ntk_context_t* ctx, const ntk_program_t* prog, int32_t dtype,
                   const ntk_state_t* in, ntk_state_t* out);

/* Workspace (bytes) one tile of t1 x t2 pairs needs; lets callers size
 * contexts (`README.md:387-396` documents the reference's batch-size limits). */
int ntk_workspace_bytes(const ntk_program_t* prog, int32_t dtype, int32_t t1, int32_t t2, int32_t H,
                        int32_t W, int32_t C, uint32_t flags, size_t* bytes);

/* ---- caller-stream entry (what an XLA custom call binds) -----------------------------------------
 * Same as ntk_gram_device, but every kernel and copy is enqueued on `cuda_stream` (a cudaStream_t of the
 * context's device, e.g. XLA's compute stream from ffi::PlatformStream) and the call returns without
 * synchronising; results are valid once work enqueued on that stream before the call's return has run.
 * The context's workspace is reused in stream order: when consecutive calls use different streams the new
 * stream first waits (cudaStreamWaitEvent, on the device) for the previous call's work.
 * Replaces the dispatch of the jitted `kernel_fn` executable, `_src/stax/requirements.py:939-953`,
 * `_src/batching.py:725,760-776`. */
int ntk_gram_device_on_stream(ntk_context_t* ctx, const ntk_program_t* prog, int32_t dtype, const void* x1,
                              int32_t n1, const void* x2, int32_t n2, int32_t H, int32_t W, int32_t C,
                              uint32_t flags, void* nngp, void* ntk, int64_t ld, void* cov1, void* cov2,
                              void* cuda_stream);

/* Kernel-in / Kernel-out on DEVICE pointers, asynchronous on `cuda_stream` (NULL = the context stream):
 * composition `kernel_fn(kernel_fn_a(x1, x2))` without the host round trip of ntk_apply_host
 * (`requirements.py:935-937`, `_src/utils/kernel.py:151-167`). */
int ntk_apply_device(ntk_context_t* ctx, const ntk_program_t* prog, int32_t dtype, const ntk_state_t* in,
                     ntk_state_t* out, void* cuda_stream);

/* ---- multi-GPU: one process (or host thread) per GPU, NCCL over NVLink / NVSwitch -----------------
 * The reference's device parallelism is `pmap` over x1 rows with x2 replicated and no collective
 * (`_src/batching.py:505-644`).  Here each rank owns row blocks of the Gram matrix; x1 / x2 are broadcast
 * once, result slabs are all-gathered, and there is no reduction.  libnccl is loaded with dlopen on the
 * first ntk_comm_* call, so the library itself has no link-time NCCL dependency.
 * All collectives are enqueued on the stream of the context the communicator was created with. */
#define NTK_COMM_ID_BYTES 128
typedef struct ntk_comm ntk_comm_t;
/* rank 0 creates the id; the launcher hands it to every rank (file / env / TCP -- not the library's business) */
int ntk_comm_unique_id(void* id_out /* NTK_COMM_ID_BYTES */);
int ntk_comm_create(ntk_context_t* ctx, const void* id, int32_t rank, int32_t world, ntk_comm_t** out);
void ntk_comm_destroy(ntk_comm_t* comm);
int ntk_comm_rank(const ntk_comm_t* comm);
int ntk_comm_world(const ntk_comm_t* comm);
int ntk_comm_nccl_version(int* version);
int ntk_comm_broadcast(ntk_comm_t* comm, void* dev_buf, size_t bytes, int32_t root);
int ntk_comm_all_gather(ntk_comm_t* comm, const void* send_dev, void* recv_dev, size_t bytes_per_rank);

/* Symmetric Gram assembled from gathered upper-trapezoid row slabs (NTK_FLAG_UPPER_ONLY):
 *   out[i, j] = slabs[row_of[i], j]  (j >= i),   slabs[row_of[j], i]  (j < i)
 * `slabs` is [rows, ld_slabs], `row_of` is int32[n] on the device, `out` is [n, ld_out]. */
int ntk_sym_assemble(ntk_context_t* ctx, int32_t dtype, const void* slabs, int64_t ld_slabs,
                     const int32_t* row_of, int32_t n, void* out, int64_t ld_out);

/* ---- the caller of the Gram path: closed-form inference on device-resident matrices ----------------
 * `predict.gp_inference` / `gradient_descent_mse_ensemble(t=None)` (`_src/predict.py:566-750,753-1100`) solve
 * (K_dd + reg I) alpha = y and return K_td alpha.  These entry points do that where the Gram kernels left their
 * results, in HBM, in float64: `ntk_chol_factor` copies / widens a symmetric [n, n] matrix of `dtype`, adds
 * diag_reg * (absolute ? 1 : trace(K) / n) to the diagonal (`_add_diagonal_regularizer`, `_src/predict.py:1186-1214`)
 * and factorises it (blocked Cholesky on the fp64 tensor cores; `_get_cho_solve`, `:1217-1240`).
 * All work is enqueued on the context stream; only ntk_chol_info synchronises. */
typedef struct ntk_chol ntk_chol_t;
int ntk_chol_factor(ntk_context_t* ctx, int32_t dtype, const void* k_dev, int32_t n, int64_t ld, double diag_reg,
                    int32_t absolute, ntk_chol_t** out);
/* info = 0: success; info = k > 0: the leading minor of order k is not positive definite (LAPACK convention). */
int ntk_chol_info(ntk_context_t* ctx, const ntk_chol_t* f, int32_t* info);
/* B <- (K + reg I)^-1 B for B [n, nrhs] float64 row-major (ldb elements); work_dev holds n * nrhs doubles. */
int ntk_chol_solve(ntk_context_t* ctx, const ntk_chol_t* f, double* b_dev, int32_t nrhs, int64_t ldb,
                   double* work_dev);
/* out[m, nrhs] (float64) = A[m, n] (dtype_a) X[n, nrhs] (float64): K_test_train alpha. */
int ntk_matmul_f64(ntk_context_t* ctx, int32_t dtype_a, const void* a_dev, int32_t m, int32_t n, int64_t lda,
                   const double* x_dev, int32_t nrhs, int64_t ldx, double* out_dev, int64_t ldo);
/* Lower Cholesky factor, [n, n] float64 row-major on the device (strict upper triangle zero). */
const double* ntk_chol_factor_ptr(const ntk_chol_t* f);
void ntk_chol_destroy(ntk_chol_t* f);

/* Finite training times work in the eigenbasis of the regularised train-train matrix (`_get_fns_in_eigenbasis`,
 * `_src/predict.py:1243-1290`: `np.linalg.eigh`).  `ntk_eigh_compute` widens a symmetric [n, n] device matrix of `dtype`
 * to float64, adds diag_reg * (absolute ? 1 : trace(K) / n) to the diagonal and diagonalises it on the device with a
 * parallel cyclic Jacobi method (one pass over the matrix per round of n / 2 disjoint rotations); it stops when the
 * off-diagonal Frobenius mass is <= tol * ||A||_F (tol <= 0: 1e-15) or after max_sweeps (<= 0: 40) sweeps.
 * Eigenvalues ascending (NumPy order); V is [n, n] float64 row-major with eigenvector k in COLUMN k, Vt its transpose;
 * both stay on the device for `ntk_matmul_f64`.  Synchronises the context stream once per sweep. */
typedef struct ntk_eigh ntk_eigh_t;
int ntk_eigh_compute(ntk_context_t* ctx, int32_t dtype, const void* k_dev, int32_t n, int64_t ld, double diag_reg,
                     int32_t absolute, int32_t max_sweeps, double tol, ntk_eigh_t** out);
int ntk_eigh_info(const ntk_eigh_t* e, int32_t* sweeps, double* off_over_norm);
int ntk_eigh_values(const ntk_eigh_t* e, double* w_host);        /* n doubles, host */
const double* ntk_eigh_values_ptr(const ntk_eigh_t* e);          /* n doubles, device */
const double* ntk_eigh_vectors_ptr(const ntk_eigh_t* e);         /* V  [n, n], device */
const double* ntk_eigh_vectors_t_ptr(const ntk_eigh_t* e);       /* V^T [n, n], device */
void ntk_eigh_destroy(ntk_eigh_t* e);

/* ---- device memory helpers (so a host language needs no CUDA binding) ---- */
int ntk_device_malloc(int32_t device, size_t bytes, void** ptr);
int ntk_device_free(int32_t device, void* ptr);
int ntk_memcpy_h2d(ntk_context_t* ctx, void* dst_dev, const void* src_host, size_t bytes);
int ntk_memcpy_d2h(ntk_context_t* ctx, void* dst_host, const void* src_dev, size_t bytes);
int ntk_memset_async(ntk_context_t* ctx, void* dst_dev, int32_t value, size_t bytes);
int ntk_context_device(const ntk_context_t* ctx);
/* Page-locked host memory: ntk_gram_host moves pageable caller arrays through the context's pinned
 * staging ring; arrays that already live in pinned memory (these) are DMA'd directly. */
int ntk_host_alloc(size_t bytes, void** ptr);
int ntk_host_free(void* ptr);
/* CUDA events on the context stream (device-side timing for a host language without a CUDA binding). */
int ntk_event_create(void** event);
int ntk_event_record(ntk_context_t* ctx, void* event);
int ntk_event_elapsed_ms(void* start, void* stop, float* ms); /* waits for `stop` */
void ntk_event_destroy(void* event);
/* A second stream on the context's device (tests of the caller-stream entry; bench timing streams). */
int ntk_stream_create(int32_t device, void** cuda_stream);
int ntk_stream_synchronize(void* cuda_stream);
int ntk_stream_query(void* cuda_stream, int32_t* done); /* done = 1 when all enqueued work has finished */
void ntk_stream_destroy(void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* NTK_B200_H_ */
