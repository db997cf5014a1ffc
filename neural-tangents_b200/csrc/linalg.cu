// Dense fp64 linear algebra on the Gram matrices where they already are -- in HBM -- for the caller of the hot path,
// `predict.gp_inference` / `gradient_descent_mse_ensemble(t=None)` (`_src/predict.py:566-750, 753-1100`;
// SURVEY §8f row 1): the regularised train-train matrix is factorised next to the Gram kernels and only the
// [n_test, out] predictions cross PCIe.  Reference: `_add_diagonal_regularizer` (`_src/predict.py:1186-1214`) and
// `_get_cho_solve` (`:1217-1240`, `jax.scipy.linalg.cho_factor / cho_solve`).
//
// Blocked right-looking Cholesky, 64-wide block columns, row-major lower factor:
//   k_potrf_block   the 64 x 64 diagonal block in shared memory + its inverse (one CTA)
//   k_dgemm         every matrix-matrix step on the fp64 tensor cores (mma.sync.m8n8k4.f64, SASS DMMA):
//                   panel <- panel * inv(L_kk)^T, trailing -= panel * panel^T (lower tiles only), the two
//                   triangular sweeps of cho_solve and K_td * alpha
// FLOPs n^3 / 3; traffic ~ n^3 / 24 bytes (every trailing update is one read + write of the remaining lower triangle).
#include <algorithm>

#include "common.cuh"

using namespace ntk;

namespace {

constexpr int kNB = 64;  // block column width == dgemm tile
constexpr int kThreads = 256;
inline int grid_for(long long n) {
  long long b = (n + kThreads - 1) / kThreads;
  const long long cap = (long long)kNumSMs * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// C[M, N] = alpha * op(A) * op(B) + beta * C, row-major with leading dimensions, fp64 accumulate.
//   TRANS_A: A is stored [K, M] (op(A) = A^T), else [M, K]
//   B_KN:    B is stored [K, N], else [N, K] (the x1 x2^T form)
//   LOWER:   only tiles that intersect the lower triangle (n0 <= m0 + 63) are computed
// 64 x 64 tiles, 256 threads = 8 warps: warp (wr, wc) owns rows [16 wr, +16) x cols [32 wc, +32) as 2 x 4 m8n8k4 tiles.
template <typename TA, bool TRANS_A, bool B_KN, bool LOWER>
__global__ void __launch_bounds__(256)
k_dgemm(int M, int N, int K, double alpha, const TA* __restrict__ A, long long lda, const double* __restrict__ B,
        long long ldb, double beta, double* __restrict__ C, long long ldc) {
  constexpr int BK = 16;
  __shared__ double sA[kNB][BK + 1];
  __shared__ double sB[kNB][BK + 1];
  const int m0 = blockIdx.y * kNB, n0 = blockIdx.x * kNB;
  if (LOWER && n0 > m0 + kNB - 1) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wr = warp >> 1, wc = warp & 1;
  double c[2][4][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j][0] = c[i][j][1] = 0.0;
  for (int k0 = 0; k0 < K; k0 += BK) {
    for (int v = tid; v < kNB * BK; v += 256) {
      int r, k;
      if (TRANS_A) {
        k = v / kNB;
        r = v % kNB;
      } else {
        r = v / BK;
        k = v % BK;
      }
      double a = 0.0;
      if (m0 + r < M && k0 + k < K)
        a = (double)(TRANS_A ? A[(long long)(k0 + k) * lda + m0 + r] : A[(long long)(m0 + r) * lda + k0 + k]);
      sA[r][k] = a;
    }
    for (int v = tid; v < kNB * BK; v += 256) {
      int r, k;
      if (B_KN) {
        k = v / kNB;
        r = v % kNB;
      } else {
        r = v / BK;
        k = v % BK;
      }
      double b = 0.0;
      if (n0 + r < N && k0 + k < K)
        b = B_KN ? B[(long long)(k0 + k) * ldb + n0 + r] : B[(long long)(n0 + r) * ldb + k0 + k];
      sB[r][k] = b;
    }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < BK; ks += 4) {
      double a[2], b[4];
#pragma unroll
      for (int i = 0; i < 2; ++i) a[i] = sA[wr * 16 + i * 8 + (lane >> 2)][ks + (lane & 3)];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sB[wc * 32 + j * 8 + (lane >> 2)][ks + (lane & 3)];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                       : "+d"(c[i][j][0]), "+d"(c[i][j][1])
                       : "d"(a[i]), "d"(b[j]));
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int row = m0 + wr * 16 + i * 8 + (lane >> 2);
    if (row >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + wc * 32 + j * 8 + 2 * (lane & 3);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        if (col + e >= N) continue;
        if (LOWER && col + e > row) continue;   // the strict upper triangle is never written (stays zero)
        double* p = C + (long long)row * ldc + col + e;
        *p = beta == 0.0 ? alpha * c[i][j][e] : alpha * c[i][j][e] + beta * *p;
      }
    }
  }
}

template <typename TA, bool TRANS_A, bool B_KN, bool LOWER>
int dgemm(cudaStream_t s, int M, int N, int K, double alpha, const TA* A, long long lda, const double* B,
          long long ldb, double beta, double* C, long long ldc) {
  if (M <= 0 || N <= 0) return NTK_OK;
  dim3 grid((N + kNB - 1) / kNB, (M + kNB - 1) / kNB);
  k_dgemm<TA, TRANS_A, B_KN, LOWER><<<grid, 256, 0, s>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  NTK_CUDA(cudaGetLastError());
  return NTK_OK;
}

// trace of a [n, n] matrix -> *out (one CTA, fixed order)
template <typename T>
__global__ void k_trace(const T* __restrict__ a, int n, long long ld, double* __restrict__ out) {
  __shared__ double sm[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += (double)a[(long long)i * ld + i];
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sm[0];
}

// L[i, j] = (double) K[i, j] + (i == j) * reg,  reg = diag_reg * (absolute ? 1 : trace / n)   (lower triangle only)
template <typename T>
__global__ void k_regularize(const T* __restrict__ k, long long ld, int n, double diag_reg, int absolute,
                             const double* __restrict__ trace, double* __restrict__ L) {
  const double reg = diag_reg * (absolute ? 1.0 : *trace / (double)n);
  const long long total = (long long)n * n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx / n), j = (int)(idx % n);
    L[idx] = j > i ? 0.0 : (double)k[(long long)i * ld + j] + (i == j ? reg : 0.0);
  }
}

// Cholesky of the nb x nb diagonal block at (k0, k0) and the inverse of its factor.  64 threads, thread t owns row t.
__global__ void __launch_bounds__(kNB)
k_potrf_block(double* __restrict__ A, long long lda, int k0, int nb, double* __restrict__ Linv, int* __restrict__ info) {
  __shared__ double sL[kNB][kNB + 1];
  const int t = threadIdx.x;
  for (int v = t; v < kNB * kNB; v += kNB) {
    const int r = v / kNB, c = v % kNB;
    sL[r][c] = (r < nb && c <= r) ? A[(long long)(k0 + r) * lda + k0 + c] : (r == c ? 1.0 : 0.0);
  }
  __syncthreads();
  for (int j = 0; j < nb; ++j) {
    if (t == j) {
      double d = sL[j][j];
      if (!(d > 0.0)) {                       // not positive definite (or NaN): report the first failing minor
        atomicCAS(info, 0, k0 + j + 1);
        d = 1.0;
      }
      sL[j][j] = sqrt(d);
    }
    __syncthreads();
    if (t > j && t < nb) sL[t][j] /= sL[j][j];
    __syncthreads();
    if (t > j && t < nb) {
      const double ltj = sL[t][j];
      for (int c = j + 1; c <= t; ++c) sL[t][c] -= ltj * sL[c][j];
    }
    __syncthreads();
  }
  // write the factor back, then the inverse of the lower factor: thread t solves L x = e_t (column t of the inverse)
  for (int v = t; v < kNB * kNB; v += kNB) {
    const int r = v / kNB, c = v % kNB;
    if (r < nb && c <= r) A[(long long)(k0 + r) * lda + k0 + c] = sL[r][c];
  }
  double x[kNB];
#pragma unroll 1
  for (int i = 0; i < kNB; ++i) x[i] = 0.0;
  if (t < nb) {
    x[t] = 1.0 / sL[t][t];
    for (int i = t + 1; i < nb; ++i) {
      double acc = 0.0;
      for (int m = t; m < i; ++m) acc += sL[i][m] * x[m];
      x[i] = -acc / sL[i][i];
    }
  }
  for (int i = 0; i < kNB; ++i) Linv[i * kNB + t] = (i < nb && t < nb) ? x[i] : 0.0;
}

}  // namespace

struct ntk_chol {
  int device = 0;
  int n = 0;
  double* L = nullptr;     // [n, n] row-major, lower triangle valid
  double* Linv = nullptr;  // [n_blocks][64][64] inverses of the diagonal blocks
  double* scratch = nullptr;
  int* info = nullptr;
};

extern "C" {

void ntk_chol_destroy(ntk_chol_t* f) {
  if (!f) return;
  cudaSetDevice(f->device);
  if (f->L) cudaFree(f->L);
  if (f->Linv) cudaFree(f->Linv);
  if (f->scratch) cudaFree(f->scratch);
  if (f->info) cudaFree(f->info);
  delete f;
}

int ntk_chol_factor(ntk_context_t* ctx, int32_t dtype, const void* k_dev, int32_t n, int64_t ld, double diag_reg,
                    int32_t absolute, ntk_chol_t** out) {
  if (!ctx || !k_dev || !out || n <= 0 || ld < n) return fail(NTK_EINVAL, "bad arguments");
  if (dtype != NTK_F32 && dtype != NTK_F64) return fail(NTK_EINVAL, "unknown dtype %d", dtype);
  const int dev = ntk_context_device(ctx);
  NTK_CUDA(cudaSetDevice(dev));
  cudaStream_t s = (cudaStream_t)ntk_context_stream(ctx);
  ntk_chol* f = new ntk_chol();
  f->device = dev;
  f->n = n;
  const int nblk = (n + kNB - 1) / kNB;
  cudaError_t e = cudaMalloc((void**)&f->L, (size_t)n * n * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc((void**)&f->Linv, (size_t)nblk * kNB * kNB * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc((void**)&f->scratch, 256);
  if (e == cudaSuccess) e = cudaMalloc((void**)&f->info, 256);
  if (e != cudaSuccess) {
    ntk_chol_destroy(f);
    return fail(NTK_ENOMEM, "cudaMalloc for the %d x %d factor -> %s", n, n, cudaGetErrorString(e));
  }
  NTK_CUDA(cudaMemsetAsync(f->info, 0, 256, s));
  if (dtype == NTK_F32) {
    k_trace<float><<<1, 256, 0, s>>>((const float*)k_dev, n, ld, f->scratch);
    k_regularize<float><<<grid_for((long long)n * n), kThreads, 0, s>>>((const float*)k_dev, ld, n, diag_reg, absolute,
                                                                         f->scratch, f->L);
  } else {
    k_trace<double><<<1, 256, 0, s>>>((const double*)k_dev, n, ld, f->scratch);
    k_regularize<double><<<grid_for((long long)n * n), kThreads, 0, s>>>((const double*)k_dev, ld, n, diag_reg, absolute,
                                                                          f->scratch, f->L);
  }
  NTK_CUDA(cudaGetLastError());
  double* A = f->L;
  const long long lda = n;
  for (int k0 = 0, b = 0; k0 < n; k0 += kNB, ++b) {
    const int nb = std::min(kNB, n - k0);
    double* Li = f->Linv + (size_t)b * kNB * kNB;
    k_potrf_block<<<1, kNB, 0, s>>>(A, lda, k0, nb, Li, f->info);
    NTK_CUDA(cudaGetLastError());
    const int m = n - (k0 + nb);
    if (m <= 0) break;
    double* P = A + (long long)(k0 + nb) * lda + k0;                      // panel [m, nb]
    // panel <- panel * inv(L_kk)^T   (one 64-wide tile column: a CTA reads its rows completely before it writes them)
    NTK_TRY((dgemm<double, false, false, false>(s, m, nb, nb, 1.0, P, lda, Li, kNB, 0.0, P, lda)));
    // trailing -= panel * panel^T, lower tiles only
    double* T = A + (long long)(k0 + nb) * lda + (k0 + nb);
    NTK_TRY((dgemm<double, false, false, true>(s, m, m, nb, -1.0, P, lda, P, lda, 1.0, T, lda)));
  }
  *out = f;
  return NTK_OK;
}

int ntk_chol_info(ntk_context_t* ctx, const ntk_chol_t* f, int32_t* info) {
  if (!ctx || !f || !info) return fail(NTK_EINVAL, "bad arguments");
  NTK_CUDA(cudaSetDevice(f->device));
  cudaStream_t s = (cudaStream_t)ntk_context_stream(ctx);
  NTK_CUDA(cudaMemcpyAsync(info, f->info, sizeof(int), cudaMemcpyDeviceToHost, s));
  NTK_CUDA(cudaStreamSynchronize(s));
  return NTK_OK;
}

// B <- (L L^T)^-1 B for B [n, nrhs] fp64 row-major (ldb); `work` holds [n, nrhs] doubles.
int ntk_chol_solve(ntk_context_t* ctx, const ntk_chol_t* f, double* b_dev, int32_t nrhs, int64_t ldb, double* work_dev) {
  if (!ctx || !f || !b_dev || !work_dev || nrhs <= 0 || ldb < nrhs) return fail(NTK_EINVAL, "bad arguments");
  NTK_CUDA(cudaSetDevice(f->device));
  cudaStream_t s = (cudaStream_t)ntk_context_stream(ctx);
  const int n = f->n;
  const long long lda = n, ldx = nrhs;
  const double* L = f->L;
  double* X = work_dev;
  // forward sweep  L Y = B:   Y_k = inv(L_kk) B_k;   B_{k+1:} -= L_{k+1:, k} Y_k      (Y in X)
  for (int k0 = 0, b = 0; k0 < n; k0 += kNB, ++b) {
    const int nb = std::min(kNB, n - k0);
    const double* Li = f->Linv + (size_t)b * kNB * kNB;
    NTK_TRY((dgemm<double, false, true, false>(s, nb, nrhs, nb, 1.0, Li, kNB, b_dev + (long long)k0 * ldb, ldb, 0.0,
                                               X + (long long)k0 * ldx, ldx)));
    const int m = n - (k0 + nb);
    if (m > 0)
      NTK_TRY((dgemm<double, false, true, false>(s, m, nrhs, nb, -1.0, L + (long long)(k0 + nb) * lda + k0, lda,
                                                 X + (long long)k0 * ldx, ldx, 1.0, b_dev + (long long)(k0 + nb) * ldb,
                                                 ldb)));
  }
  // backward sweep  L^T Z = Y:   Z_k = inv(L_kk)^T Y_k;   Y_{:k} -= L_{k, :k}^T Z_k     (Z in B)
  const int nblk = (n + kNB - 1) / kNB;
  for (int b = nblk - 1; b >= 0; --b) {
    const int k0 = b * kNB, nb = std::min(kNB, n - k0);
    const double* Li = f->Linv + (size_t)b * kNB * kNB;
    NTK_TRY((dgemm<double, true, true, false>(s, nb, nrhs, nb, 1.0, Li, kNB, X + (long long)k0 * ldx, ldx, 0.0,
                                              b_dev + (long long)k0 * ldb, ldb)));
    if (k0 > 0)
      NTK_TRY((dgemm<double, true, true, false>(s, k0, nrhs, nb, -1.0, L + (long long)k0 * lda, lda,
                                                b_dev + (long long)k0 * ldb, ldb, 1.0, X, ldx)));
  }
  return NTK_OK;
}

// out[m, nrhs] (fp64) = A[m, n] (dtype) * X[n, nrhs] (fp64): K_test_train * (K_train_train^-1 y)
int ntk_matmul_f64(ntk_context_t* ctx, int32_t dtype_a, const void* a_dev, int32_t m, int32_t n, int64_t lda,
                   const double* x_dev, int32_t nrhs, int64_t ldx, double* out_dev, int64_t ldo) {
  if (!ctx || !a_dev || !x_dev || !out_dev || m <= 0 || n <= 0 || nrhs <= 0) return fail(NTK_EINVAL, "bad arguments");
  NTK_CUDA(cudaSetDevice(ntk_context_device(ctx)));
  cudaStream_t s = (cudaStream_t)ntk_context_stream(ctx);
  if (dtype_a == NTK_F32)
    return dgemm<float, false, true, false>(s, m, nrhs, n, 1.0, (const float*)a_dev, lda, x_dev, ldx, 0.0, out_dev, ldo);
  if (dtype_a == NTK_F64)
    return dgemm<double, false, true, false>(s, m, nrhs, n, 1.0, (const double*)a_dev, lda, x_dev, ldx, 0.0, out_dev,
                                             ldo);
  return fail(NTK_EINVAL, "unknown dtype %d", dtype_a);
}

// The lower Cholesky factor as a device pointer ([n, n] fp64 row-major; the strict upper triangle is zero).
const double* ntk_chol_factor_ptr(const ntk_chol_t* f) { return f ? f->L : nullptr; }

}  // extern "C"
