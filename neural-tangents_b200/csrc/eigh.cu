// Symmetric eigendecomposition of the [n, n] Gram matrices where the Gram kernels leave them (HBM), in float64:
// the eigenbasis `predict.gradient_descent_mse` / `gradient_descent_mse_ensemble` need for finite training times
// (`_src/predict.py:1243-1290` `_get_fns_in_eigenbasis`: `np.linalg.eigh` of the regularised train-train matrix).
//
// Parallel cyclic two-sided Jacobi.  A sweep is N - 1 rounds of the round-robin tournament on N = n (+1 if n is odd)
// indices; a round rotates its N / 2 disjoint pairs (p, q) at once:
//   k_jacobi_rot    one thread per pair: (c, s) that annihilate A[p, q]  (Golub & Van Loan, symmetric Schur 2 x 2)
//   k_jacobi_apply  one thread per 2 x 2 block A[{p_i, q_i}, {p_j, q_j}] <- J_i^T A_blk J_j   (in place: one pass over A)
//   k_jacobi_vt     W <- J^T W on the rows of W = V^T (contiguous along the row: coalesced)
// The pairs of a round are (r + k, r - k) mod (N - 1): consecutive threads touch ascending p_j and descending q_j, so
// the block accesses stay sector-dense without moving data between rounds.  Every step is HBM / L2-bound elementwise
// work (32 n^2 bytes per round); convergence (off-diagonal Frobenius mass) is tested once per sweep.
#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

#include "common.cuh"

namespace ntk {
namespace {

struct JPair {
  int p, q;     // p < q; q >= n marks the bye of an odd n (identity)
  double c, s;
};

__device__ __forceinline__ void round_pair(int k, int r, int N, int& p, int& q) {
  int a, b;
  if (k == 0) {
    a = N - 1;
    b = r;
  } else {
    a = (r + k) % (N - 1);
    b = (r - k + (N - 1)) % (N - 1);
  }
  p = a < b ? a : b;
  q = a < b ? b : a;
}

__global__ void k_jacobi_rot(const double* __restrict__ A, int n, int N, int r, JPair* __restrict__ pairs) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= N / 2) return;
  JPair jp;
  round_pair(k, r, N, jp.p, jp.q);
  jp.c = 1.0;
  jp.s = 0.0;
  if (jp.q < n) {
    const double apq = A[(long long)jp.p * n + jp.q];
    if (apq != 0.0) {
      const double app = A[(long long)jp.p * n + jp.p], aqq = A[(long long)jp.q * n + jp.q];
      const double tau = (aqq - app) / (2.0 * apq);
      const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
      jp.c = 1.0 / sqrt(1.0 + t * t);
      jp.s = t * jp.c;
    }
  }
  pairs[k] = jp;
}

// J = [[c, s], [-s, c]] on (p, q);  A <- J_i^T A J_j block by block
__global__ void k_jacobi_apply(double* __restrict__ A, int n, int npairs, const JPair* __restrict__ pairs) {
  const int bj = blockIdx.x * blockDim.x + threadIdx.x;
  const int bi = blockIdx.y * blockDim.y + threadIdx.y;
  if (bi >= npairs || bj >= npairs) return;
  const JPair ri = pairs[bi], cj = pairs[bj];
  const bool hq_i = ri.q < n, hq_j = cj.q < n;
  double* r0 = A + (long long)ri.p * n;
  double* r1 = A + (long long)(hq_i ? ri.q : ri.p) * n;
  const double a00 = r0[cj.p], a01 = hq_j ? r0[cj.q] : 0.0;
  const double a10 = hq_i ? r1[cj.p] : 0.0, a11 = (hq_i && hq_j) ? r1[cj.q] : 0.0;
  const double b00 = ri.c * a00 - ri.s * a10, b01 = ri.c * a01 - ri.s * a11;
  const double b10 = ri.s * a00 + ri.c * a10, b11 = ri.s * a01 + ri.c * a11;
  double c00 = b00 * cj.c - b01 * cj.s, c01 = b00 * cj.s + b01 * cj.c;
  double c10 = b10 * cj.c - b11 * cj.s, c11 = b10 * cj.s + b11 * cj.c;
  if (bi == bj) c01 = c10 = 0.0;  // the annihilated entry, exactly
  r0[cj.p] = c00;
  if (hq_j) r0[cj.q] = c01;
  if (hq_i) {
    r1[cj.p] = c10;
    if (hq_j) r1[cj.q] = c11;
  }
}

// W = V^T: rows p, q <- J^T rows
__global__ void k_jacobi_vt(double* __restrict__ W, int n, int npairs, const JPair* __restrict__ pairs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (i >= n || b >= npairs) return;
  const JPair jp = pairs[b];
  if (jp.q >= n) return;
  double* wp = W + (long long)jp.p * n;
  double* wq = W + (long long)jp.q * n;
  const double v0 = wp[i], v1 = wq[i];
  wp[i] = jp.c * v0 - jp.s * v1;
  wq[i] = jp.s * v0 + jp.c * v1;
}

// A = (double) K + reg I  (full square);  W = I
template <typename T>
__global__ void k_eigh_init(const T* __restrict__ k, long long ld, int n, double diag_reg, int absolute,
                            const double* __restrict__ trace, double* __restrict__ A, double* __restrict__ W) {
  const double reg = diag_reg * (absolute ? 1.0 : *trace / (double)n);
  const long long total = (long long)n * n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx / n), j = (int)(idx % n);
    A[idx] = (double)k[(long long)i * ld + j] + (i == j ? reg : 0.0);
    W[idx] = i == j ? 1.0 : 0.0;
  }
}

template <typename T>
__global__ void k_eigh_trace(const T* __restrict__ a, int n, long long ld, double* __restrict__ out) {
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

// out[0] += sum of squares of the off-diagonal entries, out[1] += sum of squares of the diagonal (fixed-order per block,
// atomics across blocks: a convergence test, not a result)
__global__ void k_eigh_norms(const double* __restrict__ A, int n, double* __restrict__ out) {
  __shared__ double s_off[256], s_dia[256];
  double off = 0.0, dia = 0.0;
  const long long total = (long long)n * n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx / n), j = (int)(idx % n);
    const double v = A[idx];
    if (i == j)
      dia += v * v;
    else
      off += v * v;
  }
  s_off[threadIdx.x] = off;
  s_dia[threadIdx.x] = dia;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      s_off[threadIdx.x] += s_off[threadIdx.x + o];
      s_dia[threadIdx.x] += s_dia[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    atomicAdd(out, s_off[0]);
    atomicAdd(out + 1, s_dia[0]);
  }
}

__global__ void k_eigh_diag(const double* __restrict__ A, int n, double* __restrict__ w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) w[i] = A[(long long)i * n + i];
}

// sorted outputs: Vt[k, :] = W[perm[k], :],  V[:, k] = W[perm[k], :]^T
__global__ void k_eigh_gather(const double* __restrict__ W, const int* __restrict__ perm, int n, double* __restrict__ V,
                              double* __restrict__ Vt) {
  __shared__ double tile[32][33];
  const int k0 = blockIdx.y * 32, i0 = blockIdx.x * 32;
  for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
    const int k = k0 + dy, i = i0 + threadIdx.x;
    double v = 0.0;
    if (k < n && i < n) {
      v = W[(long long)perm[k] * n + i];
      Vt[(long long)k * n + i] = v;
    }
    tile[dy][threadIdx.x] = v;
  }
  __syncthreads();
  for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
    const int i = i0 + dy, k = k0 + threadIdx.x;
    if (k < n && i < n) V[(long long)i * n + k] = tile[threadIdx.x][dy];
  }
}

}  // namespace
}  // namespace ntk

using namespace ntk;

struct ntk_eigh {
  int device = 0;
  int n = 0;
  int sweeps = 0;
  double off_over_norm = 0.0;
  double* A = nullptr;     // work matrix -> diagonal
  double* W = nullptr;     // V^T (unsorted)
  double* V = nullptr;     // [n, n] row-major, column k = eigenvector of the k-th smallest eigenvalue
  double* Vt = nullptr;    // its transpose
  double* w = nullptr;     // [n] ascending (device)
  std::vector<double> w_host;
};

extern "C" {

void ntk_eigh_destroy(ntk_eigh_t* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  for (double* p : {e->A, e->W, e->V, e->Vt, e->w})
    if (p) cudaFree(p);
  delete e;
}

int ntk_eigh_compute(ntk_context_t* ctx, int32_t dtype, const void* k_dev, int32_t n, int64_t ld, double diag_reg,
                     int32_t absolute, int32_t max_sweeps, double tol, ntk_eigh_t** out) {
  if (!ctx || !k_dev || !out || n <= 0 || ld < n) return fail(NTK_EINVAL, "bad arguments");
  if (dtype != NTK_F32 && dtype != NTK_F64) return fail(NTK_EINVAL, "unknown dtype %d", dtype);
  if (max_sweeps <= 0) max_sweeps = 40;
  if (!(tol > 0.0)) tol = 1e-15;
  const int dev = ntk_context_device(ctx);
  NTK_CUDA(cudaSetDevice(dev));
  cudaStream_t s = (cudaStream_t)ntk_context_stream(ctx);
  ntk_eigh* e = new ntk_eigh();
  e->device = dev;
  e->n = n;
  const size_t nn = (size_t)n * n * sizeof(double);
  JPair* pairs = nullptr;
  double* scal = nullptr;
  int* perm_d = nullptr;
  cudaError_t err = cudaMalloc((void**)&e->A, nn);
  if (err == cudaSuccess) err = cudaMalloc((void**)&e->W, nn);
  if (err == cudaSuccess) err = cudaMalloc((void**)&e->w, (size_t)n * sizeof(double));
  const int N = n + (n & 1), npairs = N / 2;
  if (err == cudaSuccess) err = cudaMalloc((void**)&pairs, (size_t)std::max(npairs, 1) * sizeof(JPair));
  if (err == cudaSuccess) err = cudaMalloc((void**)&scal, 256);
  if (err == cudaSuccess) err = cudaMalloc((void**)&perm_d, (size_t)n * sizeof(int));
  auto cleanup = [&]() {
    if (pairs) cudaFree(pairs);
    if (scal) cudaFree(scal);
    if (perm_d) cudaFree(perm_d);
  };
  if (err != cudaSuccess) {
    cleanup();
    ntk_eigh_destroy(e);
    return fail(NTK_ENOMEM, "cudaMalloc for the %d x %d eigenproblem -> %s", n, n, cudaGetErrorString(err));
  }
  const int g1 = (int)std::min<long long>(((long long)n * n + 255) / 256, (long long)kNumSMs * 16);
  if (dtype == NTK_F32) {
    k_eigh_trace<float><<<1, 256, 0, s>>>((const float*)k_dev, n, ld, scal + 8);
    k_eigh_init<float><<<g1, 256, 0, s>>>((const float*)k_dev, ld, n, diag_reg, absolute, scal + 8, e->A, e->W);
  } else {
    k_eigh_trace<double><<<1, 256, 0, s>>>((const double*)k_dev, n, ld, scal + 8);
    k_eigh_init<double><<<g1, 256, 0, s>>>((const double*)k_dev, ld, n, diag_reg, absolute, scal + 8, e->A, e->W);
  }
  int rc = NTK_OK;
  double norms[2] = {0.0, 0.0};
  auto measure = [&]() -> int {
    NTK_CUDA(cudaMemsetAsync(scal, 0, 16, s));
    k_eigh_norms<<<g1, 256, 0, s>>>(e->A, n, scal);
    NTK_CUDA(cudaMemcpyAsync(norms, scal, 16, cudaMemcpyDeviceToHost, s));
    NTK_CUDA(cudaStreamSynchronize(s));
    return NTK_OK;
  };
  auto converged = [&]() { return norms[0] <= tol * tol * (norms[0] + norms[1]); };
  rc = measure();
  const dim3 blkA(32, 8), grdA((npairs + 31) / 32, (npairs + 7) / 8);
  const dim3 grdV((n + 255) / 256, npairs);
  while (rc == NTK_OK && n > 1 && !converged() && e->sweeps < max_sweeps) {
    for (int r = 0; r < N - 1; ++r) {
      k_jacobi_rot<<<(npairs + 127) / 128, 128, 0, s>>>(e->A, n, N, r, pairs);
      k_jacobi_apply<<<grdA, blkA, 0, s>>>(e->A, n, npairs, pairs);
      k_jacobi_vt<<<grdV, 256, 0, s>>>(e->W, n, npairs, pairs);
    }
    if (cudaGetLastError() != cudaSuccess) {
      rc = fail(NTK_ECUDA, "Jacobi sweep launch failed");
      break;
    }
    ++e->sweeps;
    rc = measure();
  }
  if (rc == NTK_OK) {
    e->off_over_norm = std::sqrt(norms[0] / std::max(norms[0] + norms[1], 1e-300));
    // eigenvalues = diagonal; sort ascending (np.linalg.eigh order) and gather the vectors
    std::vector<double> wd(n);
    k_eigh_diag<<<(n + 255) / 256, 256, 0, s>>>(e->A, n, e->w);
    cudaMemcpyAsync(wd.data(), e->w, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    std::vector<int> perm(n);
    std::iota(perm.begin(), perm.end(), 0);
    std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return wd[a] < wd[b]; });
    e->w_host.resize(n);
    for (int k = 0; k < n; ++k) e->w_host[k] = wd[perm[k]];
    cudaMemcpyAsync(perm_d, perm.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(e->w, e->w_host.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s);
    // the work matrix is dead: reuse it for V; Vt needs its own buffer
    err = cudaMalloc((void**)&e->Vt, nn);
    if (err != cudaSuccess) {
      rc = fail(NTK_ENOMEM, "cudaMalloc for the eigenvectors -> %s", cudaGetErrorString(err));
    } else {
      e->V = e->A;
      e->A = nullptr;
      k_eigh_gather<<<dim3((n + 31) / 32, (n + 31) / 32), dim3(32, 8), 0, s>>>(e->W, perm_d, n, e->V, e->Vt);
      if (cudaStreamSynchronize(s) != cudaSuccess || cudaGetLastError() != cudaSuccess)
        rc = fail(NTK_ECUDA, "eigenvector gather failed");
      cudaFree(e->W);
      e->W = nullptr;
    }
  }
  cleanup();
  if (rc != NTK_OK) {
    ntk_eigh_destroy(e);
    return rc;
  }
  *out = e;
  return NTK_OK;
}

int ntk_eigh_info(const ntk_eigh_t* e, int32_t* sweeps, double* off_over_norm) {
  if (!e) return fail(NTK_EINVAL, "bad arguments");
  if (sweeps) *sweeps = e->sweeps;
  if (off_over_norm) *off_over_norm = e->off_over_norm;
  return NTK_OK;
}

int ntk_eigh_values(const ntk_eigh_t* e, double* w_host) {
  if (!e || !w_host) return fail(NTK_EINVAL, "bad arguments");
  std::copy(e->w_host.begin(), e->w_host.end(), w_host);
  return NTK_OK;
}

const double* ntk_eigh_values_ptr(const ntk_eigh_t* e) { return e ? e->w : nullptr; }
const double* ntk_eigh_vectors_ptr(const ntk_eigh_t* e) { return e ? e->V : nullptr; }
const double* ntk_eigh_vectors_t_ptr(const ntk_eigh_t* e) { return e ? e->Vt : nullptr; }

}  // extern "C"
