// Input-layer Gram for [N, d] inputs on the tensor cores:  out[i, j] = x1[i] . x2[j] / d
// (`_src/stax/requirements.py:585-638`: `_cov` for 2-D inputs, `jnp.tensordot` + `/C`).
//
//  fp32: tcgen05.mma kind::tf32 with a 3xTF32 split (hi*hi + hi*lo + lo*hi) so the result is
//        FP32-exact to ~2^-21 relative; operands are staged by the CTA into shared memory in the
//        canonical no-swizzle K-major core-matrix layout, the 128 x 128 accumulator lives in TMEM
//        and is read back with tcgen05.ld for the 1/d epilogue.
//  fp64: legacy DMMA (mma.sync.m8n8k4.f64) -- tcgen05 has no f64 kind.
//
// Both kernels zero-pad ragged M / N / K edges, so any [n1, d] x [n2, d] works.
#pragma once

#include <cstdint>

#include "common.cuh"

namespace ntk {

// ------------------------------------------------------------------------------------------
// tcgen05 helpers (PTX ISA 8.7+, sm_100a)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Shared-memory matrix descriptor, K-major, SWIZZLE_NONE ("interleave"): core matrices are
// 8 rows x 16 bytes, stored contiguously (128 B); LBO = byte distance between the two core
// matrices an instruction consumes along K, SBO = byte distance between 8-row groups.
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t saddr, uint32_t lbo_bytes,
                                                     uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

// Instruction descriptor for kind::tf32, FP32 accumulate, both operands K-major.
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate));
}

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

constexpr int kGemmBM = 128, kGemmBN = 128, kGemmBK = 32;

// One CTA (128 threads) per 128 x 128 output tile.
static __global__ void __launch_bounds__(128)
k_gram_tf32x3(const float* __restrict__ x1, const float* __restrict__ x2, float* __restrict__ out,
              int n1, int n2, int d, float inv_d) {
  // [hi/lo][K/4][rows/8][8][4]: address(r, k) = (k/4)*rows*16 + (r/8)*128 + (r%8)*16 + (k%4)*4
  extern __shared__ __align__(128) unsigned char gemm_smem[];
  float(*sA)[kGemmBM * kGemmBK] = reinterpret_cast<float(*)[kGemmBM * kGemmBK]>(gemm_smem);
  float(*sB)[kGemmBN * kGemmBK] =
      reinterpret_cast<float(*)[kGemmBN * kGemmBK]>(gemm_smem + 2 * kGemmBM * kGemmBK * sizeof(float));
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int m0 = blockIdx.y * kGemmBM, n0 = blockIdx.x * kGemmBN;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_s)),
                 "r"((uint32_t)kGemmBN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_d = tmem_base_s;

  constexpr uint32_t LBO = kGemmBM * 16;  // next K core matrix (4 floats) of the same rows
  constexpr uint32_t SBO = 128;           // next group of 8 rows
  constexpr uint32_t idesc = umma_idesc_tf32(kGemmBM, kGemmBN);
  uint32_t phase = 0;
  uint32_t acc = 0;

  for (int k0 = 0; k0 < d; k0 += kGemmBK) {
    // ---- stage the operand tiles: thread t loads row (t), 32 consecutive k (coalesced per
    // 8-thread group via float4), splits into tf32 hi/lo and scatters into core matrices ----
#pragma unroll 2
    for (int it = 0; it < (kGemmBM * kGemmBK / 4) / 128; ++it) {
      const int v = it * 128 + tid;      // float4 index inside the tile
      const int r = v / (kGemmBK / 4);   // row
      const int kc = v % (kGemmBK / 4);  // k chunk of 4
      const int k = k0 + kc * 4;
      float a4[4] = {0.f, 0.f, 0.f, 0.f}, b4[4] = {0.f, 0.f, 0.f, 0.f};
      if (m0 + r < n1) {
        const float* g = x1 + (size_t)(m0 + r) * d + k;
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (k + e < d) a4[e] = g[e];
      }
      if (n0 + r < n2) {
        const float* g = x2 + (size_t)(n0 + r) * d + k;
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (k + e < d) b4[e] = g[e];
      }
      const int off = kc * (kGemmBM * 4) + (r >> 3) * 32 + (r & 7) * 4;  // in floats
      float4 ah, al, bh, bl;
      float* pah = &ah.x;
      float* pal = &al.x;
      float* pbh = &bh.x;
      float* pbl = &bl.x;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        pah[e] = to_tf32(a4[e]);
        pal[e] = to_tf32(a4[e] - pah[e]);
        pbh[e] = to_tf32(b4[e]);
        pbl[e] = to_tf32(b4[e] - pbh[e]);
      }
      *reinterpret_cast<float4*>(&sA[0][off]) = ah;
      *reinterpret_cast<float4*>(&sA[1][off]) = al;
      *reinterpret_cast<float4*>(&sB[0][off]) = bh;
      *reinterpret_cast<float4*>(&sB[1][off]) = bl;
    }
    asm volatile("fence.proxy.async.shared::cta;");  // generic-proxy writes -> async proxy
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;");
      const uint32_t a_hi = smem_u32(&sA[0][0]), a_lo = smem_u32(&sA[1][0]);
      const uint32_t b_hi = smem_u32(&sB[0][0]), b_lo = smem_u32(&sB[1][0]);
#pragma unroll
      for (int ks = 0; ks < kGemmBK / 8; ++ks) {  // one instruction = K 8 = two core matrices
        const uint32_t koff = ks * 2 * LBO;
        const uint64_t dah = umma_desc_kmajor(a_hi + koff, LBO, SBO);
        const uint64_t dal = umma_desc_kmajor(a_lo + koff, LBO, SBO);
        const uint64_t dbh = umma_desc_kmajor(b_hi + koff, LBO, SBO);
        const uint64_t dbl = umma_desc_kmajor(b_lo + koff, LBO, SBO);
        umma_tf32(tmem_d, dal, dbh, idesc, acc);  // small terms first
        umma_tf32(tmem_d, dah, dbl, idesc, 1u);
        umma_tf32(tmem_d, dah, dbh, idesc, 1u);
        acc = 1u;
      }
      // arrive on the mbarrier when every MMA issued so far has finished reading smem
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
          smem_u32(&mbar)));
    }
    // all threads wait for the MMAs before overwriting the tiles
    {
      uint32_t done = 0;
      while (!done) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(&mbar)), "r"(phase));
      }
      phase ^= 1;
    }
  }

  // ---- epilogue: TMEM -> registers -> global, scaled by 1/d -------------------------------
  asm volatile("tcgen05.fence::after_thread_sync;");
  const int row = m0 + warp * 32 + (tid & 31);  // warp w owns TMEM lanes [32w, 32w+32)
  const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
  for (int c0 = 0; c0 < kGemmBN; c0 += 8) {
    uint32_t v[8];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7])
        : "r"(taddr + (uint32_t)c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    if (row < n1) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int col = n0 + c0 + e;
        if (col < n2) out[(size_t)row * n2 + col] = __uint_as_float(v[e]) * inv_d;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d),
                 "r"((uint32_t)kGemmBN));
}

// ------------------------------------------------------------------------------------------
// Pipelined variant for the FCN path (executor.cu::fcn_gram).  The TF32 hi / lo split is done once
// by k_split_tf32 into two zero-padded [n, d_pad] arrays, so that the GEMM's operand tiles are plain
// 16-byte copies: every thread streams them with cp.async straight into the no-swizzle core-matrix
// layout (same descriptors as above), kStages deep, while one elected thread issues the three
// tcgen05.mma of each K slice and commits them to a per-stage mbarrier.  A stage is refilled as soon
// as the MMAs that read it have retired; the accumulator never leaves TMEM until the epilogue.
// ------------------------------------------------------------------------------------------
static __global__ void k_split_tf32(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo,
                                    long long n, int d, int d_pad) {
  const long long total = n * d_pad;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / d_pad;
    const int k = (int)(idx % d_pad);
    const float v = k < d ? x[r * d + k] : 0.f;
    const float h = to_tf32(v);
    hi[idx] = h;
    lo[idx] = to_tf32(v - h);
  }
}

constexpr int kGemmStages = 3;

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool valid) {
  const uint32_t bytes = valid ? 16u : 0u;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}

static __global__ void __launch_bounds__(128)
k_gram_tf32x3_pipe(const float* __restrict__ a_hi, const float* __restrict__ a_lo,
                   const float* __restrict__ b_hi, const float* __restrict__ b_lo, float* __restrict__ out,
                   int n1, int n2, int d_pad, long long ld_out, float inv_d) {
  // stage layout: [A_hi | A_lo | B_hi | B_lo], each a 128 x 32 tile in core-matrix order
  //   byte offset(r, kc) = kc * (128 * 16) + (r / 8) * 128 + (r % 8) * 16,  kc = k / 4
  extern __shared__ __align__(128) unsigned char gemm_smem[];
  constexpr uint32_t kTileBytes = kGemmBM * kGemmBK * sizeof(float);  // 16 KB
  constexpr uint32_t kStageBytes = 4 * kTileBytes;                    // 64 KB
  __shared__ __align__(8) uint64_t mbar[kGemmStages];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int m0 = blockIdx.y * kGemmBM, n0 = blockIdx.x * kGemmBN;
  const uint32_t smem0 = smem_u32(gemm_smem);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_s)),
                 "r"((uint32_t)kGemmBN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int s = 0; s < kGemmStages; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[s])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_d = tmem_base_s;

  constexpr uint32_t LBO = kGemmBM * 16;
  constexpr uint32_t SBO = 128;
  constexpr uint32_t idesc = umma_idesc_tf32(kGemmBM, kGemmBN);
  const int nk = d_pad / kGemmBK;

  // thread t copies, for each of the 4 tiles, chunks v = it*128 + t: row r = v / 8, k-chunk kc = v % 8
  auto load_stage = [&](int kt, int stage) {
    const uint32_t sbase = smem0 + (uint32_t)stage * kStageBytes;
    const int k0 = kt * kGemmBK;
#pragma unroll
    for (int it = 0; it < (kGemmBM * kGemmBK / 4) / 128; ++it) {
      const int v = it * 128 + tid;
      const int r = v >> 3, kc = v & 7;
      const uint32_t off = (uint32_t)kc * (kGemmBM * 16) + (uint32_t)(r >> 3) * 128 + (uint32_t)(r & 7) * 16;
      const bool va = m0 + r < n1, vb = n0 + r < n2;
      const size_t ga = (size_t)(va ? m0 + r : 0) * d_pad + k0 + kc * 4;
      const size_t gb = (size_t)(vb ? n0 + r : 0) * d_pad + k0 + kc * 4;
      cp_async16_zfill(sbase + off, a_hi + ga, va);
      cp_async16_zfill(sbase + kTileBytes + off, a_lo + ga, va);
      cp_async16_zfill(sbase + 2 * kTileBytes + off, b_hi + gb, vb);
      cp_async16_zfill(sbase + 3 * kTileBytes + off, b_lo + gb, vb);
    }
  };

  // prologue: kStages - 1 slices in flight
  for (int s = 0; s < kGemmStages - 1; ++s) {
    if (s < nk) load_stage(s, s);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  uint32_t acc = 0;
  for (int kt = 0; kt < nk; ++kt) {
    const int stage = kt % kGemmStages;
    asm volatile("cp.async.wait_group %0;" ::"n"(kGemmStages - 2) : "memory");  // slice kt has landed
    asm volatile("fence.proxy.async.shared::cta;");                                // generic -> async proxy
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;");
      const uint32_t sbase = smem0 + (uint32_t)stage * kStageBytes;
      const uint32_t ahi = sbase, alo = sbase + kTileBytes, bhi = sbase + 2 * kTileBytes, blo = sbase + 3 * kTileBytes;
#pragma unroll
      for (int ks = 0; ks < kGemmBK / 8; ++ks) {
        const uint32_t koff = ks * 2 * LBO;
        const uint64_t dah = umma_desc_kmajor(ahi + koff, LBO, SBO);
        const uint64_t dal = umma_desc_kmajor(alo + koff, LBO, SBO);
        const uint64_t dbh = umma_desc_kmajor(bhi + koff, LBO, SBO);
        const uint64_t dbl = umma_desc_kmajor(blo + koff, LBO, SBO);
        umma_tf32(tmem_d, dal, dbh, idesc, acc);  // small terms first
        umma_tf32(tmem_d, dah, dbl, idesc, 1u);
        umma_tf32(tmem_d, dah, dbh, idesc, 1u);
        acc = 1u;
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
          smem_u32(&mbar[stage])));
    }
    // refill the stage read by slice kt - 1 (its MMAs were committed one iteration ago) with slice kt + kStages - 1
    const int nxt = kt + kGemmStages - 1;
    if (kt >= 1) {
      const int ps = (kt - 1) % kGemmStages;
      mbar_wait(smem_u32(&mbar[ps]), (uint32_t)(((kt - 1) / kGemmStages) & 1));
    }
    if (nxt < nk) load_stage(nxt, nxt % kGemmStages);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  // the last slice's MMAs
  {
    const int last = nk - 1;
    mbar_wait(smem_u32(&mbar[last % kGemmStages]), (uint32_t)((last / kGemmStages) & 1));
  }

  asm volatile("tcgen05.fence::after_thread_sync;");
  const int row = m0 + warp * 32 + (tid & 31);
  const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
  for (int c0 = 0; c0 < kGemmBN; c0 += 8) {
    uint32_t v[8];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7])
        : "r"(taddr + (uint32_t)c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    if (row < n1) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int col = n0 + c0 + e;
        if (col < n2) out[(size_t)row * ld_out + col] = __uint_as_float(v[e]) * inv_d;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d),
                 "r"((uint32_t)kGemmBN));
}

inline int gram_pad_k(int d) { return (d + kGemmBK - 1) / kGemmBK * kGemmBK; }

// TMA-fed, warp-specialised variant of k_gram_tf32x3_pipe (gemm_tma.cu: its own translation unit)
int launch_gram_tc_tma(cudaStream_t stream, const float* a_hi, const float* a_lo, int n1, const float* b_hi,
                       const float* b_lo, int n2, int d, float* out, long long ld_out);

// hi/lo: [n, gram_pad_k(d)] each
inline int launch_split_tf32(cudaStream_t stream, const float* x, long long n, int d, float* hi, float* lo) {
  const int d_pad = gram_pad_k(d);
  const long long total = n * d_pad;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  k_split_tf32<<<(unsigned)blocks, 256, 0, stream>>>(x, hi, lo, n, d, d_pad);
  NTK_CUDA(cudaGetLastError());
  return NTK_OK;
}

inline int launch_gram_tc_pipe(cudaStream_t stream, const float* a_hi, const float* a_lo, int n1,
                               const float* b_hi, const float* b_lo, int n2, int d, float* out, long long ld_out) {
  dim3 grid((n2 + kGemmBN - 1) / kGemmBN, (n1 + kGemmBM - 1) / kGemmBM);
  constexpr size_t smem = (size_t)kGemmStages * 4 * kGemmBM * kGemmBK * sizeof(float);  // 192 KB
  NTK_TRY(ensure_dynamic_smem((const void*)k_gram_tf32x3_pipe, smem));
  k_gram_tf32x3_pipe<<<grid, 128, smem, stream>>>(a_hi, a_lo, b_hi, b_lo, out, n1, n2, gram_pad_k(d), ld_out,
                                                  (float)(1.0 / (double)d));
  NTK_CUDA(cudaGetLastError());
  return NTK_OK;
}

// ------------------------------------------------------------------------------------------
// fp64: DMMA m8n8k4.  CTA = 8 warps -> 32 x 64 output tile; warp (wr, wc) owns rows
// [8 wr, 8 wr + 8) x cols [32 wc, 32 wc + 32) as four 8x8 accumulators.
// ------------------------------------------------------------------------------------------
constexpr int kDmmaBM = 32, kDmmaBN = 64, kDmmaBK = 16;

static __global__ void __launch_bounds__(256)
k_gram_dmma(const double* __restrict__ x1, const double* __restrict__ x2, double* __restrict__ out,
            int n1, int n2, int d, double inv_d) {
  __shared__ double sA[kDmmaBM][kDmmaBK + 1];
  __shared__ double sB[kDmmaBN][kDmmaBK + 1];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wr = warp >> 1, wc = warp & 1;
  const int m0 = blockIdx.y * kDmmaBM, n0 = blockIdx.x * kDmmaBN;
  double c[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
  for (int k0 = 0; k0 < d; k0 += kDmmaBK) {
    for (int v = tid; v < kDmmaBM * kDmmaBK; v += 256) {
      const int r = v / kDmmaBK, k = v % kDmmaBK;
      sA[r][k] = (m0 + r < n1 && k0 + k < d) ? x1[(size_t)(m0 + r) * d + k0 + k] : 0.0;
    }
    for (int v = tid; v < kDmmaBN * kDmmaBK; v += 256) {
      const int r = v / kDmmaBK, k = v % kDmmaBK;
      sB[r][k] = (n0 + r < n2 && k0 + k < d) ? x2[(size_t)(n0 + r) * d + k0 + k] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < kDmmaBK; ks += 4) {
      const double a = sA[wr * 8 + (lane >> 2)][ks + (lane & 3)];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const double b = sB[wc * 32 + nt * 8 + (lane >> 2)][ks + (lane & 3)];
        asm volatile(
            "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
            : "+d"(c[nt][0]), "+d"(c[nt][1])
            : "d"(a), "d"(b));
      }
    }
    __syncthreads();
  }
  const int row = m0 + wr * 8 + (lane >> 2);
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int col = n0 + wc * 32 + nt * 8 + 2 * (lane & 3);
    if (row < n1) {
      if (col < n2) out[(size_t)row * n2 + col] = c[nt][0] * inv_d;
      if (col + 1 < n2) out[(size_t)row * n2 + col + 1] = c[nt][1] * inv_d;
    }
  }
}

inline int launch_gram_tc(cudaStream_t stream, const float* x1, int n1, const float* x2, int n2, int d,
                          float* out) {
  dim3 grid((n2 + kGemmBN - 1) / kGemmBN, (n1 + kGemmBM - 1) / kGemmBM);
  constexpr size_t smem = 2 * (kGemmBM + kGemmBN) * kGemmBK * sizeof(float);  // 64 KB
  NTK_TRY(ensure_dynamic_smem((const void*)k_gram_tf32x3, smem));
  k_gram_tf32x3<<<grid, 128, smem, stream>>>(x1, x2, out, n1, n2, d, (float)(1.0 / (double)d));
  NTK_CUDA(cudaGetLastError());
  return NTK_OK;
}

inline int launch_gram_tc(cudaStream_t stream, const double* x1, int n1, const double* x2, int n2, int d,
                          double* out) {
  dim3 grid((n2 + kDmmaBN - 1) / kDmmaBN, (n1 + kDmmaBM - 1) / kDmmaBM);
  k_gram_dmma<<<grid, 256, 0, stream>>>(x1, x2, out, n1, n2, d, 1.0 / (double)d);
  NTK_CUDA(cudaGetLastError());
  return NTK_OK;
}

}  // namespace ntk
