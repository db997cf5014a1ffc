// TMA-fed, warp-specialised input Gram for [N, d] inputs (round 2): see the comment in front of the kernel.
// Its own translation unit so that it can be iterated on without rebuilding the stage kernels.
#include "gemm_kernels.cuh"

#include <cuda.h>  // CUtensorMap types only; the encoder is fetched through the runtime (no -lcuda)

namespace ntk {
static int gemm_encode_tmap(void* map_out, const float* base, long long rows, int d_pad) {
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    NTK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(NTK_ECUDA, "cuTensorMapEncodeTiled is not available");
    encode = (encode_fn)fn;
  }
  static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap is 128 bytes");
  const cuuint64_t dims[2] = {(cuuint64_t)d_pad, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)d_pad * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)kGemmBK, (cuuint32_t)kGemmBM};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode((CUtensorMap*)map_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(NTK_ECUDA, "cuTensorMapEncodeTiled -> CUresult %d", (int)r);
  return NTK_OK;
}

// ------------------------------------------------------------------------------------------
// TMA-fed, warp-specialised variant (round 2).  Same math and operands as k_gram_tf32x3_pipe (pre-split TF32 hi / lo
// arrays, 3 MMAs per K step, 128 x 128 accumulator in TMEM), but
//   * the four operand tiles of a K slice arrive by TMA (cp.async.bulk.tensor.2d, SASS UTMALDG) into 128-byte-swizzled
//     shared memory: one elected producer thread, no per-thread address arithmetic, ragged rows zero-filled by the
//     tensor map, completion counted in bytes on a per-stage `full` mbarrier;
//   * one elected thread issues the tcgen05.mma's against SWIZZLE_128B K-major descriptors (8 rows x 128 B atoms,
//     SBO = 1024 B; the four K steps of a slice advance the start address by 32 B inside the atom) and commits each
//     slice to the stage's `empty` mbarrier, which is all the producer waits for;
//   * there is no CTA-wide barrier inside the K loop: loads run kGemmStages slices ahead of the tensor core.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t umma_desc_kmajor_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                      // LBO: unused for swizzled K-major operands
  d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32;  // SBO: next group of 8 rows
  d |= (uint64_t)1 << 46;                      // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                      // layout type SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}

struct GemmTmaMaps {
  alignas(64) unsigned char a_hi[128];
  alignas(64) unsigned char a_lo[128];
  alignas(64) unsigned char b_hi[128];
  alignas(64) unsigned char b_lo[128];
};

static __global__ void __launch_bounds__(128)
k_gram_tf32x3_tma(const __grid_constant__ GemmTmaMaps maps, float* __restrict__ out, int n1, int n2, int nk,
                  long long ld_out, float inv_d) {
  extern __shared__ __align__(1024) unsigned char gemm_smem[];
  constexpr uint32_t kTileBytes = kGemmBM * kGemmBK * sizeof(float);  // 128 rows x 128 B = 16 KB
  constexpr uint32_t kStageBytes = 4 * kTileBytes;                    // A_hi | A_lo | B_hi | B_lo
  __shared__ __align__(8) uint64_t full_bar[kGemmStages], empty_bar[kGemmStages], done_bar;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * kGemmBM, n0 = blockIdx.x * kGemmBN;
  // 1024-byte alignment of the swizzle atoms (the dynamic shared window itself is only 16-byte aligned)
  const uint32_t smem0 = (smem_u32(gemm_smem) + 1023u) & ~1023u;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"((uint32_t)kGemmBN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int s = 0; s < kGemmStages; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full_bar[s])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&empty_bar[s])));
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&done_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_d = tmem_base_s;

  if (warp == 0 && lane == 0) {
    // ---- producer: TMA loads, kGemmStages slices ahead ------------------------------------------
    for (int kt = 0; kt < nk; ++kt) {
      const int stage = kt % kGemmStages;
      if (kt >= kGemmStages) mbar_wait(smem_u32(&empty_bar[stage]), (uint32_t)(((kt / kGemmStages) - 1) & 1));
      const uint32_t bar = smem_u32(&full_bar[stage]);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kStageBytes) : "memory");
      const uint32_t sbase = smem0 + (uint32_t)stage * kStageBytes;
      const int k0 = kt * kGemmBK;
      tma_load_2d(sbase, maps.a_hi, k0, m0, bar);
      tma_load_2d(sbase + kTileBytes, maps.a_lo, k0, m0, bar);
      tma_load_2d(sbase + 2 * kTileBytes, maps.b_hi, k0, n0, bar);
      tma_load_2d(sbase + 3 * kTileBytes, maps.b_lo, k0, n0, bar);
    }
  } else if (warp == 1 && lane == 0) {
    // ---- MMA issuer ----------------------------------------------------------------------------------
    constexpr uint32_t idesc = umma_idesc_tf32(kGemmBM, kGemmBN);
    uint32_t acc = 0;
    for (int kt = 0; kt < nk; ++kt) {
      const int stage = kt % kGemmStages;
      mbar_wait(smem_u32(&full_bar[stage]), (uint32_t)((kt / kGemmStages) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;");
      const uint32_t sbase = smem0 + (uint32_t)stage * kStageBytes;
#pragma unroll
      for (int ks = 0; ks < kGemmBK / 8; ++ks) {  // one instruction = K 8 = 32 bytes inside the 128-byte atom
        const uint32_t koff = ks * 32;
        const uint64_t dah = umma_desc_kmajor_sw128(sbase + koff);
        const uint64_t dal = umma_desc_kmajor_sw128(sbase + kTileBytes + koff);
        const uint64_t dbh = umma_desc_kmajor_sw128(sbase + 2 * kTileBytes + koff);
        const uint64_t dbl = umma_desc_kmajor_sw128(sbase + 3 * kTileBytes + koff);
        umma_tf32(tmem_d, dal, dbh, idesc, acc);  // small terms first
        umma_tf32(tmem_d, dah, dbl, idesc, 1u);
        umma_tf32(tmem_d, dah, dbh, idesc, 1u);
        acc = 1u;
      }
      // the stage is free once these MMAs have read it
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
          smem_u32(&empty_bar[stage])));
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
        smem_u32(&done_bar)));
  }
  // ---- everyone: wait for the accumulator, then the epilogue ---------------------------------------
  mbar_wait(smem_u32(&done_bar), 0u);
  __syncwarp();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const int row = m0 + warp * 32 + lane;
  const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
  for (int c0 = 0; c0 < kGemmBN; c0 += 8) {
    uint32_t v[8];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)kGemmBN));
}

int launch_gram_tc_tma(cudaStream_t stream, const float* a_hi, const float* a_lo, int n1, const float* b_hi,
                              const float* b_lo, int n2, int d, float* out, long long ld_out) {
  const int d_pad = gram_pad_k(d);
  GemmTmaMaps maps;
  NTK_TRY(gemm_encode_tmap(maps.a_hi, a_hi, n1, d_pad));
  NTK_TRY(gemm_encode_tmap(maps.a_lo, a_lo, n1, d_pad));
  NTK_TRY(gemm_encode_tmap(maps.b_hi, b_hi, n2, d_pad));
  NTK_TRY(gemm_encode_tmap(maps.b_lo, b_lo, n2, d_pad));
  dim3 grid((n2 + kGemmBN - 1) / kGemmBN, (n1 + kGemmBM - 1) / kGemmBM);
  constexpr size_t smem = (size_t)kGemmStages * 4 * kGemmBM * kGemmBK * sizeof(float) + 1024;  // + alignment slack
  NTK_TRY(ensure_dynamic_smem((const void*)k_gram_tf32x3_tma, smem));
  k_gram_tf32x3_tma<<<grid, 128, smem, stream>>>(maps, out, n1, n2, d_pad / kGemmBK, ld_out, (float)(1.0 / (double)d));
  NTK_CUDA(cudaGetLastError());
  return NTK_OK;
}

}  // namespace ntk
