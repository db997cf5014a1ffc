// Micro-benchmark: is the register-file read bandwidth shared between the FMA pipe and the other
// pipes on B200, and does the operand-reuse cache lift the 3-source FFMA2 limit?
// Each test runs a loop of NI independent instruction groups per thread; rates are per SMSP.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a rf_bandwidth.cu -o rf_bandwidth
#include <cstdio>
#include <cuda_runtime.h>

// MODE 0: FFMA2 x = x*y_i + imm           (2 reg sources, distinct y per chain)
// MODE 1: PRMT  m = prmt(m, n_i)              (2 reg sources) alone
// MODE 2: MODE 0 and MODE 1 interleaved 1:1
// MODE 3: FFMA2 x = x*y + z, SAME y,z for all chains (3 reg sources, reuse-cache friendly)
// MODE 4: FFMA2 x = x*y_i + z_i           (3 reg sources, all distinct)
// MODE 5: MODE 0 interleaved 2:1 with MUFU.RSQ
// MODE 6: FFMA2 x = x*y + imm, SAME y for all chains (2 reg sources, one reusable)
// MODE 7: MODE 6 interleaved 1:1 with LOP3
// MODE 8: scalar FFMA x = x*y + z same y,z (reuse friendly)
template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
  float2 x[8], y[8], z[8];
  unsigned m[8], n[8];
  float f[8], sx[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x[i] = make_float2(threadIdx.x * 0.001f + i, 1.f + i);
    y[i] = make_float2(a + threadIdx.x * 1e-6f + i * 1e-3f, a - i * 1e-3f - threadIdx.x * 1e-6f);
    z[i] = make_float2(b * threadIdx.x * 1e-3f + i, b + i);
    m[i] = threadIdx.x * 2654435761u + i;
    n[i] = threadIdx.x * 40503u + i * 977u;
    f[i] = 1.f + threadIdx.x + i;
    sx[i] = threadIdx.x * 0.001f + i;
  }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0 || MODE == 2 || MODE == 5) x[i] = __ffma2_rn(x[i], y[i], make_float2(1.25f, 1.25f));
        if (MODE == 1 || MODE == 2 || MODE == 7) m[i] = __byte_perm(m[i], n[i], 0x5140);
        if (MODE == 3) x[i] = __ffma2_rn(x[i], y[0], z[0]);
        if (MODE == 4) x[i] = __ffma2_rn(x[i], y[i], z[i]);
        if (MODE == 5 && (i & 1)) f[i] = rsqrtf(f[i]);
        if (MODE == 6 || MODE == 7) x[i] = __ffma2_rn(x[i], y[0], make_float2(1.25f, 1.25f));
        if (MODE == 8) sx[i] = __fmaf_rn(sx[i], y[0].x, z[0].x);
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y + __uint_as_float(m[i] & 0x3f800000u) + f[i] + sx[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) ((long long*)out)[1 << 20] = t1 - t0;
}

template <int MODE>
void run(const char* name, float* d, int warps_per_sm) {
  const int iters = 2000;
  int threads = warps_per_sm * 32;
  k<MODE><<<148, threads>>>(d, iters, 1.0001f, 0.5f);
  cudaDeviceSynchronize();
  k<MODE><<<148, threads>>>(d, iters, 1.0001f, 0.5f);
  cudaDeviceSynchronize();
  long long cyc; cudaMemcpy(&cyc, (long long*)d + (1 << 20), 8, cudaMemcpyDeviceToHost);
  double groups = (double)iters * 64 * (warps_per_sm / 4.0);
  printf("%-44s warps/SM=%2d  cycles per group per SMSP = %.3f\n", name, warps_per_sm, (double)cyc / groups);
}

int main() {
  float* d; cudaMalloc(&d, (1 << 23) + 64);
  for (int w : {8, 16}) {
    run<0>("FFMA2 R,R,imm (distinct y)", d, w);
    run<6>("FFMA2 R,Rsame,imm", d, w);
    run<1>("PRMT R,R", d, w);
    run<2>("FFMA2 R,R,imm + PRMT R,R", d, w);
    run<7>("FFMA2 R,Rsame,imm + PRMT R,R", d, w);
    run<4>("FFMA2 R,R,R (distinct)", d, w);
    run<3>("FFMA2 R,Rsame,Rsame", d, w);
    run<8>("FFMA R,Rsame,Rsame", d, w);
    run<5>("FFMA2 R,R,imm + 0.5 MUFU.RSQ", d, w);
  }
  return 0;
}
