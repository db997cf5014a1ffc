// Micro-benchmark: packed FP32 (fma/mul/add .f32x2 -> SASS FFMA2/FMUL2/FADD2) issue rate on B200
// against the scalar forms, per SMSP.  Answers: does packing halve the issue slots of the
// box-filter / polynomial arithmetic of k_stage, and what is the FMA-pipe cost of one packed op?
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a fp32x2_pipe.cu -o fp32x2_pipe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pk(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float lo_of(unsigned long long v) {
  float a, b;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
  return a + b;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// MODE 0: FFMA2 R,R,R (3 distinct)   1: FMUL2   2: FADD2   3: scalar FFMA R,R,R,R   4: scalar FADD
// MODE 5: mixed 1 FFMA2 + 1 LOP3 (integer ALU work co-issued)   6: 1 scalar FFMA + 1 LOP3
// MODE 7: FFMA2 with a, a, c (two distinct sources)
template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
  unsigned long long x[8], y[8];
  float xs[8], ys[8];
  unsigned m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    xs[i] = threadIdx.x * 0.001f + i; ys[i] = a + i * 0.5f + threadIdx.x * 1e-6f;
    x[i] = pk(xs[i], xs[i] + 1.f); y[i] = pk(ys[i], ys[i] * b);
    m[i] = threadIdx.x * 2654435761u + i;
  }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) x[i] = fma2(x[i], y[i], y[(i + 1) & 7]);
        if (MODE == 1) x[i] = mul2(x[i], y[i]);
        if (MODE == 2) x[i] = add2(x[i], y[i]);
        if (MODE == 3) xs[i] = __fmaf_rn(xs[i], ys[i], ys[(i + 1) & 7]);
        if (MODE == 4) xs[i] = __fadd_rn(xs[i], ys[i]);
        if (MODE == 5) { x[i] = fma2(x[i], y[i], y[(i + 1) & 7]); m[i] = (m[i] ^ m[(i + 3) & 7]) & (m[(i + 5) & 7] | 0x55u); }
        if (MODE == 6) { xs[i] = __fmaf_rn(xs[i], ys[i], ys[(i + 1) & 7]); m[i] = (m[i] ^ m[(i + 3) & 7]) & (m[(i + 5) & 7] | 0x55u); }
        if (MODE == 7) x[i] = fma2(x[i], x[i], y[i]);
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += lo_of(x[i]) + xs[i] + __uint_as_float(m[i] & 0x3f800000u);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) ((long long*)out)[1 << 20] = t1 - t0;
}

template <int MODE>
void run(const char* name, float* d, int warps_per_sm, double flop_per_inst) {
  const int iters = 2000;
  int threads = warps_per_sm * 32;
  k<MODE><<<148, threads>>>(d, iters, 1.0001f, 0.5f);
  cudaDeviceSynchronize();
  k<MODE><<<148, threads>>>(d, iters, 1.0001f, 0.5f);
  cudaDeviceSynchronize();
  long long cyc; cudaMemcpy(&cyc, (long long*)d + (1 << 20), 8, cudaMemcpyDeviceToHost);
  double inst_per_warp = (double)iters * 64;
  double ipc_smsp = inst_per_warp * (warps_per_sm / 4.0) / (double)cyc;
  printf("%-26s warps/SM=%2d cycles=%9lld  (FP) warp-instr/clk/SMSP=%.3f  lane-ops/clk/SMSP=%.1f\n", name,
         warps_per_sm, cyc, ipc_smsp, ipc_smsp * 32 * flop_per_inst);
}

int main() {
  float* d; cudaMalloc(&d, (1 << 23) + 64);
  for (int w : {4, 8, 16}) {
    run<0>("FFMA2 R,R,R,R", d, w, 2);
    run<7>("FFMA2 x,x,y", d, w, 2);
    run<1>("FMUL2", d, w, 2);
    run<2>("FADD2", d, w, 2);
    run<3>("FFMA R,R,R,R", d, w, 1);
    run<4>("FADD", d, w, 1);
    run<5>("FFMA2 + LOP3 pair", d, w, 2);
    run<6>("FFMA + LOP3 pair", d, w, 1);
  }
  return 0;
}
