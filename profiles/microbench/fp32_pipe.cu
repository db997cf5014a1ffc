// Micro-benchmark: FP32 pipe throughput on B200 for register-register vs immediate forms.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a fp32_pipe.cu -o fp32_pipe
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
  float x[8], y[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 0.001f + i; y[i] = a + i * 0.5f; }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) x[i] = __fmul_rn(x[i], y[i]);                 // FMUL R,R,R
        if (MODE == 1) x[i] = __fmaf_rn(x[i], y[i], y[(i + 1) & 7]); // FFMA R,R,R,R (3 distinct src)
        if (MODE == 2) x[i] = __fmaf_rn(x[i], y[i], 1.25f);          // FFMA R,R,R,imm
        if (MODE == 3) x[i] = __fadd_rn(x[i], y[i]);                 // FADD R,R,R
        if (MODE == 4) x[i] = __fmaf_rn(x[i], x[i], x[i]);           // FFMA same reg
        if (MODE == 5) x[i] = __fmaf_rn(x[i], 0.999f, y[i]);         // FFMA R,R,imm,R
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) ((long long*)out)[1 << 20] = t1 - t0;
}

template <int MODE>
void run(const char* name, float* d, int warps_per_sm) {
  const int iters = 2000;
  int threads = warps_per_sm * 32;
  k<MODE><<<148, threads>>>(d, iters, 1.0001f, 0.5f);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<148, threads>>>(d, iters, 1.0001f, 0.5f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long cyc; cudaMemcpy(&cyc, (long long*)d + (1 << 20), 8, cudaMemcpyDeviceToHost);
  double inst_per_warp = (double)iters * 64;
  double ipc_smsp = inst_per_warp * (warps_per_sm / 4.0) / (double)cyc;
  printf("%-22s warps/SM=%2d  cycles=%lld  warp-instr/clk/SMSP=%.3f  (%.1f us)\n", name, warps_per_sm, cyc, ipc_smsp, ms * 1e3);
}

int main() {
  float* d; cudaMalloc(&d, (1 << 23) + 64);
  for (int w : {4, 8, 16}) {
    run<0>("FMUL R,R,R", d, w);
    run<1>("FFMA R,R,R,R", d, w);
    run<2>("FFMA R,R,R,imm", d, w);
    run<3>("FADD R,R,R", d, w);
    run<4>("FFMA x,x,x", d, w);
    run<5>("FFMA R,imm,R", d, w);
  }
  return 0;
}
