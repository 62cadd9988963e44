// Microbenchmark: peak FP32 FMA issue rate on B200 with scalar FFMA (3-register form) vs packed fma.rn.f32x2.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ffma_peak tools/ffma_peak.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a0, float b0) {
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = threadIdx.x * 1e-3f + i;
  float a[4] = {a0, a0 + 1e-3f, a0 + 2e-3f, a0 + 3e-3f}, b[4] = {b0, b0 * 1.01f, b0 * 1.02f, b0 * 1.03f};
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = fmaf(a[i & 3], b[(i >> 2) & 3], acc[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        unsigned long long c, av, bv;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(acc[i]), "f"(acc[i + 1]));
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(av) : "f"(a[i & 3]), "f"(a[(i + 1) & 3]));
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(bv) : "f"(b[(i >> 2) & 3]), "f"(b[(i >> 2) & 3]));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(av), "l"(bv));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(acc[i]), "=f"(acc[i + 1]) : "l"(c));
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name) {
  float* out;
  const int blocks = 148 * 8, iters = 20000;
  cudaMalloc(&out, blocks * 256 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, 256>>>(out, 100, 1.0001f, 0.9999f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<blocks, 256>>>(out, iters, 1.0001f, 0.9999f);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double flops = 2.0 * 32 * (double)iters * blocks * 256;
  printf("%s: %.3f ms  %.1f TFLOP/s (%s)\n", name, ms, flops / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  run<0>("FFMA  (scalar, 3-reg)");
  run<1>("FFMA2 (fma.rn.f32x2)");
  return 0;
}
