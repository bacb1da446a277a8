// Throughput of individual FP64 instructions and of FP64 + MUFU.RSQ64H mixes (chip-wide, 4 warps / scheduler).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(128, 4) k(int iters, double *out, double seed) {
  double a[8], b = seed, c = seed * 0.5;
  for (int i = 0; i < 8; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i) + seed;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = __fma_rn(a[i], b, c);
      if (MODE == 1) a[i] = __dmul_rn(a[i], b);
      if (MODE == 2) a[i] = __dadd_rn(a[i], c);
      if (MODE == 3) { a[i] = __dmul_rn(a[i], a[(i + 1) & 7]); }            // two register sources
      if (MODE == 4) { a[i] = __fma_rn(a[i], a[(i + 1) & 7], a[(i + 2) & 7]); }  // three register sources
      if (MODE == 5) {                                                       // 1 MUFU.RSQ64H per 8 DFMA
        a[i] = __fma_rn(a[i], b, c);
        if (i == 0) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a[7])); a[7] = y + 1.0; }
      }
      if (MODE == 6) {                                                       // 1 MUFU.RSQ64H per 2 DFMA
        a[i] = __fma_rn(a[i], b, c);
        if ((i & 1) == 0) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a[i | 1])); a[i | 1] = y + 1.0; }
      }
    }
  }
  double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name, double fp_per_iter, double *d) {
  int iters = 50000; cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 4, 128>>>(iters / 10, d, 1e-3);
  cudaEventRecord(e0); k<MODE><<<148 * 4, 128>>>(iters, d, 1e-3); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = 148.0 * 4 * 128 * iters * fp_per_iter;
  printf("%-34s %.3e FP64 lane-ops/s = %.1f%% of 1.861e13 nominal\n", name, ops / (ms * 1e-3), 100 * ops / (ms * 1e-3) / 1.861e13);
}
int main() {
  double *d; cudaMalloc(&d, 8 * 148 * 4 * 128);
  run<0>("DFMA reg,const,const", 8, d); run<1>("DMUL reg,const", 8, d); run<2>("DADD reg,const", 8, d);
  run<3>("DMUL reg,reg", 8, d); run<4>("DFMA reg,reg,reg", 8, d);
  run<5>("DFMA + RSQ64H (1 per 8)", 9, d); run<6>("DFMA + RSQ64H (1 per 2)", 12, d);
  printf("%s\n", cudaGetErrorString(cudaGetLastError())); return 0;
}
