// Instruction-mix microbenchmark: the arithmetic of one Gibbs kernel evaluation (variants A and C)
// in a register-only loop, to separate pipe limits from memory/sync effects.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mix mix.cu ; run: ./mix
#include <cstdio>
#include <cuda_runtime.h>
#include "../../kerneldensityestimate.jl_b200/csrc/common.cuh"
using namespace kdeb200;
namespace kdeb200 { void set_error(const char*, ...) {} }

// VAR: 0 = A (hoisted), 1 = C (per-dim rcp + rsqrt), 2 = C' (one rsqrt of the product, reciprocals by products)
template <int VAR, int UNR, bool MUFU>
__global__ void __launch_bounds__(128, 4) k(int iters, const double *in, double *out, const __grid_constant__ ExpConsts ec) {
  __shared__ double tab[KDE_EXP_TAB];
  for (int i = threadIdx.x; i < KDE_EXP_TAB; i += blockDim.x) tab[i] = in[i & 63];
  __syncthreads();
  double mu[3], ich[3], cadd[3];
  for (int k = 0; k < 3; ++k) { mu[k] = in[64 + k] + threadIdx.x * 1e-3; ich[k] = -in[67 + k]; cadd[k] = in[70 + k]; }
  double S = 0;
  double m[UNR][3], b[UNR][3], lw[UNR];
  for (int u = 0; u < UNR; ++u) for (int k = 0; k < 3; ++k) { m[u][k] = in[80 + u * 3 + k]; b[u][k] = in[100 + u * 3 + k]; lw[u] = -in[120 + u]; }
  for (int it = 0; it < iters; ++it) {
    double p[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      double arg, sc = 1.0;
      if (VAR == 0) {
        arg = lw[u];
#pragma unroll
        for (int k = 0; k < 3; ++k) { double df = __dadd_rn(m[u][k], -mu[k]); arg = __fma_rn(__dmul_rn(df, df), ich[k], arg); }
      } else if (VAR == 2) {
        double c[3], df2[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) { c[k] = __dadd_rn(b[u][k], cadd[k]); double df = __dadd_rn(m[u][k], -mu[k]); df2[k] = __dmul_rn(df, df); }
        const double c01 = __dmul_rn(c[0], c[1]);
        const double P = __dmul_rn(c01, c[2]);
        sc = kde_rsqrt(P);
        const double Rv = __dmul_rn(sc, sc);
        const double t = __dmul_rn(c[2], Rv);
        double quad = __dmul_rn(df2[2], __dmul_rn(c01, Rv));
        quad = __fma_rn(df2[0], __dmul_rn(c[1], t), quad);
        quad = __fma_rn(df2[1], __dmul_rn(c[0], t), quad);
        arg = __fma_rn(quad, -0.5, lw[u]);
      } else {
        double quad = 0, prod = 1.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          double c = __dadd_rn(b[u][k], cadd[k]);
          double df = __dadd_rn(m[u][k], -mu[k]);
          double ic;
          if (MUFU) ic = kde_rcp(c); else { double e = __fma_rn(-c, 0.3, 1.0); ic = __fma_rn(0.3, __fma_rn(e, e, e), 0.3); }
          quad = __fma_rn(__dmul_rn(df, df), ic, quad);
          prod = __dmul_rn(prod, c);
        }
        arg = __fma_rn(quad, -0.5, lw[u]);
        if (MUFU) sc = kde_rsqrt(prod); else { double e = __fma_rn(-prod, 0.25, 1.0); double uu = __dmul_rn(__fma_rn(e, 0.375, 0.5), e); sc = __fma_rn(0.5, uu, 0.5); }
      }
      double e = kde_exp_flush(arg, tab, ec);
      p[u] = VAR ? __dmul_rn(e, sc) : e;
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) S = __dadd_rn(S, p[u]);
#pragma unroll
    for (int k = 0; k < 3; ++k) mu[k] = __dadd_rn(mu[k], 1e-9);  // keep the loop honest
    if (VAR) {
#pragma unroll
      for (int u = 0; u < UNR; ++u)
#pragma unroll
        for (int k = 0; k < 3; ++k) b[u][k] = __dadd_rn(b[u][k], 1e-9);  // per-node variances change every iteration
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = S;
}

template <int VAR, int UNR, bool MUFU>
void run(const char *name, int fp64_per_eval, int bps, const double *din, double *dout) {
  int iters = 20000;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<VAR, UNR, MUFU><<<148 * bps, 128>>>(iters / 10, din, dout, make_exp_consts());
  cudaEventRecord(a);
  k<VAR, UNR, MUFU><<<148 * bps, 128>>>(iters, din, dout, make_exp_consts());
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double evals = 148.0 * bps * 128 * (double)iters * UNR;
  double fp = evals * (fp64_per_eval + 3.0 / UNR + (VAR ? 3.0 : 0.0));
  printf("%-28s bps=%d  %.3e evals/s   FP64 lane-ops/s %.3e = %.1f%% of 1.826e13\n", name, bps, evals / (ms * 1e-3), fp / (ms * 1e-3), 100 * fp / (ms * 1e-3) / 1.826e13);
}

int main() {
  double h[256]; for (int i = 0; i < 256; ++i) h[i] = 0.5 + 0.01 * i; for (int j = 0; j < 64; ++j) h[j] = exp2(j / 64.0);
  double *din, *dout; cudaMalloc(&din, sizeof(h)); cudaMalloc(&dout, 8 * 148 * 8 * 128); cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice);
  for (int bps = 2; bps <= 4; bps *= 2) {
    run<0, 1, true>("A unr1", 17, bps, din, dout);
    run<0, 2, true>("A unr2", 17, bps, din, dout);
    run<0, 4, true>("A unr4", 17, bps, din, dout);
    run<1, 1, true>("C unr1 (MUFU)", 41, bps, din, dout);
    run<1, 2, true>("C unr2 (MUFU)", 41, bps, din, dout);
    run<1, 4, true>("C unr4 (MUFU)", 41, bps, din, dout);
    run<1, 2, false>("C unr2 (no MUFU, same DFMAs)", 41, bps, din, dout);
    run<2, 1, true>("C' unr1 (product trick)", 38, bps, din, dout);
    run<2, 2, true>("C' unr2 (product trick)", 38, bps, din, dout);
    run<2, 4, true>("C' unr4 (product trick)", 38, bps, din, dout);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
