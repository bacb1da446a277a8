/* product_multi_c.c -- ONE plain-C process drives every GPU of the box through the C-ABI (no Python, no torch, no
 * MPI): the C4-shaped product of BASELINE.json (8 densities x 4096 components, 3-D, Niter = 5) with
 * samples_per_gpu x ngpus product samples, block-partitioned over the GPUs by kdeb200_gibbs itself after
 * kdeb200_init_multi; every GPU copies its shard straight into the caller's host arrays.
 *   gcc -std=c99 -O2 -Iinclude examples/product_multi_c.c -o product_multi_c -Lkerneldensityestimate.jl_b200 -lkdeb200 -lm
 *   ./product_multi_c [ngpus (0 = all)] [samples_per_gpu (default 1000000)]
 * Prints one JSON line: samples/s on 1 GPU and on all GPUs (wall clock around kdeb200_gibbs: staging, kernels and
 * the D2H of 88 bytes per sample included), and whether a slice of the multi-GPU result equals the 1-GPU result. */
#define _POSIX_C_SOURCE 199309L
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "kdeb200.h"

#define CHECK(call)                                                   \
  do {                                                                \
    if ((call) != 0) {                                                \
      fprintf(stderr, "%s failed: %s\n", #call, kdeb200_last_error()); \
      return 1;                                                       \
    }                                                                 \
  } while (0)

static double now(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static double gauss(unsigned *s) { /* Box-Muller on a tiny LCG: example data only */
  double u1, u2;
  *s = *s * 1664525u + 1013904223u; u1 = ((*s >> 8) + 0.5) / 16777216.0;
  *s = *s * 1664525u + 1013904223u; u2 = ((*s >> 8) + 0.5) / 16777216.0;
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

static int make_tree(int d, int64_t N, const double *pts, const double *var, kdeb200_tree_t *out) {
  const int64_t NN = 2 * N;
  double *centers = calloc(NN * d, 8), *ranges = calloc(NN * d, 8), *w2 = calloc(NN, 8), *means = calloc(NN * d, 8),
         *bw = calloc(NN * d, 8), *w = malloc(N * 8);
  int64_t *l = calloc(NN, 8), *r = calloc(NN, 8), *lo = calloc(NN, 8), *hi = calloc(NN, 8), *perm = calloc(NN, 8);
  int rc;
  for (int64_t i = 0; i < N; ++i) w[i] = 1.0 / (double)N;
  rc = kdeb200_tree_build_host(d, N, pts, w, var, centers, ranges, w2, means, bw, l, r, lo, hi, perm);
  if (rc == 0) rc = kdeb200_tree_create(d, N, means, bw, w2, l, r, perm, out);
  free(centers); free(ranges); free(w2); free(means); free(bw); free(w); free(l); free(r); free(lo); free(hi); free(perm);
  return rc;
}

int main(int argc, char **argv) {
  enum { D = 3, M = 8, N = 4096, NITER = 5 };
  int want = argc > 1 ? atoi(argv[1]) : 0, ndev = 0, G = 0;
  int64_t per_gpu = argc > 2 ? atoll(argv[2]) : 1000000;
  kdeb200_tree_t trees[M];
  unsigned seed = 2026u;
  static double pts[D * N];
  CHECK(kdeb200_device_count(&ndev));
  if (ndev == 0) { fprintf(stderr, "no CUDA device: libkdeb200 has no CPU fallback\n"); return 2; }
  CHECK(kdeb200_init(0));
  for (int j = 0; j < M; ++j) { /* 4-component mixtures on corners of {+-2}^3, shifted 0.25 j along dim 1 */
    double var[D];
    for (int i = 0; i < N; ++i) {
      seed = seed * 1664525u + 1013904223u;
      const unsigned c = (seed >> 13) & 3u;
      pts[i * D + 0] = -2.0 + 0.6 * gauss(&seed) + 0.25 * j;
      pts[i * D + 1] = ((c & 2u) ? 2.0 : -2.0) + 0.6 * gauss(&seed);
      pts[i * D + 2] = ((c & 1u) ? 2.0 : -2.0) + 0.6 * gauss(&seed);
    }
    for (int k = 0; k < D; ++k) { /* Silverman */
      double m = 0, s = 0;
      for (int i = 0; i < N; ++i) m += pts[i * D + k] / N;
      for (int i = 0; i < N; ++i) s += (pts[i * D + k] - m) * (pts[i * D + k] - m) / (N - 1);
      const double h = sqrt(s) * pow(4.0 / ((D + 2.0) * N), 1.0 / (D + 4.0));
      var[k] = h * h;
    }
    CHECK(make_tree(D, N, pts, var, &trees[j]));
  }
  const int64_t n1 = per_gpu;
  double *p1 = malloc(sizeof(double) * D * n1);
  int64_t *i1 = malloc(sizeof(int64_t) * M * n1);
  /* 1 GPU: warm-up, then timed */
  CHECK(kdeb200_gibbs(trees, M, 4096, NITER, 1, NULL, NULL, 0, NULL, 0, 7ull, 0, 4096, p1, i1, NULL));
  CHECK(kdeb200_init_multi(want));
  CHECK(kdeb200_multi_count(&G));
  const int64_t nG = per_gpu * G;
  double *pG = malloc(sizeof(double) * D * nG);
  int64_t *iG = malloc(sizeof(int64_t) * M * nG);
  if (!p1 || !i1 || !pG || !iG) { fprintf(stderr, "out of host memory\n"); return 4; }
  CHECK(kdeb200_gibbs(trees, M, nG, NITER, 1, NULL, NULL, 0, NULL, 0, 7ull, 0, 4096 * (int64_t)G, pG, iG, NULL)); /* replicate + warm */
  double t0 = now();
  CHECK(kdeb200_gibbs(trees, M, nG, NITER, 1, NULL, NULL, 0, NULL, 0, 20261017ull, 0, nG, pG, iG, NULL));
  const double tG = now() - t0;
  double kms = 0; int nl = 0;
  kdeb200_last_kernel_ms(&kms, &nl);
  CHECK(kdeb200_init_multi(1));
  t0 = now();
  CHECK(kdeb200_gibbs(trees, M, nG, NITER, 1, NULL, NULL, 0, NULL, 0, 20261017ull, 0, n1, p1, i1, NULL));
  const double t1 = now() - t0;
  /* the first per_gpu samples of the nG-sample run, drawn on one GPU, must equal the multi-GPU run's rows bit for bit;
   * a slice from the LAST device's block is recomputed too */
  int same = memcmp(p1, pG, sizeof(double) * D * n1) == 0 && memcmp(i1, iG, sizeof(int64_t) * M * n1) == 0;
  const int64_t a = nG - 4096;
  CHECK(kdeb200_gibbs(trees, M, nG, NITER, 1, NULL, NULL, 0, NULL, 0, 20261017ull, a, nG, p1, i1, NULL));
  same = same && memcmp(p1, pG + a * D, sizeof(double) * D * 4096) == 0 && memcmp(i1, iG + a * M, sizeof(int64_t) * M * 4096) == 0;
  printf("{\"workload\": \"C4 from one C process: 8 x 4096 components, 3-D, Niter=5\", \"n_gpus\": %d, \"samples\": %lld, "
         "\"wall_s\": %.4f, \"samples_per_s\": %.1f, \"slowest_kernel_ms\": %.2f, \"launches\": %d, "
         "\"one_gpu_samples\": %lld, \"one_gpu_wall_s\": %.4f, \"one_gpu_samples_per_s\": %.1f, \"speedup\": %.3f, "
         "\"identical_to_one_gpu\": %s}\n",
         G, (long long)nG, tG, nG / tG, kms, nl, (long long)n1, t1, n1 / t1, (nG / tG) / (n1 / t1), same ? "true" : "false");
  for (int j = 0; j < M; ++j) CHECK(kdeb200_tree_destroy(trees[j]));
  CHECK(kdeb200_shutdown());
  free(p1); free(i1); free(pG); free(iG);
  return same ? 0 : 3;
}
