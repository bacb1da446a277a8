/* product_c.c -- the C-ABI used from plain C (no Python, no torch): the README product of the reference
 * (p = kde!(randn(2,100)), q = kde!(2 .+ randn(2,100)), 100 product samples, Niter = 5).
 *   gcc -std=c99 -Iinclude examples/product_c.c -o product_c -Lkerneldensityestimate.jl_b200 -lkdeb200 -lm
 *   LD_LIBRARY_PATH=kerneldensityestimate.jl_b200 ./product_c
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "kdeb200.h"

#define CHECK(call)                                                          \
  do {                                                                       \
    if ((call) != 0) {                                                       \
      fprintf(stderr, "%s failed: %s\n", #call, kdeb200_last_error());        \
      return 1;                                                              \
    }                                                                        \
  } while (0)

static double gauss(unsigned *s) { /* Box-Muller on a tiny LCG: this is only example data */
  double u1, u2;
  *s = *s * 1664525u + 1013904223u; u1 = ((*s >> 8) + 0.5) / 16777216.0;
  *s = *s * 1664525u + 1013904223u; u2 = ((*s >> 8) + 0.5) / 16777216.0;
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

static int make_tree(int d, int64_t N, const double *pts, double sigma, kdeb200_tree_t *out) {
  const int64_t NN = 2 * N;
  double *centers = calloc(NN * d, 8), *ranges = calloc(NN * d, 8), *w2 = calloc(NN, 8), *means = calloc(NN * d, 8),
         *bw = calloc(NN * d, 8), *w = malloc(N * 8), var[KDEB200_MAX_DIM];
  int64_t *l = calloc(NN, 8), *r = calloc(NN, 8), *lo = calloc(NN, 8), *hi = calloc(NN, 8), *perm = calloc(NN, 8);
  int rc;
  for (int64_t i = 0; i < N; ++i) w[i] = 1.0 / (double)N;
  for (int k = 0; k < d; ++k) var[k] = sigma * sigma;
  rc = kdeb200_tree_build_host(d, N, pts, w, var, centers, ranges, w2, means, bw, l, r, lo, hi, perm);
  if (rc == 0) rc = kdeb200_tree_create(d, N, means, bw, w2, l, r, perm, out);
  free(centers); free(ranges); free(w2); free(means); free(bw); free(w); free(l); free(r); free(lo); free(hi); free(perm);
  return rc;
}

int main(void) {
  enum { D = 2, N = 100, NP = 100, NITER = 5 };
  double p[D * N], q[D * N], pts[D * NP], mean[D] = {0, 0};
  int64_t ind[2 * NP];
  kdeb200_tree_t trees[2];
  unsigned seed = 12345u;
  int ndev = 0, nlev = 0;
  int64_t perU = 0, perN = 0, evals = 0;
  for (int i = 0; i < D * N; ++i) { p[i] = gauss(&seed); q[i] = 2.0 + gauss(&seed); }
  CHECK(kdeb200_device_count(&ndev));
  if (ndev == 0) { fprintf(stderr, "no CUDA device: libkdeb200 has no CPU fallback\n"); return 2; }
  CHECK(kdeb200_init(0));
  CHECK(make_tree(D, N, p, 0.4, &trees[0]));
  CHECK(make_tree(D, N, q, 0.4, &trees[1]));
  CHECK(kdeb200_gibbs_sizes(trees, 2, NITER, &nlev, &perU, &perN, &evals));
  CHECK(kdeb200_gibbs(trees, 2, NP, NITER, 1, NULL, NULL, 0, NULL, 0, 20261017ull, 0, NP, pts, ind, NULL));
  for (int s = 0; s < NP; ++s) for (int k = 0; k < D; ++k) mean[k] += pts[s * D + k] / NP;
  printf("levels=%d uniforms/sample=%lld evals/sample=%lld  product mean = (%.3f, %.3f)  [expected near (1, 1)]\n",
         nlev, (long long)perU, (long long)evals, mean[0], mean[1]);
  CHECK(kdeb200_tree_destroy(trees[0]));
  CHECK(kdeb200_tree_destroy(trees[1]));
  CHECK(kdeb200_shutdown());
  return (fabs(mean[0] - 1.0) < 0.5 && fabs(mean[1] - 1.0) < 0.5) ? 0 : 3;
}
