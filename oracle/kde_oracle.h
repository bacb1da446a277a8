/*
 * kde_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, never linked into the product).
 *
 * Plain-C restatement, in reference loop order, of the hot path of
 * JuliaRobotics/KernelDensityEstimate.jl v0.5.13 (see kde_oracle.c for per-function
 * file:line citations).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.
 *
 * Indices stored inside the tree arrays are 1-based node ids exactly as in the
 * reference (NO_CHILD = -1, permutation 0 for internal nodes); array storage is
 * 0-based C, i.e. reference node i lives at C index i-1.
 */
#ifndef KDE_ORACLE_H
#define KDE_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct okde {
  int64_t dims, num_points;
  /* BallTree (src/BallTree01.jl:10-28) */
  double *centers, *ranges, *weights;
  int64_t *left_child, *right_child, *lowest_leaf, *highest_leaf, *permutation;
  int64_t next;
  /* BallTreeDensity (src/BallTreeDensity01.jl:11-24) */
  int64_t multibandwidth;
  double *means, *bandwidth, *bandwidthMin, *bandwidthMax;
} okde;

/* construction / destruction */
okde *okde_make(int64_t d, int64_t N, const double *points /*d x N col-major*/,
                const double *weights /*N, used as given*/, const double *bwvar /*d variances*/);
okde *okde_kde_bw(int64_t d, int64_t N, const double *points, const double *ks /*nks std-devs*/,
                  int64_t nks, const double *weights /*N or NULL => ones*/);
okde *okde_kde_lcv(int64_t d, int64_t N, const double *points, int64_t *n_loo_calls /*optional*/);
void okde_free(okde *p);

/* getters (src/KDE01.jl:91-136) */
void okde_get_points(const okde *p, double *out /*d x N*/);
void okde_get_bw(const okde *p, double *out /*d x N std-devs*/);
void okde_get_weights(const okde *p, double *out /*N*/);
okde *okde_marginal(const okde *p, const int64_t *ind /*1-based dims*/, int64_t nind);

/* raw array access for the python binding */
int64_t okde_dims(const okde *p);
int64_t okde_npts(const okde *p);
const double *okde_arr_f(const okde *p, int which);  /* 0 centers 1 ranges 2 weights 3 means 4 bandwidth 5 bwMin 6 bwMax */
const int64_t *okde_arr_i(const okde *p, int which); /* 0 left 1 right 2 lowest 3 highest 4 perm */

/* evaluation (src/DualTree01.jl) */
int okde_evaluate(const okde *bd, const okde *loc /* ==bd => leave-one-out */, double *p /*M, original order*/);
int okde_eval_points(const okde *bd, int64_t M, const double *pos /*d x M*/, double *p);
double okde_eval_avg_logl(const okde *bd1, const okde *bd2);
double okde_entropy(const okde *bd);

/* cross validation (src/CrossValidation.jl) */
double okde_nloo_ll(double alpha, okde *bd);
void okde_neighbor_minmax(const okde *bd, double *minm, double *maxm);
okde *okde_ksize(const okde *bd, int64_t *n_loo_calls);

/* multiscale Gibbs (src/MSGibbs01.jl) */
int64_t okde_gibbs_nlevels(const okde *const *trees, int64_t ndens);
/* samples s in [s0, s1) (0-based); pts / ind are the FULL d*Np and ndens*Np buffers */
int okde_gibbs(int64_t ndens, const okde *const *trees, int64_t Np, int64_t Niter,
               double *pts, int64_t *ind, const double *randU, int64_t nU,
               const double *randN, int64_t nN, int add_entropy,
               const uint8_t *mask /* ndens*d bytes (1 = active) or NULL */,
               int64_t s0, int64_t s1);

/* same with labelsChoosen recording (src/MSGibbs01.jl:109-112): record[(s*ndens + j)*Nlevels + (l-1)] */
int okde_gibbs_record(int64_t ndens, const okde *const *trees, int64_t Np, int64_t Niter, double *pts,
                      int64_t *ind, const double *randU, int64_t nU, const double *randN, int64_t nN,
                      int add_entropy, const uint8_t *mask, int64_t s0, int64_t s1, int64_t *record);

/* bounded-sample timing helpers for the CPU baseline (OpenMP over independent rows/chains) */
int okde_gibbs_omp(int64_t ndens, const okde *const *trees, int64_t Np, int64_t Niter,
                   double *pts, int64_t *ind, const double *randU, int64_t nU,
                   const double *randN, int64_t nN, int add_entropy, const uint8_t *mask,
                   int nthreads);
int okde_eval_points_omp(const okde *bd, int64_t M, const double *pos, double *p, int nthreads);
int okde_max_threads(void);
/* LOO densities of the leaf rows [j0, j1) (leaf order) of evaluate(bd, bd): full-size spot checks */
int okde_loo_rows(const okde *bd, int64_t j0, int64_t j1, double *out, int nthreads);
/* emulate Julia's @simd sum(lambdas) for Ndens >= 16 with vf lanes x ic interleaved accumulators; (0,0) = sequential */
void okde_set_sum_simd(int vf, int ic);

#ifdef __cplusplus
}
#endif
#endif
