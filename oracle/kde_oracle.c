/*
 * kde_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; never part of the product path).
 *
 * Literal, scalar, reference-loop-order restatement in plain C of the hot path of
 * JuliaRobotics/KernelDensityEstimate.jl v0.5.13 (pure Julia; cannot be executed in this
 * image -- no julia binary -- hence this restatement).  Every function cites the
 * reference file:line it follows.  Compile with -ffp-contract=off so that no FMA is
 * introduced (Julia does not contract outside @fastmath).
 *
 * Parity pin: tests/test_oracle_golden.py checks this file against every
 * enabled golden fixture of the reference's own test-suite (test/testdata/test1DResult.txt,
 * test2DResult.txt, test2DvarResult.txt, test1Dlcv100Result.txt -> tests/golden/):
 *   - tree build: PINNED (all arrays, indices exact);
 *   - kernel sums / leave-one-out / likelihood / golden-section / kde!: PINNED through the LOOCV
 *     fixture (bandwidth 0.00272597, 20 nLOO_LL calls) and analytic normal-pdf answers;
 *   - Gibbs labels and points: PARITY UNPINNED by reference fixtures -- the reference has only
 *     statistical tests for them and cannot be executed here (no julia binary); that part of the
 *     oracle is held by the reference's statistical bands, the analytic Gaussian product, the
 *     precision-weighted-mean identity and the verified stream-consumption counts.
 *
 * Known, documented deviations from bit-level Julia behaviour (all at the 1e-16 level):
 *   - Julia's exp/log/@fastmath exp are not bit-identical to glibc's.
 *   - Julia's sum(::Vector) may be SIMD-reassociated; here every sum is sequential.
 */
#include "kde_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NO_CHILD (-1) /* src/BallTree01.jl:5 */

/* 1-based accessors mirroring src/BallTree01.jl:64-94 and src/BallTreeDensity01.jl:86-101 */
#define CEN(p, i, k) ((p)->centers[((i)-1) * (p)->dims + ((k)-1)])
#define RNG(p, i, k) ((p)->ranges[((i)-1) * (p)->dims + ((k)-1)])
#define MEA(p, i, k) ((p)->means[((i)-1) * (p)->dims + ((k)-1)])
#define BWD(p, i, k) ((p)->bandwidth[((i)-1) * (p)->dims + ((k)-1)])
#define WGT(p, i) ((p)->weights[(i)-1])
#define LEFT(p, i) ((p)->left_child[(i)-1])
#define RIGHT(p, i) ((p)->right_child[(i)-1])
#define PERM(p, i) ((p)->permutation[(i)-1])

static int valid_index(const okde *p, int64_t ind) { /* src/BallTree01.jl:83 */
  return (0 < ind) && (ind <= 2 * p->num_points);
}
/* bwMin/bwMax(bd,i,k): index (i-1)*dims*multibandwidth + k; multibandwidth==0 always via the
 * typed constructors (src/BallTreeDensity01.jl:95-99,209-215) */
static double bw_min(const okde *p, int64_t i, int64_t k) {
  return p->bandwidthMin[(i - 1) * p->dims * p->multibandwidth + (k - 1)];
}
static double bw_max(const okde *p, int64_t i, int64_t k) {
  return p->bandwidthMax[(i - 1) * p->dims * p->multibandwidth + (k - 1)];
}

/* ---------------------------------------------------------------- tree build ------- */

/* swapDensity! + swapBall!  (src/BallTreeDensity01.jl:112-139, src/BallTree01.jl:109-138) */
static void swap_leaves(okde *p, int64_t i, int64_t j) {
  if (i == j) return;
  double t = WGT(p, i);
  WGT(p, i) = WGT(p, j);
  WGT(p, j) = t;
  int64_t q = PERM(p, i);
  PERM(p, i) = PERM(p, j);
  PERM(p, j) = q;
  for (int64_t k = 1; k <= p->dims; ++k) {
    t = CEN(p, i, k); CEN(p, i, k) = CEN(p, j, k); CEN(p, j, k) = t;
  }
  for (int64_t k = 1; k <= p->dims; ++k) {
    t = MEA(p, i, k); MEA(p, i, k) = MEA(p, j, k); MEA(p, j, k) = t;
    t = BWD(p, i, k); BWD(p, i, k) = BWD(p, j, k); BWD(p, j, k) = t;
    /* non-uniform branch (bandwidthMax/Min swap) unreachable: multibandwidth == 0 */
  }
}

/* most_spread_coord (src/BallTree01.jl:142-173): NB the strided range stops at leaf
 * high-1 (the last leaf is excluded) and the weight is 1/(high-low). */
static int64_t most_spread_coord(const okde *p, int64_t low, int64_t high) {
  double max_variance = 0.0;
  int64_t max_dim = 1;
  const int64_t d = p->dims;
  double w = 1.0 / (double)(high - low);
  for (int64_t dimension = 1; dimension <= d; ++dimension) {
    double mean = 0.0;
    for (int64_t pt = d * (low - 1) + dimension; pt <= d * (high - 1); pt += d)
      mean = mean + w * p->centers[pt - 1];
    double variance = 0.0;
    for (int64_t pt = d * (low - 1) + dimension; pt <= d * (high - 1); pt += d) {
      double df = p->centers[pt - 1] - mean;
      variance += df * df;
    }
    if (variance > max_variance) {
      max_variance = variance;
      max_dim = dimension;
    }
  }
  return max_dim;
}

/* select! (src/BallTree01.jl:223-242) */
static void select_leaves(okde *p, int64_t dimension, int64_t position, int64_t low, int64_t high) {
  while (low < high) {
    int64_t r = (low + high) / 2;
    swap_leaves(p, r, low);
    int64_t m = low;
    const int64_t lo0 = low, hi0 = high;
    for (int64_t i = lo0; i <= hi0; ++i) {
      if (CEN(p, i, dimension) - CEN(p, lo0, dimension) < 0.0) {
        m += 1;
        swap_leaves(p, m, i);
      }
    }
    swap_leaves(p, low, m);
    if (m <= position) low = m + 1;
    if (m >= position) high = m - 1;
  }
}

/* calcStatsDensity! (src/BallTreeDensity01.jl:141-187) which first runs calcStatsBall!
 * (src/BallTree01.jl:282-336, getMiniMaxi :249-278) */
static void calc_stats(okde *p, int64_t root) {
  const int64_t leftI = LEFT(p, root), rightI = RIGHT(p, root);
  if (!valid_index(p, leftI) || !valid_index(p, rightI)) return;
  for (int64_t d = 1; d <= p->dims; ++d) {
    double a = CEN(p, leftI, d) + RNG(p, leftI, d);
    double b = CEN(p, rightI, d) + RNG(p, rightI, d);
    double maxi = (a > b) ? a : b;
    double c = CEN(p, leftI, d) - RNG(p, leftI, d);
    double c2 = CEN(p, rightI, d) - RNG(p, rightI, d);
    double mini = (c < c2) ? c : c2;
    double halfspan = (maxi - mini) / 2.0;
    RNG(p, root, d) = halfspan;
    CEN(p, root, d) = mini + halfspan;
  }
  if (leftI != rightI)
    WGT(p, root) = WGT(p, leftI) + WGT(p, rightI);
  else
    WGT(p, root) = WGT(p, leftI);

  double wtL = WGT(p, leftI), wtR = WGT(p, rightI);
  double wtT = wtL + wtR + DBL_EPSILON; /* eps(Float64), :161 */
  wtL /= wtT;
  wtR /= wtT;
  for (int64_t k = 1; k <= p->dims; ++k) {
    double mL = MEA(p, leftI, k), mR = MEA(p, rightI, k);
    double m = wtL * mL + wtR * mR;
    MEA(p, root, k) = m;
    BWD(p, root, k) = wtL * (BWD(p, leftI, k) + mL * mL) + wtR * (BWD(p, rightI, k) + mR * mR) - m * m;
  }
}

/* buildBall! (src/BallTree01.jl:342-411) */
static void build_ball(okde *p, int64_t low, int64_t high, int64_t root) {
  if (low == high) { /* N == 1 special case */
    p->lowest_leaf[root - 1] = low;
    p->highest_leaf[root - 1] = high;
    LEFT(p, root) = low;
    RIGHT(p, root) = high;
    calc_stats(p, root);
    RIGHT(p, root) = NO_CHILD;
    return;
  }
  int64_t coord = most_spread_coord(p, low, high);
  int64_t split = (low + high) / 2;
  select_leaves(p, coord, split, low, high);
  int64_t left, right;
  if (split <= low) left = low; else { left = p->next; p->next += 1; }
  if (split + 1 >= high) right = high; else { right = p->next; p->next += 1; }
  p->lowest_leaf[root - 1] = low;
  p->highest_leaf[root - 1] = high;
  LEFT(p, root) = left;
  RIGHT(p, root) = right;
  if (left != low) build_ball(p, low, split, left);
  if (right != high) build_ball(p, split + 1, high, right);
  calc_stats(p, root);
}

/* makeBallTreeDensity + makeBallTree + buildTree!
 * (src/BallTreeDensity01.jl:192-231, src/BallTree01.jl:415-463) */
okde *okde_make(int64_t d, int64_t N, const double *points, const double *weights, const double *bwvar) {
  okde *p = (okde *)calloc(1, sizeof(okde));
  p->dims = d;
  p->num_points = N;
  p->multibandwidth = 0;
  p->centers = (double *)calloc((size_t)(2 * N * d + 1), sizeof(double));
  p->ranges = (double *)calloc((size_t)(2 * N * d + 1), sizeof(double));
  p->weights = (double *)calloc((size_t)(2 * N + 1), sizeof(double));
  p->means = (double *)calloc((size_t)(2 * N * d + 1), sizeof(double));
  p->bandwidth = (double *)calloc((size_t)(2 * N * d + 1), sizeof(double));
  p->bandwidthMin = (double *)calloc((size_t)(N * d + 1), sizeof(double));
  p->bandwidthMax = (double *)calloc((size_t)(N * d + 1), sizeof(double));
  p->left_child = (int64_t *)calloc((size_t)(2 * N + 1), sizeof(int64_t));
  p->right_child = (int64_t *)calloc((size_t)(2 * N + 1), sizeof(int64_t));
  p->lowest_leaf = (int64_t *)calloc((size_t)(2 * N + 1), sizeof(int64_t));
  p->highest_leaf = (int64_t *)calloc((size_t)(2 * N + 1), sizeof(int64_t));
  p->permutation = (int64_t *)calloc((size_t)(2 * N + 1), sizeof(int64_t));
  for (int64_t i = 0; i < 2 * N; ++i) { /* ones(Int,2Np) x4, zeros(Int,2Np) */
    p->left_child[i] = p->right_child[i] = p->lowest_leaf[i] = p->highest_leaf[i] = 1;
    p->permutation[i] = 0;
  }
  memcpy(p->centers + N * d, points, (size_t)(N * d) * sizeof(double));
  memcpy(p->means + N * d, points, (size_t)(N * d) * sizeof(double));
  memcpy(p->weights + N, weights, (size_t)N * sizeof(double));
  for (int64_t i = 0; i < N; ++i)
    for (int64_t k = 0; k < d; ++k) {
      p->bandwidth[(N + i) * d + k] = bwvar[k]; /* repeat(_bwMatrix, Np) */
      p->bandwidthMin[i * d + k] = bwvar[k];
      p->bandwidthMax[i * d + k] = bwvar[k];
    }
  if (N > 0) {
    int64_t i = N;
    for (int64_t j = 1; j <= N; ++j) { /* buildTree! :415-434 */
      for (int64_t k = 1; k <= d; ++k) p->ranges[i * d + k - 1] = 0;
      i += 1;
      p->lowest_leaf[i - 1] = i;
      p->highest_leaf[i - 1] = i;
      p->left_child[i - 1] = i;
      p->right_child[i - 1] = NO_CHILD;
      p->permutation[i - 1] = j;
    }
    p->next = 2;
    build_ball(p, N + 1, 2 * N, 1);
  }
  return p;
}

void okde_free(okde *p) {
  if (!p) return;
  free(p->centers); free(p->ranges); free(p->weights); free(p->means); free(p->bandwidth);
  free(p->bandwidthMin); free(p->bandwidthMax); free(p->left_child); free(p->right_child);
  free(p->lowest_leaf); free(p->highest_leaf); free(p->permutation);
  free(p);
}

/* kde!(points, ks, weights) (src/KDE01.jl:34-57): ks repeated if scalar, squared;
 * weights ./ sum(weights) */
okde *okde_kde_bw(int64_t d, int64_t N, const double *points, const double *ks, int64_t nks,
                  const double *weights) {
  double *var = (double *)malloc((size_t)d * sizeof(double));
  for (int64_t k = 0; k < d; ++k) {
    double s = (nks == 1) ? ks[0] : ks[k];
    var[k] = s * s;
  }
  double *w = (double *)malloc((size_t)(N > 0 ? N : 1) * sizeof(double));
  double sum = 0.0;
  for (int64_t i = 0; i < N; ++i) sum += weights ? weights[i] : 1.0;
  for (int64_t i = 0; i < N; ++i) w[i] = (weights ? weights[i] : 1.0) / sum;
  okde *p = okde_make(d, N, points, w, var);
  free(var);
  free(w);
  return p;
}

/* ---------------------------------------------------------------- getters ---------- */

int64_t okde_dims(const okde *p) { return p->dims; }
int64_t okde_npts(const okde *p) { return p->num_points; }
const double *okde_arr_f(const okde *p, int which) {
  switch (which) {
    case 0: return p->centers;
    case 1: return p->ranges;
    case 2: return p->weights;
    case 3: return p->means;
    case 4: return p->bandwidth;
    case 5: return p->bandwidthMin;
    default: return p->bandwidthMax;
  }
}
const int64_t *okde_arr_i(const okde *p, int which) {
  switch (which) {
    case 0: return p->left_child;
    case 1: return p->right_child;
    case 2: return p->lowest_leaf;
    case 3: return p->highest_leaf;
    default: return p->permutation;
  }
}

/* getPoints (src/KDE01.jl:91-101) */
void okde_get_points(const okde *p, double *out) {
  const int64_t N = p->num_points, d = p->dims;
  for (int64_t i = 1; i <= N; ++i) {
    int64_t o = PERM(p, N + i);
    for (int64_t k = 1; k <= d; ++k) out[(o - 1) * d + k - 1] = CEN(p, N + i, k);
  }
}
/* getBW (src/KDE01.jl:109-120) -- std-devs */
void okde_get_bw(const okde *p, double *out) {
  const int64_t N = p->num_points, d = p->dims;
  for (int64_t i = 1; i <= N; ++i) {
    int64_t o = PERM(p, N + i);
    for (int64_t k = 1; k <= d; ++k) out[(o - 1) * d + k - 1] = sqrt(BWD(p, N + i, k));
  }
}
/* getWeights (src/KDE01.jl:127-136) */
void okde_get_weights(const okde *p, double *out) {
  const int64_t N = p->num_points;
  for (int64_t i = 1; i <= N; ++i) out[PERM(p, N + i) - 1] = WGT(p, N + i);
}

/* marginal (src/KDE01.jl:143-153): sig = getBW(bd,[1]) (bandwidth of original point 1) */
okde *okde_marginal(const okde *p, const int64_t *ind, int64_t nind) {
  const int64_t N = p->num_points, d = p->dims;
  double *pts = (double *)malloc((size_t)(N * d) * sizeof(double));
  double *bw = (double *)malloc((size_t)(N * d) * sizeof(double));
  double *w = (double *)malloc((size_t)N * sizeof(double));
  okde_get_points(p, pts);
  okde_get_bw(p, bw);
  okde_get_weights(p, w);
  double *mp = (double *)malloc((size_t)(N * nind) * sizeof(double));
  double *ms = (double *)malloc((size_t)nind * sizeof(double));
  for (int64_t i = 0; i < N; ++i)
    for (int64_t a = 0; a < nind; ++a) mp[i * nind + a] = pts[i * d + ind[a] - 1];
  for (int64_t a = 0; a < nind; ++a) ms[a] = bw[ind[a] - 1];
  okde *m = okde_kde_bw(nind, N, mp, ms, nind, w);
  free(pts); free(bw); free(w); free(mp); free(ms);
  return m;
}

/* ---------------------------------------------------------------- evaluation ------- */

/* maxDistGauss! -> distGauss! at leaf x leaf (src/DualTree01.jl:14-47,51-57):
 * mainop = +, minmaxFnc = bwMin, minmaxFncUni = bwMax, saturate = false. */
static double max_dist_gauss(const okde *bd, int64_t dRoot, const okde *at, int64_t aRoot) {
  double acc = 0.0;
  for (int64_t k = 1; k <= at->dims; ++k) {
    double r = fabs(CEN(at, aRoot, k) - CEN(bd, dRoot, k));
    r = r + RNG(at, aRoot, k);
    r = r + RNG(bd, dRoot, k);
    acc += (r * r) / bw_min(bd, dRoot, k);
    if (bd->multibandwidth != 0) acc += log(bw_max(bd, dRoot, k));
  }
  return exp(-0.5 * acc);
}

/* evalDirect restricted to one query leaf j (src/DualTree01.jl:143-158) */
static void eval_direct_row(const okde *bd, const okde *at, int64_t j, int loo, double *pMin, double *pMax) {
  const int64_t first = bd->lowest_leaf[0], last = bd->highest_leaf[0];
  for (int64_t i = first; i <= last; ++i) {
    if (!loo || i != j) {
      double r = max_dist_gauss(bd, i, at, j);
      r *= WGT(bd, i);
      *pMin += r;
      *pMax += r;
    }
  }
}

/* norm of the top-level evaluate (src/DualTree01.jl:325-330) */
static double eval_norm(const okde *bd) {
  double norm = pow(2.0 * M_PI, (double)bd->dims / 2.0);
  if (bd->multibandwidth == 0)
    for (int64_t i = 1; i <= bd->dims; ++i) norm *= sqrt(bd->bandwidthMax[i - 1]);
  return norm;
}

/* evaluate(bd, locations, p, maxErr) with FORCE_EVAL_DIRECT = true
 * (src/DualTree01.jl:303-346 -> :248-299 -> evalDirect :130-162). loc == bd => LOO. */
int okde_evaluate(const okde *bd, const okde *loc, double *p) {
  if (bd->dims != loc->dims) return 1;
  const int loo = (bd == loc);
  const double norm = eval_norm(bd);
  const int64_t first = loc->lowest_leaf[0], last = loc->highest_leaf[0];
  for (int64_t j = first; j <= last; ++j) {
    double pMin = 0.0, pMax = 0.0;
    eval_direct_row(bd, loc, j, loo, &pMin, &pMax);
    if (loo)
      p[PERM(loc, j) - 1] = 0.5 * (pMin + pMax) / norm / (1.0 - WGT(bd, j));
    else
      p[PERM(loc, j) - 1] = 0.5 * (pMin + pMax) / norm;
  }
  return 0;
}

/* Rows [j0, j1) (0-based leaf positions, leaf order) of evaluate(bd, bd, ...) -- the same evalDirect sums as
 * okde_evaluate, restricted to a block of query leaves so that full-size configurations (1e5 x 1e5) can be
 * spot-checked in seconds.  out[j - j0] is the LOO density of leaf N+1+j; rows are independent (OpenMP). */
int okde_loo_rows(const okde *bd, int64_t j0, int64_t j1, double *out, int nthreads) {
  if (j0 < 0 || j1 > bd->num_points || j0 > j1) return 1;
  const double norm = eval_norm(bd);
  const int64_t first = bd->lowest_leaf[0];
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
#pragma omp parallel for schedule(static)
  for (int64_t r = j0; r < j1; ++r) {
    const int64_t j = first + r;
    double pMin = 0.0, pMax = 0.0;
    eval_direct_row(bd, bd, j, 1, &pMin, &pMax);
    out[r - j0] = 0.5 * (pMin + pMax) / norm / (1.0 - WGT(bd, j));
  }
  return 0;
}

/* evaluateDualTree(bd, pos::Matrix) (src/DualTree01.jl:370-390).  The reference builds a
 * throw-away tree over pos (bw 1, weights 1/M) whose only effect under brute force is the
 * order in which rows are visited; each row's sum runs over bd's leaves in leaf order, so
 * the values are identical when rows are visited in original order (done here). */
static void eval_point(const okde *bd, const double *x, double norm, double *out) {
  const int64_t first = bd->lowest_leaf[0], last = bd->highest_leaf[0];
  double pMin = 0.0, pMax = 0.0;
  for (int64_t i = first; i <= last; ++i) {
    double acc = 0.0;
    for (int64_t k = 1; k <= bd->dims; ++k) {
      double r = fabs(x[k - 1] - CEN(bd, i, k));
      r = r + 0.0;
      r = r + RNG(bd, i, k);
      acc += (r * r) / bw_min(bd, i, k);
    }
    double e = exp(-0.5 * acc);
    e *= WGT(bd, i);
    pMin += e;
    pMax += e;
  }
  *out = 0.5 * (pMin + pMax) / norm;
}

int okde_eval_points(const okde *bd, int64_t M, const double *pos, double *p) {
  const double norm = eval_norm(bd);
  for (int64_t j = 0; j < M; ++j) eval_point(bd, pos + j * bd->dims, norm, p + j);
  return 0;
}

int okde_eval_points_omp(const okde *bd, int64_t M, const double *pos, double *p, int nthreads) {
  const double norm = eval_norm(bd);
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < M; ++j) eval_point(bd, pos + j * bd->dims, norm, p + j);
  return 0;
}

/* evalAvgLogL (src/DualTree01.jl:450-470) */
double okde_eval_avg_logl(const okde *bd1, const okde *bd2) {
  const int64_t M = bd2->num_points;
  double *L = (double *)malloc((size_t)M * sizeof(double));
  double *W = (double *)malloc((size_t)M * sizeof(double));
  okde_evaluate(bd1, bd2, L);
  okde_get_weights(bd2, W);
  int bad = 0;
  for (int64_t j = 0; j < M; ++j)
    if (L[j] == 0.0 && W[j] != 0.0) bad = 1;
  double ll;
  if (bad) {
    ll = -INFINITY;
  } else {
    ll = 0.0;
    for (int64_t j = 0; j < M; ++j) {
      double l = (L[j] == 0.0) ? 1.0 : L[j];
      ll += log(l) * W[j];
    }
  }
  free(L);
  free(W);
  return ll;
}

/* entropy (src/DualTree01.jl:505-508): same object twice => leave-one-out */
double okde_entropy(const okde *bd) { return -okde_eval_avg_logl(bd, bd); }

/* ---------------------------------------------------------------- LOOCV ------------ */

/* updateBandwidth! (src/CrossValidation.jl:5-12) */
static void update_bandwidth(okde *bd, double factor, int divide) {
  const int64_t N = bd->num_points, d = bd->dims;
  for (int64_t i = 0; i < 2 * N * d; ++i)
    bd->bandwidth[i] = divide ? bd->bandwidth[i] / factor : bd->bandwidth[i] * factor;
  for (int64_t i = 0; i < N * d; ++i)
    bd->bandwidthMax[i] = bd->bandwidthMin[i] = bd->bandwidth[N * d + i];
}

/* nLOO_LL (src/CrossValidation.jl:15-24): multiply, evaluate, divide back (drifts by ulps) */
double okde_nloo_ll(double alpha, okde *bd) {
  alpha = alpha * alpha;
  update_bandwidth(bd, alpha, 0);
  double H = okde_entropy(bd);
  update_bandwidth(bd, alpha, 1);
  return H;
}

/* golden (src/CrossValidation.jl:44-98) */
static double golden(okde *bd, double ax, double bx, double cx, double tol, double *fmin, int64_t *ncalls) {
  const double C = (3.0 - sqrt(5.0)) / 2.0;
  const double R = 1.0 - C;
  double x0 = ax, x3 = cx, x1, x2;
  if (fabs(cx - bx) > fabs(bx - ax)) {
    x1 = bx;
    x2 = bx + C * (cx - bx);
  } else {
    x1 = bx - C * (bx - ax);
    x2 = bx;
  }
  double f1 = okde_nloo_ll(x1, bd);
  double f2 = okde_nloo_ll(x2, bd);
  int64_t n = 2;
  while (fabs(x3 - x0) > tol * (fabs(x1) + fabs(x2))) {
    if (f2 < f1) {
      x0 = x1;
      x1 = x2;
      x2 = R * x1 + C * x3;
      f1 = f2;
      f2 = okde_nloo_ll(x2, bd);
    } else {
      x3 = x2;
      x2 = x1;
      x1 = R * x2 + C * x0;
      f2 = f1;
      f1 = okde_nloo_ll(x1, bd);
    }
    n += 1;
  }
  if (ncalls) *ncalls += n;
  if (f1 < f2) { *fmin = f1; return x1; }
  *fmin = f2;
  return x2;
}

/* neighborMinMax (src/CrossValidation.jl:100-108): ranges of nodes 1..N-1 (min) and root (max) */
void okde_neighbor_minmax(const okde *bd, double *minm, double *maxm) {
  const int64_t N = bd->num_points, d = bd->dims;
  double s = 0.0;
  for (int64_t k = 1; k <= d; ++k) { double t = 2.0 * RNG(bd, 1, k); s += t * t; }
  *maxm = sqrt(s);
  double mn = INFINITY;
  for (int64_t i = 1; i <= N - 1; ++i) {
    s = 0.0;
    for (int64_t k = 1; k <= d; ++k) { double t = 2.0 * RNG(bd, i, k); s += t * t; }
    double v = sqrt(s);
    if (v < mn) mn = v;
  }
  *minm = (mn > 1e-6) ? mn : 1e-6;
}

/* ksize (src/CrossValidation.jl:110-120) */
okde *okde_ksize(const okde *bd, int64_t *ncalls) {
  const int64_t N = bd->num_points, d = bd->dims;
  double minm, maxm;
  okde_neighbor_minmax(bd, &minm, &maxm);
  double *pts = (double *)malloc((size_t)(N * d) * sizeof(double));
  double *w = (double *)malloc((size_t)N * sizeof(double));
  okde_get_points(bd, pts);
  okde_get_weights(bd, w);
  double k0 = (minm + maxm) / 2.0;
  okde *p = okde_kde_bw(d, N, pts, &k0, 1, w);
  double fmin;
  double ks = golden(p, 2.0 * minm / (minm + maxm), 1.0, 2.0 * maxm / (minm + maxm), 1e-2, &fmin, ncalls);
  ks = ks * (minm + maxm) / 2.0;
  okde_get_points(p, pts);
  okde_get_weights(p, w);
  okde *npd = okde_kde_bw(d, N, pts, &ks, 1, w);
  okde_free(p);
  free(pts);
  free(w);
  return npd;
}

/* kde!(points) (src/KDE01.jl:3-27): per-dimension LOOCV on the 1-D marginals */
okde *okde_kde_lcv(int64_t d, int64_t N, const double *points, int64_t *ncalls) {
  double one = 1.0;
  if (ncalls) *ncalls = 0;
  okde *p = okde_kde_bw(d, N, points, &one, 1, NULL);
  double *bwds = (double *)malloc((size_t)d * sizeof(double));
  double *bw = (double *)malloc((size_t)N * sizeof(double));
  for (int64_t i = 1; i <= d; ++i) {
    okde *m = okde_marginal(p, &i, 1);
    okde *pp = okde_ksize(m, ncalls);
    okde_get_bw(pp, bw);
    bwds[i - 1] = bw[0]; /* getBW(pp)[1] */
    okde_free(m);
    okde_free(pp);
  }
  okde_free(p);
  p = okde_kde_bw(d, N, points, bwds, d, NULL);
  free(bwds);
  free(bw);
  return p;
}

/* ---------------------------------------------------------------- MS Gibbs --------- */

/* GbGlb working state (src/MSGibbs01.jl:1-33), only the fields gibbs1 touches */
typedef struct {
  int64_t Ndim, Ndens, Nlevels;
  const okde *const *trees;
  double *particles, *variance; /* [Ndim x Ndens] column-major */
  double *p;
  int64_t *ind;
  double *Malmost, *Calmost;
  const double *randU, *randN;
  double *newPoints;
  int64_t *levelList, *levelListNew; /* [Ndens x maxNp] column-major */
  int64_t *dNpts;
  int64_t ruptr, rnptr, maxNp;
  double mn, vn;
  double *calclambdas, *calcmu;
  const uint8_t *mask; /* mask[j*Ndim + k], NULL = all active */
  int64_t *record;     /* labelsChoosen[s][j][level] flattened [Np][Ndens][Nlevels], or NULL */
} gbglb;

#define LL(g, j, z) ((g)->levelList[((z)-1) * (g)->Ndens + ((j)-1)])
#define LLN(g, j, z) ((g)->levelListNew[((z)-1) * (g)->Ndens + ((j)-1)])
#define PART(g, dim, j) ((g)->particles[((j)-1) * (g)->Ndim + ((dim)-1)])
#define VARI(g, dim, j) ((g)->variance[((j)-1) * (g)->Ndim + ((dim)-1)])

static int mask_on(const gbglb *g, int64_t j, int64_t dim) {
  return g->mask ? (g->mask[(j - 1) * g->Ndim + (dim - 1)] != 0) : 1;
}

/* updateGlbParticlesVariance! (src/MSGibbs01.jl:89-115); idx/level > 0 <=> recordChoosen (:109-112) */
static void update_particles_variance_rec(gbglb *g, int64_t j, int64_t idx, int64_t level);
static void update_particles_variance(gbglb *g, int64_t j) { update_particles_variance_rec(g, j, 0, 0); }
static void update_particles_variance_rec(gbglb *g, int64_t j, int64_t idx, int64_t level) {
  for (int64_t dim = 1; dim <= g->Ndim; ++dim) {
    if (!mask_on(g, j, dim)) {
      PART(g, dim, j) = 0.0;
      VARI(g, dim, j) = 0.0;
    } else {
      PART(g, dim, j) = MEA(g->trees[j - 1], g->ind[j - 1], dim);
      VARI(g, dim, j) = BWD(g->trees[j - 1], g->ind[j - 1], dim);
    }
  }
  if (g->record && idx > 0)
    g->record[((idx - 1) * g->Ndens + (j - 1)) * g->Nlevels + (level - 1)] = PERM(g->trees[j - 1], g->ind[j - 1]);
}

/* calcIndices! (src/MSGibbs01.jl:123-130) */
static void calc_indices(gbglb *g) {
  for (int64_t j = 1; j <= g->Ndens; ++j) update_particles_variance(g, j);
}

/* getEuclidLambda = sum(lambdas) (src/MSGibbs01.jl:141).  Julia's Base.sum over a Vector{Float64} is a plain
 * left-to-right loop for length < 16; from length 16 on it is `v = a[1] + a[2]; @simd for i in 3:n v += a[i]`
 * (Base.mapreduce_impl, block 1024), which LLVM may vectorise with reassociation: VF lanes x IC interleaved
 * accumulators consume VF*IC elements per trip while at least that many remain, the accumulators are folded
 * (interleave copies left to right, then a log2 shuffle tree inside the vector) and the tail is added
 * sequentially.  Whether the vector body runs at all for the 14 remaining elements of n = 16 depends on the
 * host ISA (VF*IC = 16 on AVX2 and 32 on AVX-512 => scalar, i.e. sequential; 8 on SSE2 => one vector trip).
 * okde_set_sum_simd(vf, ic) selects the emulated shape; (0, 0) = sequential (default, and what n < 16 always is).
 * KDEB200_MAX_DENS == 16, so M = 16 is the only reachable length on that edge (tests/test_oracle_stats.py). */
static int g_sum_vf = 0, g_sum_ic = 0;
void okde_set_sum_simd(int vf, int ic) {
  g_sum_vf = vf;
  g_sum_ic = ic;
}
static double julia_sum(const double *a, int64_t n) {
  if (n < 16 || g_sum_vf <= 0 || g_sum_ic <= 0 || g_sum_vf * g_sum_ic > 64) {
    double s = 0.0;
    for (int64_t j = 0; j < n; ++j) s += a[j];
    return s;
  }
  const int vf = g_sum_vf, ic = g_sum_ic, W = vf * ic;
  double acc[64];
  for (int l = 0; l < W; ++l) acc[l] = 0.0;
  acc[0] = a[0] + a[1]; /* the scalar start value enters lane 0 of the first accumulator */
  int64_t i = 2;
  for (; i + W <= n; i += W)
    for (int l = 0; l < W; ++l) acc[l] += a[i + l];
  double v[64];
  for (int l = 0; l < vf; ++l) { /* fold the interleaved copies: ((v0 + v1) + v2) + ... */
    double t = acc[l];
    for (int c = 1; c < ic; ++c) t = acc[c * vf + l] + t;
    v[l] = t;
  }
  for (int h = vf / 2; h >= 1; h /= 2) /* shuffle tree: upper half onto lower half */
    for (int l = 0; l < h; ++l) v[l] = v[l] + v[l + h];
  double s = v[0];
  for (; i < n; ++i) s += a[i];
  return s;
}

/* gaussianProductMeanCov! with getEuclidLambda / getEuclidMu
 * (src/MSGibbs01.jl:176-216, :141, :152-161) */
static void gaussian_product_mean_cov(gbglb *g, int64_t dim, double *destMu, double *destCov, int64_t skip) {
  *destMu = 0.0;
  *destCov = 0.0;
  int any = 0;
  for (int64_t j = 1; j <= g->Ndens; ++j)
    if (mask_on(g, j, dim) && !(skip > 0 && j == skip)) any = 1;
  if (!any) return;
  for (int64_t j = 1; j <= g->Ndens; ++j) {
    if (j != skip && mask_on(g, j, dim)) {
      g->calclambdas[j - 1] = 1.0 / VARI(g, dim, j);
      g->calcmu[j - 1] = PART(g, dim, j);
    } else {
      g->calclambdas[j - 1] = 0.0;
      g->calcmu[j - 1] = 0.0;
    }
  }
  *destCov = julia_sum(g->calclambdas, g->Ndens); /* sum(lambdas) */
  *destCov = 1.0 / *destCov;
  double lambdamu = 0.0;
  for (int64_t z = 0; z < g->Ndens; ++z) lambdamu += g->calcmu[z] * g->calclambdas[z];
  *destMu = (*destCov) * lambdamu;
}

/* makeFasterSampleIndex! (src/MSGibbs01.jl:250-328) */
static void make_faster_sample_index(int64_t j, gbglb *g, const double *muValue, const double *covValue,
                                     int64_t offset, int doCalmost) {
  const okde *tr = g->trees[j - 1];
  double pT = 0.0;
  int64_t zz = LL(g, j, 1);
  const int64_t n = g->dNpts[j - 1];
  for (int64_t z = 1; z <= n; ++z) {
    g->p[z - 1] = 0.0;
    for (int64_t i = 1; i <= g->Ndim; ++i) {
      int others = 0; /* dimmask: OR over all densities but j (:270-274) */
      for (int64_t jj = 1; jj <= g->Ndens; ++jj)
        if (jj != j && mask_on(g, jj, i)) others = 1;
      if (!mask_on(g, j, i) || !others) continue;
      double tmpC = BWD(tr, zz, i);
      if (doCalmost) tmpC += covValue[i - 1];
      double tmpM = MEA(tr, zz, i) - muValue[i - 1 + offset];
      double distr = (tmpM * tmpM) / tmpC;
      if (!isnan(distr)) {
        g->p[z - 1] += distr;
        g->p[z - 1] += log(tmpC);
      }
    }
    g->p[z - 1] = exp(-0.5 * g->p[z - 1]) * WGT(tr, zz);
    if (isnan(g->p[z - 1])) g->p[z - 1] = 0.0;
    pT += g->p[z - 1];
    if (z < n) zz = LL(g, j, z + 1);
  }
  if (pT < 1e-99) { /* :311-315 */
    double w = WGT(tr, zz);
    for (int64_t z = 0; z < n; ++z) g->p[z] = w;
    pT = 0.0;
    for (int64_t z = 0; z < n; ++z) pT += g->p[z];
  }
  for (int64_t z = 0; z < n; ++z) g->p[z] /= pT;
  for (int64_t z = 1; z < n; ++z) g->p[z] += g->p[z - 1];
}

/* selectLabelOnLevel (src/MSGibbs01.jl:330-351).  ruptr is the reference's 1-based index
 * (starts at 0 and is only dereferenced when dNp > 1). */
static void select_label_on_level(gbglb *g, int64_t j) {
  const int64_t dNp = g->dNpts[j - 1];
  int64_t z = 1;
  int64_t zz = LL(g, j, z);
  while (z <= dNp - 1) {
    if (g->randU[g->ruptr - 1] <= g->p[z - 1]) break;
    z += 1;
    if (z <= dNp) zz = LL(g, j, z);
  }
  g->ind[j - 1] = zz;
  g->ruptr += 1;
}

/* sampleIndices! (src/MSGibbs01.jl:364-385) */
static void sample_indices(gbglb *g, int64_t offset) {
  for (int64_t j = 1; j <= g->Ndens; ++j) {
    make_faster_sample_index(j, g, g->newPoints, NULL, offset, 0);
    select_label_on_level(g, j);
  }
  calc_indices(g);
}

/* sampleIndex (src/MSGibbs01.jl:404-429) */
static void sample_index(int64_t j, gbglb *g, int64_t idx, int64_t level) {
  for (int64_t i = 1; i <= g->Ndim; ++i)
    gaussian_product_mean_cov(g, i, &g->Malmost[i - 1], &g->Calmost[i - 1], j);
  make_faster_sample_index(j, g, g->Malmost, g->Calmost, 0, 1);
  select_label_on_level(g, j);
  update_particles_variance_rec(g, j, idx, level);
}

/* samplePoint! (src/MSGibbs01.jl:440-463) */
static void sample_point(gbglb *g, int64_t idx, int addEntropy) {
  for (int64_t dim = 1; dim <= g->Ndim; ++dim) {
    gaussian_product_mean_cov(g, dim, &g->mn, &g->vn, -1);
    g->rnptr += 1;
    if (addEntropy)
      g->newPoints[dim - 1 + idx] = g->mn + sqrt(g->vn) * g->randN[g->rnptr - 1];
    else
      g->newPoints[dim - 1 + idx] = g->mn;
  }
}

/* levelInit! (src/MSGibbs01.jl:467-475) */
static void level_init(gbglb *g) {
  for (int64_t j = 1; j <= g->Ndens; ++j) {
    g->dNpts[j - 1] = 1;
    LL(g, j, 1) = 1; /* root() */
  }
}

/* initIndices! (src/MSGibbs01.jl:477-497) */
static void init_indices(gbglb *g) {
  for (int64_t j = 1; j <= g->Ndens; ++j) {
    const int64_t dNp = g->dNpts[j - 1];
    int64_t zz = LL(g, j, 1);
    int64_t z = 1;
    while (z <= dNp) {
      g->p[z - 1] = WGT(g->trees[j - 1], zz);
      z += 1;
      if (z <= dNp) zz = LL(g, j, z);
    }
    for (z = 2; z <= dNp; ++z) g->p[z - 1] += g->p[z - 2];
    select_label_on_level(g, j);
  }
}

/* levelDown! (src/MSGibbs01.jl:500-523) */
static void level_down(gbglb *g) {
  for (int64_t j = 1; j <= g->Ndens; ++j) {
    const okde *tr = g->trees[j - 1];
    int64_t z = 1;
    for (int64_t y = 1; y <= g->dNpts[j - 1]; ++y) {
      int64_t cur = LL(g, j, y);
      if (valid_index(tr, LEFT(tr, cur))) { LLN(g, j, z) = LEFT(tr, cur); z += 1; }
      if (valid_index(tr, RIGHT(tr, cur))) { LLN(g, j, z) = RIGHT(tr, cur); z += 1; }
      if (g->ind[j - 1] == cur) g->ind[j - 1] = LLN(g, j, z - 1);
    }
    g->dNpts[j - 1] = z - 1;
  }
  int64_t *tmp = g->levelList;
  g->levelList = g->levelListNew;
  g->levelListNew = tmp;
}

/* glbs.Nlevels (src/MSGibbs01.jl:555-568): from the trees only */
int64_t okde_gibbs_nlevels(const okde *const *trees, int64_t ndens) {
  int64_t maxNp = 0;
  for (int64_t j = 0; j < ndens; ++j)
    if (maxNp < trees[j]->num_points) maxNp = trees[j]->num_points;
  return (int64_t)floor((log((double)maxNp) / log(2.0)) + 1.0);
}

/* gibbs1 (src/MSGibbs01.jl:527-629), Euclidean ops only.  Samples [s0,s1) of the full run:
 * the stream pointers advance by a constant per sample (Ndens*(1+L*(1+Niter)) uniforms,
 * Ndim*(L+1) normals -- verified against the sequential run in tests/), so starting at s0
 * just offsets them. */
int okde_gibbs_record(int64_t ndens, const okde *const *trees, int64_t Np, int64_t Niter, double *pts,
                      int64_t *ind, const double *randU, int64_t nU, const double *randN, int64_t nN,
                      int add_entropy, const uint8_t *mask, int64_t s0, int64_t s1, int64_t *record);

int okde_gibbs(int64_t ndens, const okde *const *trees, int64_t Np, int64_t Niter, double *pts, int64_t *ind,
               const double *randU, int64_t nU, const double *randN, int64_t nN, int add_entropy,
               const uint8_t *mask, int64_t s0, int64_t s1) {
  return okde_gibbs_record(ndens, trees, Np, Niter, pts, ind, randU, nU, randN, nN, add_entropy, mask, s0, s1, NULL);
}

/* record: labelsChoosen (glbs.recordChoosen = true), [Np][ndens][Nlevels], entries never written stay untouched */
int okde_gibbs_record(int64_t ndens, const okde *const *trees, int64_t Np, int64_t Niter, double *pts,
                      int64_t *ind, const double *randU, int64_t nU, const double *randN, int64_t nN,
                      int add_entropy, const uint8_t *mask, int64_t s0, int64_t s1, int64_t *record) {
  gbglb G;
  gbglb *g = &G;
  memset(g, 0, sizeof(G));
  g->Ndens = ndens;
  g->trees = trees;
  g->newPoints = pts;
  g->randU = randU;
  g->randN = randN;
  g->mask = mask;
  g->record = record;
  g->Ndim = 0;
  int64_t maxNp = 0;
  for (int64_t j = 0; j < ndens; ++j) {
    if (trees[j]->dims > g->Ndim) g->Ndim = trees[j]->dims;
    if (maxNp < trees[j]->num_points) maxNp = trees[j]->num_points;
  }
  for (int64_t j = 0; j < ndens; ++j)
    if (trees[j]->dims != g->Ndim) return 2;
  g->maxNp = maxNp;
  g->Nlevels = okde_gibbs_nlevels(trees, ndens);
  const int64_t perU = ndens * (1 + g->Nlevels * (1 + Niter));
  const int64_t perN = g->Ndim * (g->Nlevels + 1);
  if (s0 < 0 || s1 > Np || s0 > s1) return 3;
  if (s1 * perU > nU + 1 || s1 * perN > nN) return 4; /* the reference would throw BoundsError */

  g->ind = (int64_t *)malloc((size_t)ndens * sizeof(int64_t));
  g->p = (double *)calloc((size_t)maxNp, sizeof(double));
  g->Malmost = (double *)calloc((size_t)g->Ndim, sizeof(double));
  g->Calmost = (double *)calloc((size_t)g->Ndim, sizeof(double));
  g->calcmu = (double *)calloc((size_t)ndens, sizeof(double));
  g->calclambdas = (double *)calloc((size_t)ndens, sizeof(double));
  g->particles = (double *)calloc((size_t)(g->Ndim * ndens), sizeof(double));
  g->variance = (double *)calloc((size_t)(g->Ndim * ndens), sizeof(double));
  g->dNpts = (int64_t *)calloc((size_t)ndens, sizeof(int64_t));
  g->levelList = (int64_t *)malloc((size_t)(ndens * maxNp) * sizeof(int64_t));
  g->levelListNew = (int64_t *)malloc((size_t)(ndens * maxNp) * sizeof(int64_t));
  for (int64_t i = 0; i < ndens; ++i) g->ind[i] = 1;
  for (int64_t i = 0; i < ndens * maxNp; ++i) g->levelList[i] = g->levelListNew[i] = 1;
  g->ruptr = s0 * perU;
  g->rnptr = s0 * perN;

  for (int64_t s = s0 + 1; s <= s1; ++s) {
    const int64_t frm = (s - 1) * g->Ndim;
    level_init(g);
    init_indices(g);
    calc_indices(g);
    for (int64_t l = 1; l <= g->Nlevels; ++l) {
      sample_point(g, frm, 1); /* :594 omits the addEntropy flag => always noisy */
      level_down(g);
      sample_indices(g, frm);
      for (int64_t i = 1; i <= Niter; ++i)
        for (int64_t j = 1; j <= ndens; ++j) sample_index(j, g, s, l);
    }
    for (int64_t j = 1; j <= ndens; ++j) /* :612-616, the "+1" quirk */
      ind[(s - 1) * ndens + (j - 1)] = PERM(trees[j - 1], g->ind[j - 1]) + 1;
    sample_point(g, frm, add_entropy);
  }
  free(g->ind); free(g->p); free(g->Malmost); free(g->Calmost); free(g->calcmu); free(g->calclambdas);
  free(g->particles); free(g->variance); free(g->dNpts);
  free(g->levelList); free(g->levelListNew);
  return 0;
}

int okde_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* CPU-baseline helper: the same literal chain code, independent chains spread over threads. */
int okde_gibbs_omp(int64_t ndens, const okde *const *trees, int64_t Np, int64_t Niter, double *pts, int64_t *ind,
                   const double *randU, int64_t nU, const double *randN, int64_t nN, int add_entropy,
                   const uint8_t *mask, int nthreads) {
  int nt = nthreads > 0 ? nthreads : okde_max_threads();
  if (nt > Np) nt = (int)(Np > 0 ? Np : 1);
  int rc = 0;
#ifdef _OPENMP
  omp_set_num_threads(nt);
#endif
#pragma omp parallel for schedule(static, 1)
  for (int t = 0; t < nt; ++t) {
    int64_t a = Np * t / nt, b = Np * (t + 1) / nt;
    int r = okde_gibbs(ndens, trees, Np, Niter, pts, ind, randU, nU, randN, nN, add_entropy, mask, a, b);
    if (r) rc = r;
  }
  return rc;
}
