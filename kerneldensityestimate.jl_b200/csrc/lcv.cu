// lcv.cu -- kde!(points): per-dimension bandwidth selection by leave-one-out likelihood cross
// validation, natively behind one C call.
//
// Computes what the loop of kde!(points) src/KDE01.jl:13-23 computes for every dimension i:
//   pp = ksize(marginal(p, [i]));  bwds[i] = getBW(pp)[1]
// i.e. marginal :143-153 (1-D tree on points[i,:] with renormalised weights), neighborMinMax
// src/CrossValidation.jl:100-108, the working density of ksize :110-120, golden :44-98 over
// nLOO_LL :15-24 (multiply / evaluate / divide back, ulp drift included) and the final scaling.
//
// Two device paths, bit-identical to each other and to the per-call kdeb200_loo_entropy route:
//   N <= LCV_FUSED_MAX : ONE launch for all dimensions; CTA i runs the whole golden-section search of
//                        dimension i on chip (points in shared memory, one thread per row, the same
//                        sequential leaf-order sums and the same 1024-slot reduction tree as eval.cu)
//   larger N           : host golden loop over the tiled LOO kernel of eval.cu, one scalar back per step
#include <cmath>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#include "eval_shared.cuh"
#include "tree.cuh"

namespace kdeb200 {

int tree_build_host(int d, int64_t N, const double *points, const double *weights, const double *bw_var,
                    double *centers, double *ranges, double *wout, double *means, double *bandwidth,
                    int64_t *left, int64_t *right, int64_t *lowest, int64_t *highest, int64_t *perm);
int tree_create(int d, int64_t N, const double *means, const double *bandwidth, const double *weights,
                const int64_t *left, const int64_t *right, const int64_t *perm, bool gibbs_records,
                kdeb200_tree_t *out);
int tree_destroy(kdeb200_tree_t t);
int loo_partial_device(kdeb200_tree_t bd, const double *bw_var, int64_t j0, int64_t j1, double *d_sum, int *d_flag,
                       cudaStream_t st, int *launches);

constexpr int LCV_FUSED_MAX = 512;  // one row per thread and a single component tile: the sums of eval.cu at S = 1
constexpr int LCV_THREADS = 1024;   // == the block of loglik_reduce_kernel (same reduction tree)

struct LcvDim {
  double b0;      // leaf variance of the working density, ((minm + maxm) / 2)^2
  double ax, cx;  // golden bracket [2 minm / (minm + maxm), 1, 2 maxm / (minm + maxm)]
};

struct LcvParams {
  const double *leaf;  // [dim][N][2]: leaf-ordered (x, w) records of every dimension's working density
  const LcvDim *dims;
  double *out;         // [dim][3]: xmin, fmin, number of nLOO_LL calls
  const double *exptab;
  ExpConsts ec;
  int N;
  double norm0, tol, Cg, Rg;  // (2 pi)^(1/2); golden constants computed on the host
};

// nLOO_LL(alpha) for the CTA's dimension; every thread returns the same H and updates its copy of b identically
__device__ __forceinline__ double lcv_nloo(double alpha, double &b, const double *__restrict__ rec, int N,
                                           const double *__restrict__ tab, const ExpConsts &ec, double norm0,
                                           double *sh, int *shf) {
  const double a2 = alpha * alpha;  // src/CrossValidation.jl:17
  b = b * a2;                       // updateBandwidth!(bd, bd.bandwidth * alpha)
  const double ich = -0.5 / b;
  const double norm = norm0 * sqrt(b);
  const int j = threadIdx.x;
  double s = 0.0;
  int f = 0;
  if (j < N) {
    const double x = rec[2 * j], wj = rec[2 * j + 1];
    double sum = 0.0;
#pragma unroll 4
    for (int i = 0; i < N; ++i) {  // evalDirect, leaf order, i != j (src/DualTree01.jl:130-162)
      const double2 r = *reinterpret_cast<const double2 *>(rec + 2 * i);
      const double df = __dadd_rn(x, -r.x);
      const double e = kde_exp_flush(__fma_rn(__dmul_rn(df, df), ich, 0.0), tab, ec);
      if (i != j) sum = __fma_rn(e, r.y, sum);
    }
    if (sum < EV_TINY) sum = exact_row(rec, 2, 1, N, &x, &ich, j);
    double v = 0.5 * (sum + sum) / norm;
    v = v / (1.0 - wj);
    loglik_term(v, wj, s, f);
  }
  __syncthreads();  // everyone has consumed the previous H
  loglik_block_reduce(s, f, sh, shf);
  const double H = shf[0] ? INFINITY : -sh[0];  // entropy = -evalAvgLogL, -Inf under the zero rule
  b = b / a2;                                   // updateBandwidth!(bd, bd.bandwidth / alpha)
  return H;
}

__global__ void __launch_bounds__(LCV_THREADS) lcv_golden_kernel(const __grid_constant__ LcvParams P) {
  __shared__ __align__(16) double tab[KDE_EXP_TAB];
  __shared__ __align__(16) double rec[2 * LCV_FUSED_MAX];
  __shared__ double sh[LCV_THREADS];
  __shared__ int shf[LCV_THREADS];
  const int dim = blockIdx.x, N = P.N;
  for (int i = threadIdx.x; i < KDE_EXP_TAB; i += LCV_THREADS) tab[i] = P.exptab[i];
  for (int i = threadIdx.x; i < 2 * N; i += LCV_THREADS) rec[i] = P.leaf[(size_t)dim * 2 * N + i];
  __syncthreads();

  // golden (src/CrossValidation.jl:44-98), executed redundantly (and identically) by every thread
  const LcvDim D = P.dims[dim];
  double b = D.b0;
  const double ax = D.ax, bx = 1.0, cx = D.cx, Cg = P.Cg, Rg = P.Rg;
  double x0 = ax, x3 = cx, x1, x2;
  if (fabs(cx - bx) > fabs(bx - ax)) {
    x1 = bx;
    x2 = __dadd_rn(bx, __dmul_rn(Cg, cx - bx));  // explicit roundings: no FMA contraction, like the host loop
  } else {
    x1 = __dadd_rn(bx, -__dmul_rn(Cg, bx - ax));
    x2 = bx;
  }
  double f1 = lcv_nloo(x1, b, rec, N, tab, P.ec, P.norm0, sh, shf);
  double f2 = lcv_nloo(x2, b, rec, N, tab, P.ec, P.norm0, sh, shf);
  int n = 2;
  while (fabs(x3 - x0) > P.tol * (fabs(x1) + fabs(x2))) {
    if (f2 < f1) {
      x0 = x1;
      x1 = x2;
      x2 = __dadd_rn(__dmul_rn(Rg, x1), __dmul_rn(Cg, x3));
      f1 = f2;
      f2 = lcv_nloo(x2, b, rec, N, tab, P.ec, P.norm0, sh, shf);
    } else {
      x3 = x2;
      x2 = x1;
      x1 = __dadd_rn(__dmul_rn(Rg, x2), __dmul_rn(Cg, x0));
      f2 = f1;
      f1 = lcv_nloo(x1, b, rec, N, tab, P.ec, P.norm0, sh, shf);
    }
    ++n;
    if (n > 4096) break;  // NaN-proofing: the reference would spin forever
  }
  if (threadIdx.x == 0) {
    P.out[dim * 3 + 0] = (f1 < f2) ? x1 : x2;
    P.out[dim * 3 + 1] = (f1 < f2) ? f1 : f2;
    P.out[dim * 3 + 2] = (double)n;
  }
}

namespace {

// everything ksize needs from one dimension's marginal, in the reference's order of operations
struct Marginal {
  std::vector<double> leaf;  // N x (x, w2) in leaf order
  std::vector<double> means, weights;
  std::vector<int64_t> perm;
  double minm = 0, maxm = 0;
};

int build_marginal(int64_t N, const double *x, const std::vector<double> &w1, Marginal &m) {
  const int64_t NN = 2 * N;
  std::vector<double> centers(NN), ranges(NN), wout(NN), bandwidth(NN);
  std::vector<int64_t> left(NN), right(NN), lowest(NN), highest(NN);
  m.means.assign(NN, 0.0);
  m.perm.assign(NN, 0);
  const double one = 1.0;  // marginal(): getBW(bd, [1]) of kde!(points, [1.0])
  if (int rc = tree_build_host(1, N, x, w1.data(), &one, centers.data(), ranges.data(), wout.data(), m.means.data(),
                               bandwidth.data(), left.data(), right.data(), lowest.data(), highest.data(),
                               m.perm.data()))
    return rc;
  // neighborMinMax (src/CrossValidation.jl:100-108) in 1-D
  auto nrm = [&](int64_t i) { const double t = 2.0 * ranges[i]; return std::sqrt(t * t); };
  m.maxm = nrm(0);
  double mn = nrm(0);
  for (int64_t i = 1; i < N - 1; ++i) mn = std::fmin(mn, nrm(i));
  m.minm = std::fmax(mn, 1e-6);
  // ksize: p = kde!(getPoints(bd), [(minm+maxm)/2], getWeights(bd)) -- same points => same tree, the weights are
  // renormalised once more by kde! (sequential sum in original order)
  std::vector<double> w_orig(N);
  for (int64_t s = 0; s < N; ++s) w_orig[m.perm[N + s] - 1] = wout[N + s];
  double ssum = 0.0;
  for (int64_t i = 0; i < N; ++i) ssum += w_orig[i];
  m.weights.assign(NN, 0.0);
  m.leaf.resize(2 * N);
  for (int64_t s = 0; s < N; ++s) {
    const double w2 = w_orig[m.perm[N + s] - 1] / ssum;
    m.weights[N + s] = w2;
    m.leaf[2 * s] = centers[N + s];
    m.leaf[2 * s + 1] = w2;
  }
  return 0;
}

struct Golden {
  double Cg, Rg;
  Golden() {
    Cg = (3.0 - std::sqrt(5.0)) / 2.0;
    Rg = 1.0 - Cg;
  }
};

}  // namespace

// j0, j1 and allreduce: this process owns the leaf rows [j0, j1) of every nLOO_LL evaluation and the callback sums the
// partial likelihood (and ORs the zero flag) over the processes -- the multi-GPU split of SURVEY.md 8e.  allreduce ==
// nullptr: single process, all rows.  Small N runs the fused kernel redundantly on every process (no exchange needed).
int kde_lcv(int d, int64_t N, const double *points, int64_t j0, int64_t j1, kdeb200_allreduce_fn allreduce,
            void *user, double *bw_std_out, int *ncalls_out) {
  if (int rc = ensure_init()) return rc;
  if (d < 1) KDE_FAIL(3, "kde_lcv: d must be >= 1");
  if (N < 2) KDE_FAIL(3, "kde_lcv: at least two points are needed for cross validation");
  if (2 * N >= (int64_t)std::numeric_limits<int32_t>::max()) KDE_FAIL(3, "kde_lcv: N too large");
  if (!allreduce) {
    j0 = 0;
    j1 = N;
  }
  if (j0 < 0 || j1 > N || j0 > j1) KDE_FAIL(3, "kde_lcv: bad row range [%lld,%lld) of %lld", (long long)j0, (long long)j1, (long long)N);
  Context &c = ctx();
  const Golden G;
  const double tol = 1e-2;

  // kde!(points, [1.0]): weights ones(N) / N; marginal(): renormalised by their sequential sum
  std::vector<double> w1(N, 1.0 / (double)N);
  {
    double ssum = 0.0;
    for (int64_t i = 0; i < N; ++i) ssum += w1[i];
    for (int64_t i = 0; i < N; ++i) w1[i] = w1[i] / ssum;
  }
  // the d marginal trees are independent: one host thread each (the builder is a pure function of its arguments)
  std::vector<Marginal> margs(d);
  {
    std::vector<int> rcs(d, 0);
    auto work = [&](int k) {
      std::vector<double> x(N);
      for (int64_t i = 0; i < N; ++i) x[i] = points[i * d + k];
      rcs[k] = build_marginal(N, x.data(), w1, margs[k]);
    };
    if (d == 1 || N < 4096) {
      for (int k = 0; k < d; ++k) work(k);
    } else {
      std::vector<std::thread> th;
      for (int k = 0; k < d; ++k) th.emplace_back(work, k);
      for (auto &t : th) t.join();
    }
    for (int k = 0; k < d; ++k)
      if (rcs[k]) KDE_FAIL(rcs[k], "kde_lcv: building the marginal tree of dimension %d failed", k + 1);
  }
  auto finish = [&](int k, double xmin) {
    const Marginal &m = margs[k];
    const double ks = xmin * (m.minm + m.maxm) / 2.0;  // src/CrossValidation.jl:117
    bw_std_out[k] = std::sqrt(ks * ks);                // getBW of kde!(.., [ks]) (variance ks^2)
  };
  c.last_launches = 0;

  if (N <= LCV_FUSED_MAX) {
    std::vector<double> leaf((size_t)d * 2 * N);
    std::vector<LcvDim> dims(d);
    for (int k = 0; k < d; ++k) {
      const Marginal &m = margs[k];
      std::copy(m.leaf.begin(), m.leaf.end(), leaf.begin() + (size_t)k * 2 * N);
      const double h = (m.minm + m.maxm) / 2.0;
      dims[k].b0 = h * h;
      dims[k].ax = 2.0 * m.minm / (m.minm + m.maxm);
      dims[k].cx = 2.0 * m.maxm / (m.minm + m.maxm);
    }
    const size_t b_leaf = sizeof(double) * leaf.size(), b_dims = sizeof(LcvDim) * d, b_out = sizeof(double) * 3 * d;
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    char *base = nullptr;
    KDE_CUDA(cudaMallocAsync(&base, up(b_leaf) + up(b_dims) + up(b_out), c.stream));
    LcvParams P;
    P.leaf = reinterpret_cast<double *>(base);
    P.dims = reinterpret_cast<LcvDim *>(base + up(b_leaf));
    P.out = reinterpret_cast<double *>(base + up(b_leaf) + up(b_dims));
    P.exptab = c.d_exptab;
    P.ec = make_exp_consts();
    P.N = (int)N;
    P.norm0 = std::pow(2.0 * M_PI, 0.5);  // src/DualTree01.jl:325 with d = 1
    P.tol = tol;
    P.Cg = G.Cg;
    P.Rg = G.Rg;
    std::vector<double> out(3 * d);
    cudaError_t e = cudaMemcpyAsync(base, leaf.data(), b_leaf, cudaMemcpyHostToDevice, c.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(base + up(b_leaf), dims.data(), b_dims, cudaMemcpyHostToDevice, c.stream);
    if (e == cudaSuccess) {
      cudaEventRecord(c.ev0, c.stream);
      lcv_golden_kernel<<<d, LCV_THREADS, 0, c.stream>>>(P);
      e = cudaGetLastError();
      cudaEventRecord(c.ev1, c.stream);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out.data(), P.out, b_out, cudaMemcpyDeviceToHost, c.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c.stream);
    cudaFreeAsync(base, c.stream);
    if (e != cudaSuccess) KDE_FAIL(100 + (int)e, "kde_lcv: fused golden-section kernel: %s", cudaGetErrorString(e));
    float ms = 0;
    cudaEventElapsedTime(&ms, c.ev0, c.ev1);
    c.last_ms = ms;
    c.last_launches = 1;
    for (int k = 0; k < d; ++k) {
      finish(k, out[3 * k]);
      if (ncalls_out) ncalls_out[k] = (int)out[3 * k + 2];
    }
    return 0;
  }

  // large N: host golden loop, one tiled LOO launch sequence and one scalar back per step
  double *d_sum = nullptr;
  int *d_flag = nullptr;
  KDE_CUDA(cudaMallocAsync(&d_sum, 256, c.stream));
  d_flag = reinterpret_cast<int *>(d_sum + 8);
  int rc = 0, launches = 0;
  for (int k = 0; k < d && rc == 0; ++k) {
    Marginal &m = margs[k];
    const double h = (m.minm + m.maxm) / 2.0;
    double b = h * h;
    std::vector<double> bandwidth(2 * N, b);
    kdeb200_tree_t t = nullptr;
    rc = tree_create(1, N, m.means.data(), bandwidth.data(), m.weights.data(), nullptr, nullptr, m.perm.data(), false, &t);
    if (rc) break;
    int ncalls = 0;
    auto nloo = [&](double alpha, double &H) -> int {
      const double a2 = alpha * alpha;
      b = b * a2;
      double hs[9] = {0};
      if (j1 > j0) {
        if (int r = loo_partial_device(t, &b, j0, j1, d_sum, d_flag, c.stream, &launches)) return r;
        KDE_CUDA(cudaMemcpyAsync(hs, d_sum, sizeof(hs), cudaMemcpyDeviceToHost, c.stream));
        KDE_CUDA(cudaStreamSynchronize(c.stream));
      }
      int flag;
      std::memcpy(&flag, &hs[8], sizeof(int));
      if (allreduce) {
        if (int r = allreduce(&hs[0], &flag, user)) KDE_FAIL(9, "kde_lcv: the all-reduce callback failed (%d)", r);
      }
      H = flag ? std::numeric_limits<double>::infinity() : -hs[0];
      b = b / a2;
      ++ncalls;
      return 0;
    };
    const double ax = 2.0 * m.minm / (m.minm + m.maxm), bx = 1.0, cx = 2.0 * m.maxm / (m.minm + m.maxm);
    double x0 = ax, x3 = cx, x1, x2, f1 = 0, f2 = 0;
    if (std::fabs(cx - bx) > std::fabs(bx - ax)) {
      x1 = bx;
      x2 = bx + G.Cg * (cx - bx);
    } else {
      x1 = bx - G.Cg * (bx - ax);
      x2 = bx;
    }
    rc = nloo(x1, f1);
    if (!rc) rc = nloo(x2, f2);
    while (!rc && std::fabs(x3 - x0) > tol * (std::fabs(x1) + std::fabs(x2)) && ncalls <= 4096) {
      if (f2 < f1) {
        x0 = x1;
        x1 = x2;
        x2 = G.Rg * x1 + G.Cg * x3;
        f1 = f2;
        rc = nloo(x2, f2);
      } else {
        x3 = x2;
        x2 = x1;
        x1 = G.Rg * x2 + G.Cg * x0;
        f2 = f1;
        rc = nloo(x1, f1);
      }
    }
    tree_destroy(t);
    if (rc) break;
    finish(k, (f1 < f2) ? x1 : x2);
    if (ncalls_out) ncalls_out[k] = ncalls;
  }
  cudaFreeAsync(d_sum, c.stream);
  c.last_launches = launches;
  return rc;
}

}  // namespace kdeb200
