// eval.cu -- K2: brute-force N x M Gaussian-kernel sums (evaluation, leave-one-out evaluation and
// the fused leave-one-out log-likelihood).
//
// Computes what evaluate(bd, locations, p, ...) src/DualTree01.jl:303-346 computes with
// FORCE_EVAL_DIRECT = true: evalDirect :130-162 over distGauss! :14-47 at leaf x leaf, then the
// normalisation :325-341; evalAvgLogL :450-470 / entropy :505-508 for the fused likelihood.
//
// Mapping: one thread owns Q = ev_q(d) query points (registers), the CTA streams the component records
// [x_0..x_{d-1}, w] (leaf order, the reference's summation order) through a 3-stage shared-memory
// ring filled by 1-D TMA bulk copies; every lane reads the same record (broadcast LDS), so the
// kernel is bound by the FP64 pipe (3d + 8 FP64 instructions per pair), not by memory.
#include <cmath>
#include <vector>

#include "eval_shared.cuh"
#include "tree.cuh"

namespace kdeb200 {

#ifndef EV_R
#define EV_R 2
#endif
template <int D, int Q, bool LOO>
__global__ void __launch_bounds__(EV_THREADS) eval_kernel(const __grid_constant__ EvalParams P) {
  constexpr int SE = Rec<D>::SE;
  constexpr int R = EV_R;  // component records per loop trip
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *tiles = reinterpret_cast<double *>(smem_raw);
  __shared__ __align__(16) double tab[KDE_EXP_TAB];
  __shared__ __align__(8) uint64_t bars[EV_STAGES];

  const int tid = threadIdx.x;
  for (int i = tid; i < KDE_EXP_TAB; i += EV_THREADS) tab[i] = P.exptab[i];
  if (tid == 0) {
    for (int s = 0; s < EV_STAGES; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  const int64_t c0 = (int64_t)blockIdx.y * P.chunk;
  const int64_t c1 = (c0 + P.chunk < P.N) ? c0 + P.chunk : P.N;
  const int TN = P.tile_nodes;
  const int ntiles = (int)((c1 - c0 + TN - 1) / TN);

  auto issue = [&](int t) {
    const int64_t a = c0 + (int64_t)t * TN;
    const int64_t cnt = (c1 - a < TN) ? (c1 - a) : TN;
    const uint32_t bytes = (uint32_t)(cnt * SE * sizeof(double));
    uint64_t *bar = &bars[t % EV_STAGES];
    mbar_expect_tx(bar, bytes);
    tma_bulk_g2s(tiles + (size_t)(t % EV_STAGES) * (EV_TILE_BYTES / 8), P.comps + a * SE, bytes, bar);
  };
  if (tid == 0)
    for (int t = 0; t < EV_STAGES && t < ntiles; ++t) issue(t);

  // this thread's queries
  const int64_t qbase = (int64_t)blockIdx.x * (EV_THREADS * Q);
  double x[Q][D];
  double sum[Q];
  int64_t self[Q];
#pragma unroll
  for (int i = 0; i < Q; ++i) {
    int64_t qi = qbase + tid + (int64_t)i * EV_THREADS;
    const bool ok = qi < P.M;
    if (!ok) qi = P.M - 1;
    const double *src = P.queries + (P.q0 + qi) * (int64_t)P.qstride;
#pragma unroll
    for (int k = 0; k < D; ++k) x[i][k] = src[k];
    sum[i] = 0.0;
    self[i] = LOO ? (P.q0 + qi) : -1;
  }
  const double *ich = P.ich;  // kernel-parameter space: used as c[0x0][..] operands, not registers

  // does the CTA's own index range overlap this split? (warp-uniform; only then test i == j)
  const int64_t qlo = P.q0 + qbase, qhi = qlo + EV_THREADS * Q;

  for (int t = 0; t < ntiles; ++t) {
    const int64_t a = c0 + (int64_t)t * TN;
    const int cnt = (int)((c1 - a < TN) ? (c1 - a) : TN);
    mbar_wait(&bars[t % EV_STAGES], (uint32_t)((t / EV_STAGES) & 1));
    const double *rec = tiles + (size_t)(t % EV_STAGES) * (EV_TILE_BYTES / 8);
    const bool check = LOO && (a < qhi) && (a + cnt > qlo);
    if (!check) {
      // branch-free exp for every pair, R records x Q queries per iteration (independent chains in one basic
      // block).  Exponents below -700 are clamped (terms <= 1e-304); rows whose total ends up below
      // EV_TINY are recomputed exactly at the end, so this never shows in a result.
      int c = 0;
      for (; c + R <= cnt; c += R) {
        double rr[R][SE], e[R][Q];
#pragma unroll
        for (int r = 0; r < R; ++r) load_rec<SE>(rec + (c + r) * SE, rr[r]);
#pragma unroll
        for (int i = 0; i < Q; ++i) {
#pragma unroll
          for (int r = 0; r < R; ++r) e[r][i] = kde_exp_flush(quad<D>(x[i], rr[r], ich), tab, P.ec);
        }
#pragma unroll
        for (int i = 0; i < Q; ++i) {
#pragma unroll
          for (int r = 0; r < R; ++r) sum[i] = __fma_rn(e[r][i], rr[r][D], sum[i]);  // component order
        }
      }
      for (; c < cnt; ++c) {
        double ra[SE];
        load_rec<SE>(rec + c * SE, ra);
#pragma unroll
        for (int i = 0; i < Q; ++i) sum[i] = __fma_rn(kde_exp_flush(quad<D>(x[i], ra, ich), tab, P.ec), ra[D], sum[i]);
      }
    } else {
      for (int c = 0; c < cnt; ++c) {
        double ra[SE];
        load_rec<SE>(rec + c * SE, ra);
#pragma unroll
        for (int i = 0; i < Q; ++i) {
          const double e = kde_exp_flush(quad<D>(x[i], ra, ich), tab, P.ec);
          if (a + c != self[i]) sum[i] = __fma_rn(e, ra[D], sum[i]);  // leave-one-out (src/DualTree01.jl:146)
        }
      }
    }
    __syncthreads();  // everyone is done with this stage
    if (tid == 0 && t + EV_STAGES < ntiles) issue(t + EV_STAGES);
  }

#pragma unroll
  for (int i = 0; i < Q; ++i) {
    const int64_t qi = qbase + tid + (int64_t)i * EV_THREADS;
    if (qi >= P.M) continue;
    if (P.S > 1) {
      P.partial[(int64_t)blockIdx.y * P.M + qi] = sum[i];
    } else {
      double sv = sum[i];
      if (sv < EV_TINY)
        sv = exact_row(P.comps, SE, D, P.N, P.queries + (P.q0 + qi) * (int64_t)P.qstride, P.ich, LOO ? P.q0 + qi : -1);
      double v = 0.5 * (sv + sv) / P.norm;  // 0.5*(pMin+pMax)/norm, :335-339
      if (LOO) v = v / (1.0 - P.comps[(P.q0 + qi) * SE + D]);
      const int64_t o = (LOO && P.perm) ? P.perm[P.q0 + qi] : qi;
      P.out[o] = v;
    }
  }
}

// sums the S partials of each query in split order and applies the epilogue
__global__ void eval_finalize_kernel(const double *partial, int S, int64_t M, int64_t q0, const double *comps, int SE,
                                     int D, int64_t N, const double *queries, int qstride, EvalParams P, double norm,
                                     int loo, const int64_t *perm, double *out) {
  const int64_t qi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= M) return;
  double s = 0.0;
  for (int y = 0; y < S; ++y) s += partial[(int64_t)y * M + qi];
  if (s < EV_TINY) s = exact_row(comps, SE, D, N, queries + (q0 + qi) * (int64_t)qstride, P.ich, loo ? q0 + qi : -1);
  double v = 0.5 * (s + s) / norm;
  if (loo) v = v / (1.0 - comps[(q0 + qi) * SE + D]);
  const int64_t o = (loo && perm) ? perm[q0 + qi] : qi;
  out[o] = v;
}

// sum_j W_j log L_j over leaf-ordered L (one CTA, fixed order => deterministic), with the
// reference's zero rule (src/DualTree01.jl:460-468): L_j == 0 && W_j != 0 => flag (-Inf).
__global__ void loglik_reduce_kernel(const double *L, const double *comps, int SE, int D, int64_t q0, int64_t n,
                                     double *sum_out, int *flag_out) {
  __shared__ double sh[1024];
  __shared__ int shf[1024];
  double s = 0.0;
  int f = 0;
  for (int64_t j = threadIdx.x; j < n; j += blockDim.x) loglik_term(L[j], comps[(q0 + j) * SE + D], s, f);
  loglik_block_reduce(s, f, sh, shf);
  if (threadIdx.x == 0) {
    *sum_out = sh[0];
    *flag_out = shf[0];
  }
}

// ---------------------------------------------------------------- host side --------------
// query points per thread: more rows per record load and per loop trip.  Swept on the B200 at 200k x 200k
// (tools/bench_eval_dims.py): d = 1: 6 beats 2 by 8 % and 8 by 4 %; d = 2: 4; d = 3, 4: 2 (3 and 4 are 1-11 % slower);
// d = 5, 6: 2 beats 1 by 2-10 %; d = 7, 8: 1.
#ifndef EV_Q1
#define EV_Q1 6
#endif
#ifndef EV_Q2
#define EV_Q2 4
#endif
#ifndef EV_Q4
#define EV_Q4 2
#endif
#ifndef EV_Q6
#define EV_Q6 2
#endif
constexpr int ev_q(int d) { return d == 1 ? EV_Q1 : (d == 2 ? EV_Q2 : (d <= 4 ? EV_Q4 : (d <= 6 ? EV_Q6 : 1))); }

template <int D, bool LOO>
static cudaError_t launch_eval_d(const EvalParams &P, dim3 grid, size_t smem, cudaStream_t st) {
  constexpr int Q = ev_q(D);
  auto kern = eval_kernel<D, Q, LOO>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  kern<<<grid, EV_THREADS, smem, st>>>(P);
  return cudaGetLastError();
}

template <bool LOO>
static cudaError_t launch_eval(int d, const EvalParams &P, dim3 grid, size_t smem, cudaStream_t st) {
  switch (d) {
    case 1: return launch_eval_d<1, LOO>(P, grid, smem, st);
    case 2: return launch_eval_d<2, LOO>(P, grid, smem, st);
    case 3: return launch_eval_d<3, LOO>(P, grid, smem, st);
    case 4: return launch_eval_d<4, LOO>(P, grid, smem, st);
    case 5: return launch_eval_d<5, LOO>(P, grid, smem, st);
    case 6: return launch_eval_d<6, LOO>(P, grid, smem, st);
    case 7: return launch_eval_d<7, LOO>(P, grid, smem, st);
    case 8: return launch_eval_d<8, LOO>(P, grid, smem, st);
  }
  return cudaErrorInvalidValue;
}

static int queries_per_cta(int d) { return EV_THREADS * ev_q(d); }

int eval_pruned_device(kdeb200_tree_t bd, const double *d_pos, int64_t M, int loo, int64_t q0, bool scatter,
                       const double *bw_var, double *d_out, cudaStream_t st, int *launches);
bool pruning_can_help(const kdeb200_tree_s *bd, const double *bw_var);
bool loo_sym_applicable(const kdeb200_tree_s *bd);
int loo_sym_device(kdeb200_tree_t bd, const double *bw_var, double *d_L, cudaStream_t st, int *launches, int part,
                   int nparts, double *d_tot);

// 0: brute force everywhere; 1 (default): the error-bounded tile-pruned kernel (eval_pruned.cu, <= 1e-13 relative)
// serves the leave-one-out LIKELIHOOD path (nLOO_LL / entropy / kde!(points)); 2: also plain FP64 evaluations --
// the counterpart of the reference's setForceEvalDirect!(false) (src/DualTree01.jl:3-9)
static int g_prune_mode = 1;
void set_prune_mode(int m) { g_prune_mode = m; }
int get_prune_mode() { return g_prune_mode; }

// Evaluate rows.  loo: rows are bd's own leaves q0..q0+M-1 (d_pos ignored); d_out is written
// through perm when `scatter`, else in row order.  bw_var overrides the tree's variances.
// prune: 0 = brute force, 1 = the pruned kernel where it can drop something, 2 = the pruned kernel always.
int eval_device(kdeb200_tree_t bd, const double *d_pos, int64_t M, int loo, int64_t q0, bool scatter,
                const double *bw_var, double *d_out, cudaStream_t st, int *launches, int prune) {
  Context &c = ctx();
  if (M <= 0) return 0;
  // the pruned kernel has one CTA per block of 256 queries (no component splits): it needs enough blocks to fill the chip
  if (prune == 2 || (prune == 1 && bd->N >= 4096 && M >= (int64_t)512 * c.sm_count && pruning_can_help(bd, bw_var)))
    return eval_pruned_device(bd, d_pos, M, loo, q0, scatter, bw_var, d_out, st, launches);
  const int d = bd->d;
  EvalParams P;
  P.comps = bd->d_leaf;
  P.N = bd->N;
  P.M = M;
  P.q0 = loo ? q0 : 0;
  P.queries = loo ? bd->d_leaf : d_pos;
  P.qstride = loo ? bd->SE : d;
  P.perm = (loo && scatter) ? bd->d_perm : nullptr;
  P.out = d_out;
  P.exptab = c.d_exptab;
  P.ec = make_exp_consts();
  double norm = std::pow(2.0 * M_PI, (double)d / 2.0);  // src/DualTree01.jl:325-330
  for (int k = 0; k < d; ++k) {
    const double v = bw_var ? bw_var[k] : bd->hvar[k];
    if (!(v > 0.0) || !std::isfinite(v)) KDE_FAIL(5, "eval: bandwidth variance must be finite and > 0 (dim %d: %g)", k + 1, v);
    P.ich[k] = -0.5 / v;
    norm *= std::sqrt(v);
  }
  P.norm = norm;
  const int SE = bd->SE;
  int TN = 1;
  while (TN * 2 * SE * 8 <= EV_TILE_BYTES) TN *= 2;
  P.tile_nodes = TN;
  const int bq = queries_per_cta(d);
  const int64_t nqb = (M + bq - 1) / bq;
  const int64_t ntile_total = (bd->N + TN - 1) / TN;
  int64_t S = (8LL * c.sm_count + nqb - 1) / nqb;
  if (S > ntile_total) S = ntile_total;
  if (S > 65535) S = 65535;
  if (S < 1) S = 1;
  int64_t chunk = ((ntile_total + S - 1) / S) * TN;
  S = (bd->N + chunk - 1) / chunk;
  P.S = (int)S;
  P.chunk = chunk;
  double *d_partial = nullptr;
  if (S > 1) KDE_CUDA(cudaMallocAsync(&d_partial, sizeof(double) * S * M, st));
  P.partial = d_partial;
  const size_t smem = EV_STAGES * EV_TILE_BYTES;
  dim3 grid((unsigned)nqb, (unsigned)S);
  cudaError_t e = loo ? launch_eval<true>(d, P, grid, smem, st) : launch_eval<false>(d, P, grid, smem, st);
  if (e != cudaSuccess) KDE_FAIL(100 + (int)e, "eval kernel launch: %s", cudaGetErrorString(e));
  if (launches) *launches += 1;
  if (S > 1) {
    eval_finalize_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(d_partial, (int)S, M, P.q0, bd->d_leaf, SE, d,
                                                                      bd->N, P.queries, P.qstride, P, norm, loo,
                                                                      P.perm, d_out);
    KDE_CUDA(cudaGetLastError());
    KDE_CUDA(cudaFreeAsync(d_partial, st));
    if (launches) *launches += 1;
  }
  return 0;
}

// sum_j W_j log L_j (+ zero flag) of leaf-ordered LOO densities d_L of rows q0..q0+n-1
int loglik_reduce_device(kdeb200_tree_t bd, const double *d_L, int64_t q0, int64_t n, double *d_sum, int *d_flag,
                         cudaStream_t st, int *launches) {
  loglik_reduce_kernel<<<1, 1024, 0, st>>>(d_L, bd->d_leaf, bd->SE, bd->d, q0, n, d_sum, d_flag);
  KDE_CUDA(cudaGetLastError());
  if (launches) *launches += 1;
  return 0;
}

// may the all-rows LOO sums of this density be shared out over the GPUs of the set by row block (symmetric kernel)?
bool loo_sym_shardable(const kdeb200_tree_s *bd) { return g_prune_mode >= 1 && loo_sym_applicable(bd); }

int loo_partial_device(kdeb200_tree_t bd, const double *bw_var, int64_t j0, int64_t j1, double *d_sum, int *d_flag,
                       cudaStream_t st, int *launches) {
  const int64_t n = j1 - j0;
  double *d_L = nullptr;
  KDE_CUDA(cudaMallocAsync(&d_L, sizeof(double) * (n > 0 ? n : 1), st));
  // all rows of one density on one device: each unordered pair once (symmetric kernel); row ranges (multi-GPU shards)
  // and small / very large densities: the row-by-row kernels
  int rc;
  if (g_prune_mode >= 1 && j0 == 0 && j1 == bd->N && loo_sym_applicable(bd))
    rc = loo_sym_device(bd, bw_var, d_L, st, launches, 0, 1, nullptr);
  else
    rc = eval_device(bd, nullptr, n, 1, j0, false, bw_var, d_L, st, launches, g_prune_mode >= 1 ? 1 : 0);
  if (rc) return rc;
  loglik_reduce_kernel<<<1, 1024, 0, st>>>(d_L, bd->d_leaf, bd->SE, bd->d, j0, n, d_sum, d_flag);
  KDE_CUDA(cudaGetLastError());
  if (launches) *launches += 1;
  KDE_CUDA(cudaFreeAsync(d_L, st));
  return 0;
}

}  // namespace kdeb200

namespace kdeb200 {
// FP32 (MUFU ex2) variant: defined in eval_f32.cu
}
