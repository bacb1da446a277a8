// eval_pruned.cu -- K6: error-bounded, tile-pruned N x M Gaussian-kernel sums -- the B200 counterpart of the
// reference's dual-tree evaluation.
//
// The reference's evaluate (src/DualTree01.jl:248-299 over recurseMinMax :164-242) walks a pair of ball trees
// and stops descending where the bounds on a node pair's contribution are within errTol.  On the GPU the same idea
// is applied one level, flat and exactly accounted:
//   * the component leaves, in ball-tree leaf order, are cut into the TMA tiles of eval.cu (TN consecutive leaves: a
//     spatially compact set); the query points are cut into blocks of BQ = 128 x Q points that are spatially compact
//     too -- leaves of the same tree for leave-one-out, otherwise the queries sorted by a Morton key on the device;
//   * every (block, tile) pair whose bounding boxes are further apart than CUT in the metric of the kernel
//     (0.5 sum_k gap_k^2 / var_k > PR_CUT) is dropped.  What a row loses is at most
//     sum_{dropped tiles} W_tile exp(-PR_CUT) <= exp(-PR_CUT) sum_i w_i = 1e-26 x total weight;
//   * every row whose kept sum is below 1e13 x that bound is recomputed over ALL components, so each result is either
//     within 1e-13 relative of the brute-force sum or IS the brute-force sum -- comfortably inside the 1e-12 parity bar
//     and 10 orders of magnitude tighter than the reference's default errTol = 1e-3.  Exact zeros and subnormals (the
//     likelihood's zero rule) go through the same exact_row fallback as eval.cu.
// The surviving pairs run through the arithmetic of eval.cu unchanged (same records, same exp, same leaf order inside
// a row), so the kernel is bound by the same FP64 pipe -- there is just less to do: bench.py reports the evaluated-pair
// fraction next to the nominal N x M rate.
#include <cub/device/device_radix_sort.cuh>

#include <cmath>
#include <vector>

#include "eval_shared.cuh"
#include "tree.cuh"

namespace kdeb200 {

constexpr double PR_CUT = 59.86721;       // ln(1e26): a dropped pair's kernel value is below 1e-26
constexpr double PR_DELTA = 1e-26;        // exp(-PR_CUT)
constexpr double PR_REL = 1e-13;          // guaranteed relative accuracy of a pruned row
constexpr int PR_Q = 2;                   // queries per thread => blocks of 256 points (tight boxes, many CTAs)
constexpr int PR_BQ = EV_THREADS * PR_Q;
constexpr int PR_BOUNDS_BLOCKS = 296;

// ---------------------------------------------------------------- boxes ------------------------------------
// one warp per group of `count` consecutive records (stride `stride` doubles, first d entries = coordinates):
// box[g] = [min_0.., max_0..], wsum[g] = sum of entry d (may be null)
__global__ void boxes_kernel(const double *__restrict__ rec, int stride, int d, int64_t n, int count, double *__restrict__ box,
                             double *__restrict__ wsum) {
  const int64_t g = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x & 31;
  const int64_t a = g * count;
  if (a >= n) return;
  const int64_t b = (a + count < n) ? a + count : n;
  double lo[KDEB200_MAX_DIM], hi[KDEB200_MAX_DIM], w = 0.0;
  for (int k = 0; k < d; ++k) {
    lo[k] = INFINITY;
    hi[k] = -INFINITY;
  }
  for (int64_t i = a + lane; i < b; i += 32) {
    const double *r = rec + i * stride;
    for (int k = 0; k < d; ++k) {
      lo[k] = fmin(lo[k], r[k]);
      hi[k] = fmax(hi[k], r[k]);
    }
    if (wsum) w += r[d];
  }
  for (int off = 16; off > 0; off >>= 1) {
    for (int k = 0; k < d; ++k) {
      lo[k] = fmin(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], off));
      hi[k] = fmax(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], off));
    }
    w += __shfl_xor_sync(0xffffffffu, w, off);
  }
  if (lane == 0) {
    for (int k = 0; k < d; ++k) {
      box[g * 2 * d + k] = lo[k];
      box[g * 2 * d + d + k] = hi[k];
    }
    if (wsum) wsum[g] = w;
  }
}

// ---------------------------------------------------------------- Morton order of free queries -------------
__device__ __forceinline__ uint64_t spread_bits(uint64_t v, int d, int bits) {  // bit i of v -> bit i*d
  uint64_t r = 0;
  for (int i = 0; i < bits; ++i) r |= ((v >> i) & 1ull) << (i * d);
  return r;
}
__global__ void morton_kernel(const double *__restrict__ pos, int d, int64_t M, const double *__restrict__ lohi, int bits,
                              uint64_t *__restrict__ keys, uint32_t *__restrict__ idx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  uint64_t key = 0;
  const double cells = (double)(1ull << bits);
  for (int k = 0; k < d; ++k) {
    const double lo = lohi[k], hi = lohi[d + k];
    double u = (hi > lo) ? (pos[i * d + k] - lo) / (hi - lo) : 0.0;
    u = fmin(fmax(u, 0.0), 0.999999999);  // NaN -> 0
    key |= spread_bits((uint64_t)(u * cells), d, bits) << k;
  }
  keys[i] = key;
  idx[i] = (uint32_t)i;
}
__global__ void gather_kernel(const double *__restrict__ pos, int d, int64_t M, const uint32_t *__restrict__ idx,
                              double *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const double *s = pos + (int64_t)idx[i] * d;
  for (int k = 0; k < d; ++k) out[i * d + k] = s[k];
}
// lo/hi of all queries in two steps: per-block partials part[block][2d], then one warp folds them into lohi
__global__ void bounds_kernel(const double *__restrict__ pos, int d, int64_t M, double *__restrict__ part) {
  __shared__ double sh[2 * KDEB200_MAX_DIM][8];
  double lo[KDEB200_MAX_DIM], hi[KDEB200_MAX_DIM];
  for (int k = 0; k < d; ++k) {
    lo[k] = INFINITY;
    hi[k] = -INFINITY;
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x)
    for (int k = 0; k < d; ++k) {
      const double v = pos[i * d + k];
      lo[k] = fmin(lo[k], v);
      hi[k] = fmax(hi[k], v);
    }
  for (int off = 16; off > 0; off >>= 1)
    for (int k = 0; k < d; ++k) {
      lo[k] = fmin(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], off));
      hi[k] = fmax(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], off));
    }
  const int w = threadIdx.x / 32, lane = threadIdx.x & 31;
  if (lane == 0)
    for (int k = 0; k < d; ++k) {
      sh[k][w] = lo[k];
      sh[d + k][w] = hi[k];
    }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nw = blockDim.x / 32;
    for (int k = 0; k < d; ++k) {
      double a = INFINITY, b = -INFINITY;
      for (int i = 0; i < nw; ++i) {
        a = fmin(a, sh[k][i]);
        b = fmax(b, sh[d + k][i]);
      }
      part[(int64_t)blockIdx.x * 2 * d + k] = a;
      part[(int64_t)blockIdx.x * 2 * d + d + k] = b;
    }
  }
}
__global__ void bounds_final_kernel(const double *__restrict__ part, int nblocks, int d, double *__restrict__ lohi) {
  const int k = threadIdx.x;
  if (k >= d) return;
  double a = INFINITY, b = -INFINITY;
  for (int i = 0; i < nblocks; ++i) {
    a = fmin(a, part[(int64_t)i * 2 * d + k]);
    b = fmax(b, part[(int64_t)i * 2 * d + d + k]);
  }
  lohi[k] = a;
  lohi[d + k] = b;
}

// ---------------------------------------------------------------- prune mask --------------------------------
// mask[b][t / 32] bit t % 32 = tile t can contribute to query block b.  One warp tests 32 tiles of one block.
struct MaskParams {
  const double *qbox;  // nqb x 2d
  const double *tbox;  // ntile x 2d
  uint32_t *mask;      // nqb x words
  unsigned long long *kept;  // total number of kept (block, tile) pairs (statistics)
  unsigned int *count;       // kept tiles per query block (longest-first launch order)
  int nqb, ntile, words, d;
  double cut;          // a pair is dropped when 0.5 sum gap^2 / var exceeds this (PR_CUT; the FP32 route uses a smaller one)
  int sym_bq, sym_tn;  // symmetric leave-one-out (sym_bq > 0): rows per block / nodes per tile; tiles that end at or before
                       // the block's first row are dropped (those pairs are produced by the transposed block)
  double half_ivar[KDEB200_MAX_DIM];  // 0.5 / variance_k
};
__global__ void iota_negate_kernel(const unsigned int *count, unsigned int *keys, uint32_t *idx, int n) {  // keys = ~count => ascending sort = longest first
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    keys[i] = ~count[i];
    idx[i] = (uint32_t)i;
  }
}
// the tile with ordinal k (0-based) among the set bits of a mask row; ntile when there are fewer
__device__ __forceinline__ int kth_tile(const uint32_t *__restrict__ row, int words, int ntile, int k) {
  for (int w = 0; w < words; ++w) {
    uint32_t bits = row[w];
    const int c = __popc(bits);
    if (k < c) {
      for (; k > 0; --k) bits &= bits - 1;
      return w * 32 + __ffs(bits) - 1;
    }
    k -= c;
  }
  return ntile;
}
__global__ void mask_kernel(const __grid_constant__ MaskParams P) {
  const int warp = threadIdx.x / 32, lane = threadIdx.x & 31;
  const int64_t unit = (int64_t)blockIdx.x * (blockDim.x / 32) + warp;
  if (unit >= (int64_t)P.nqb * P.words) return;
  const int b = (int)(unit / P.words), wd = (int)(unit % P.words);
  const int t = wd * 32 + lane;
  bool keep = false;
  if (t < P.ntile) {
    const double *q = P.qbox + (int64_t)b * 2 * P.d, *c = P.tbox + (int64_t)t * 2 * P.d;
    double acc = 0.0;
    for (int k = 0; k < P.d; ++k) {
      const double gap = fmax(0.0, fmax(q[k] - c[P.d + k], c[k] - q[P.d + k]));
      acc += gap * gap * P.half_ivar[k];
    }
    keep = !(acc > P.cut);  // NaN boxes are kept
    if (P.sym_bq > 0 && (int64_t)(t + 1) * P.sym_tn <= (int64_t)b * P.sym_bq) keep = false;
  }
  const unsigned bits = __ballot_sync(0xffffffffu, keep);
  if (lane == 0) {
    P.mask[(int64_t)b * P.words + wd] = bits;
    if (bits) {
      atomicAdd(P.kept, (unsigned long long)__popc(bits));
      atomicAdd(&P.count[b], (unsigned)__popc(bits));
    }
  }
}

// ---------------------------------------------------------------- main kernel -------------------------------
struct PrunedParams {
  EvalParams E;            // comps, queries, perm/out, exptab, ich, norm, N, M, q0, qstride, tile_nodes as in eval.cu
  const uint32_t *mask;    // nqb x words
  const uint32_t *qidx;    // sorted position -> original query index (free queries), or null
  const uint32_t *order;   // CTA -> query block, longest block first (the tail of the launch is made of short blocks)
  int words;
  double thresh;           // rows with a kept sum below this are recomputed over all components
  int64_t *redo;           // list of such rows (position in the block order) ...
  unsigned int *nredo;     // ... and its length
};

template <int D, bool LOO>
__global__ void __launch_bounds__(EV_THREADS) eval_pruned_kernel(const __grid_constant__ PrunedParams P) {
  constexpr int SE = Rec<D>::SE;
  constexpr int Q = PR_Q;
  constexpr int R = 2;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *tiles = reinterpret_cast<double *>(smem_raw);
  __shared__ __align__(16) double tab[KDE_EXP_TAB];
  __shared__ __align__(8) uint64_t bars[EV_STAGES];
  const EvalParams &E = P.E;
  const int tid = threadIdx.x;
  for (int i = tid; i < KDE_EXP_TAB; i += EV_THREADS) tab[i] = E.exptab[i];
  if (tid == 0) {
    for (int s = 0; s < EV_STAGES; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  const int TN = E.tile_nodes;
  const int ntile = (int)((E.N + TN - 1) / TN);
  const int blk = (int)P.order[blockIdx.x];
  const uint32_t *row = P.mask + (int64_t)blk * P.words;

  auto issue = [&](int t, int slot) {
    const int64_t a = (int64_t)t * TN;
    const int64_t cnt = (E.N - a < TN) ? (E.N - a) : TN;
    const uint32_t bytes = (uint32_t)(cnt * SE * sizeof(double));
    uint64_t *bar = &bars[slot % EV_STAGES];
    mbar_expect_tx(bar, bytes);
    tma_bulk_g2s(tiles + (size_t)(slot % EV_STAGES) * (EV_TILE_BYTES / 8), E.comps + a * SE, bytes, bar);
  };
  // producer state (thread 0): the tile sequence of this block, EV_STAGES ahead of the consumers
  int p_tile = 0, p_slot = 0;
  if (tid == 0) {
    p_tile = next_tile(row, P.words, ntile, 0);
    while (p_tile < ntile && p_slot < EV_STAGES) {
      issue(p_tile, p_slot);
      ++p_slot;
      p_tile = next_tile(row, P.words, ntile, p_tile + 1);
    }
  }

  const int64_t qbase = (int64_t)blk * PR_BQ;
  double x[Q][D], sum[Q];
  int64_t self[Q];
#pragma unroll
  for (int i = 0; i < Q; ++i) {
    int64_t qi = qbase + tid + (int64_t)i * EV_THREADS;
    if (qi >= E.M) qi = E.M - 1;
    const double *src = E.queries + (E.q0 + qi) * (int64_t)E.qstride;
#pragma unroll
    for (int k = 0; k < D; ++k) x[i][k] = src[k];
    sum[i] = 0.0;
    self[i] = LOO ? (E.q0 + qi) : -1;
  }
  const double *ich = E.ich;
  const int64_t qlo = E.q0 + qbase, qhi = qlo + PR_BQ;

  int slot = 0;
  for (int t = next_tile(row, P.words, ntile, 0); t < ntile; t = next_tile(row, P.words, ntile, t + 1), ++slot) {
    const int64_t a = (int64_t)t * TN;
    const int cnt = (int)((E.N - a < TN) ? (E.N - a) : TN);
    mbar_wait(&bars[slot % EV_STAGES], (uint32_t)((slot / EV_STAGES) & 1));
    const double *rec = tiles + (size_t)(slot % EV_STAGES) * (EV_TILE_BYTES / 8);
    const bool check = LOO && (a < qhi) && (a + cnt > qlo);
    if (!check) {
      int c = 0;
      for (; c + R <= cnt; c += R) {
        double rr[R][SE], e[R][Q];
#pragma unroll
        for (int r = 0; r < R; ++r) load_rec<SE>(rec + (c + r) * SE, rr[r]);
#pragma unroll
        for (int i = 0; i < Q; ++i) {
#pragma unroll
          for (int r = 0; r < R; ++r) e[r][i] = kde_exp_flush(quad<D>(x[i], rr[r], ich), tab, E.ec);
        }
#pragma unroll
        for (int i = 0; i < Q; ++i) {
#pragma unroll
          for (int r = 0; r < R; ++r) sum[i] = __fma_rn(e[r][i], rr[r][D], sum[i]);
        }
      }
      for (; c < cnt; ++c) {
        double ra[SE];
        load_rec<SE>(rec + c * SE, ra);
#pragma unroll
        for (int i = 0; i < Q; ++i) sum[i] = __fma_rn(kde_exp_flush(quad<D>(x[i], ra, ich), tab, E.ec), ra[D], sum[i]);
      }
    } else {
      for (int c = 0; c < cnt; ++c) {
        double ra[SE];
        load_rec<SE>(rec + c * SE, ra);
#pragma unroll
        for (int i = 0; i < Q; ++i) {
          const double e = kde_exp_flush(quad<D>(x[i], ra, ich), tab, E.ec);
          if (a + c != self[i]) sum[i] = __fma_rn(e, ra[D], sum[i]);
        }
      }
    }
    __syncthreads();
    if (tid == 0 && p_tile < ntile) {
      issue(p_tile, p_slot);
      ++p_slot;
      p_tile = next_tile(row, P.words, ntile, p_tile + 1);
    }
  }

#pragma unroll
  for (int i = 0; i < Q; ++i) {
    const int64_t qi = qbase + tid + (int64_t)i * EV_THREADS;
    if (qi >= E.M) continue;
    const double sv = sum[i];
    int64_t o = qi;
    if (LOO && E.perm) o = E.perm[E.q0 + qi];
    if (!LOO && P.qidx) o = P.qidx[qi];
    if (!(sv >= P.thresh)) {  // too small for the bound (or NaN): the row goes to the exact pass
      const unsigned k = atomicAdd(P.nredo, 1u);
      P.redo[k] = qi;
      continue;
    }
    double v = 0.5 * (sv + sv) / E.norm;
    if (LOO) v = v / (1.0 - E.comps[(E.q0 + qi) * SE + D]);
    E.out[o] = v;
  }
}

// exact pass: one CTA per listed row, all N components, block-strided partial sums folded in a fixed order; totals
// below EV_TINY go through exact_row (libdevice exp, leaf order) like eval.cu
__global__ void __launch_bounds__(256) redo_rows_kernel(const __grid_constant__ PrunedParams P, int d, int SE, int loo) {
  __shared__ double sh[256];
  __shared__ __align__(16) double tab[KDE_EXP_TAB];
  const EvalParams &E = P.E;
  for (int i = threadIdx.x; i < KDE_EXP_TAB; i += blockDim.x) tab[i] = E.exptab[i];
  __syncthreads();
  const unsigned n = *P.nredo;
  for (unsigned r = blockIdx.x; r < n; r += gridDim.x) {
    const int64_t qi = P.redo[r];
    const double *xq = E.queries + (E.q0 + qi) * (int64_t)E.qstride;
    const int64_t self = loo ? E.q0 + qi : -1;
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < E.N; i += blockDim.x) {
      if (i == self) continue;
      const double *c = E.comps + i * SE;
      double acc = 0.0;
      for (int k = 0; k < d; ++k) {
        const double df = __dadd_rn(xq[k], -c[k]);
        acc = __fma_rn(__dmul_rn(df, df), E.ich[k], acc);
      }
      s = __fma_rn(kde_exp_flush(acc, tab, E.ec), c[d], s);
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int off = blockDim.x / 2; off > 0; off >>= 1) {
      if ((int)threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      double sv = sh[0];
      if (sv < EV_TINY) sv = exact_row(E.comps, SE, d, E.N, xq, E.ich, self);
      double v = 0.5 * (sv + sv) / E.norm;
      if (loo) v = v / (1.0 - E.comps[(E.q0 + qi) * SE + d]);
      int64_t o = qi;
      if (loo && E.perm) o = E.perm[E.q0 + qi];
      if (!loo && P.qidx) o = P.qidx[qi];
      E.out[o] = v;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------- host side ---------------------------------
template <int D>
static cudaError_t launch_pruned_d(const PrunedParams &P, bool loo, unsigned grid, size_t smem, cudaStream_t st) {
  auto go = [&](auto kern) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    kern<<<grid, EV_THREADS, smem, st>>>(P);
    return cudaGetLastError();
  };
  return loo ? go(eval_pruned_kernel<D, true>) : go(eval_pruned_kernel<D, false>);
}

static cudaError_t launch_pruned(int d, const PrunedParams &P, bool loo, unsigned grid, size_t smem, cudaStream_t st) {
  switch (d) {
    case 1: return launch_pruned_d<1>(P, loo, grid, smem, st);
    case 2: return launch_pruned_d<2>(P, loo, grid, smem, st);
    case 3: return launch_pruned_d<3>(P, loo, grid, smem, st);
    case 4: return launch_pruned_d<4>(P, loo, grid, smem, st);
    case 5: return launch_pruned_d<5>(P, loo, grid, smem, st);
    case 6: return launch_pruned_d<6>(P, loo, grid, smem, st);
    case 7: return launch_pruned_d<7>(P, loo, grid, smem, st);
    case 8: return launch_pruned_d<8>(P, loo, grid, smem, st);
  }
  return cudaErrorInvalidValue;
}

// Is there anything to prune?  The kept window of a query has half-width sqrt(2 PR_CUT var_k) in dimension k; if that
// covers the components' whole extent in every dimension no (block, tile) pair can be dropped and the brute-force
// kernel (more queries per thread, no mask pass) is the faster exact route.
bool pruning_can_help(const kdeb200_tree_s *bd, const double *bw_var) {
  for (int k = 0; k < bd->d; ++k) {
    const double v = bw_var ? bw_var[k] : bd->hvar[k];
    if (!(v > 0.0)) return false;
    if (std::sqrt(2.0 * PR_CUT * v) < bd->extent[k]) return true;  // extent = max |x - root mean| ~ half the diameter
  }
  return false;
}

// statistics of the last pruned call on this context (for bench.py / tests)
struct PrunedStats {
  unsigned long long kept_pairs = 0, all_pairs = 0;
  unsigned int redo_rows = 0;
};
static PrunedStats g_stats[KDEB200_MAX_GPUS];
static unsigned long long *g_stat_kept[KDEB200_MAX_GPUS] = {nullptr};
static unsigned int *g_stat_redo[KDEB200_MAX_GPUS] = {nullptr};
static double g_stat_block_pairs[KDEB200_MAX_GPUS] = {0};

int pruned_last_stats(double *kept_fraction, int64_t *redo_rows) {
  Context &c = ctx();
  const int s = c.slot;
  if (g_stat_kept[s]) {  // read back lazily: the counters live in device memory so that the call itself stays asynchronous
    unsigned long long k = 0;
    unsigned int r = 0;
    KDE_CUDA(cudaMemcpy(&k, g_stat_kept[s], sizeof(k), cudaMemcpyDeviceToHost));
    KDE_CUDA(cudaMemcpy(&r, g_stat_redo[s], sizeof(r), cudaMemcpyDeviceToHost));
    g_stats[s].kept_pairs = k;
    g_stats[s].redo_rows = r;
  }
  if (kept_fraction) *kept_fraction = g_stat_block_pairs[s] > 0 ? (double)g_stats[s].kept_pairs / g_stat_block_pairs[s] : 1.0;
  if (redo_rows) *redo_rows = g_stats[s].redo_rows;
  return 0;
}

// tile boxes + weight sums of the components: cached with the tree (they do not depend on the bandwidth)
static int ensure_tile_boxes(kdeb200_tree_t bd, int TN, int ntile, int *launches) {
  if (bd->d_tilebox) return 0;
  Context &c = ctx();
  const int d = bd->d;
  const size_t bytes = sizeof(double) * ((size_t)ntile * 2 * d + ntile);
  KDE_CUDA(cudaMallocAsync(&bd->d_tilebox, bytes, c.stream));
  boxes_kernel<<<(unsigned)((ntile + 3) / 4), 128, 0, c.stream>>>(bd->d_leaf, bd->SE, d, bd->N, TN, bd->d_tilebox,
                                                                 bd->d_tilebox + (size_t)ntile * 2 * d);
  KDE_CUDA(cudaGetLastError());
  std::vector<double> wsum(ntile);
  KDE_CUDA(cudaMemcpyAsync(wsum.data(), bd->d_tilebox + (size_t)ntile * 2 * d, sizeof(double) * ntile, cudaMemcpyDeviceToHost, c.stream));
  KDE_CUDA(cudaStreamSynchronize(c.stream));
  double wt = 0.0;
  for (double w : wsum) wt += std::fabs(w);
  bd->wtotal = wt;
  if (launches) *launches += 1;
  return 0;
}

// statistics stay on the device (one small persistent buffer per context), read only when asked for
static int publish_stats(const unsigned long long *d_kept, const unsigned int *d_nredo, double block_pairs, cudaStream_t st) {
  const int s = ctx().slot;
  if (!g_stat_kept[s]) {
    char *sb = nullptr;
    KDE_CUDA(cudaMalloc(&sb, 64));
    g_stat_kept[s] = reinterpret_cast<unsigned long long *>(sb);
    g_stat_redo[s] = reinterpret_cast<unsigned int *>(sb + 16);
  }
  KDE_CUDA(cudaMemcpyAsync(g_stat_kept[s], d_kept, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
  KDE_CUDA(cudaMemcpyAsync(g_stat_redo[s], d_nredo, sizeof(unsigned int), cudaMemcpyDeviceToDevice, st));
  g_stat_block_pairs[s] = block_pairs;
  return 0;
}

// Same contract as eval_device (eval.cu): rows of bd's own leaves q0.. (loo) or the M free queries d_pos; d_out through
// perm when `scatter` (loo) / in the caller's query order (free queries).
int eval_pruned_device(kdeb200_tree_t bd, const double *d_pos, int64_t M, int loo, int64_t q0, bool scatter,
                       const double *bw_var, double *d_out, cudaStream_t st, int *launches) {
  Context &c = ctx();
  if (M <= 0) return 0;
  const int d = bd->d, SE = bd->SE;
  if (M >= (int64_t)1 << 31) KDE_FAIL(3, "eval (pruned): at most 2^31 - 1 query points per call");
  int TN = 1;
  while (TN * 2 * SE * 8 <= EV_TILE_BYTES) TN *= 2;
  const int ntile = (int)((bd->N + TN - 1) / TN);
  const int nqb = (int)((M + PR_BQ - 1) / PR_BQ);
  const int words = (ntile + 31) / 32;

  PrunedParams P;
  EvalParams &E = P.E;
  E.comps = bd->d_leaf;
  E.N = bd->N;
  E.M = M;
  E.q0 = loo ? q0 : 0;
  E.qstride = loo ? SE : d;
  E.perm = (loo && scatter) ? bd->d_perm : nullptr;
  E.out = d_out;
  E.partial = nullptr;
  E.exptab = c.d_exptab;
  E.ec = make_exp_consts();
  E.S = 1;
  E.chunk = 0;
  E.tile_nodes = TN;
  MaskParams MP;
  double norm = std::pow(2.0 * M_PI, (double)d / 2.0);
  for (int k = 0; k < d; ++k) {
    const double v = bw_var ? bw_var[k] : bd->hvar[k];
    if (!(v > 0.0) || !std::isfinite(v)) KDE_FAIL(5, "eval: bandwidth variance must be finite and > 0 (dim %d: %g)", k + 1, v);
    E.ich[k] = -0.5 / v;
    MP.half_ivar[k] = 0.5 / v;
    norm *= std::sqrt(v);
  }
  E.norm = norm;

  if (int rc = ensure_tile_boxes(bd, TN, ntile, launches)) return rc;
  P.thresh = PR_DELTA * bd->wtotal / PR_REL;

  // scratch: query boxes | mask | redo list | counters (| sorted queries, keys, indices, CUB temp for free queries)
  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t b_qbox = up(sizeof(double) * (size_t)nqb * 2 * d), b_mask = up(sizeof(uint32_t) * (size_t)nqb * words),
               b_redo = up(sizeof(int64_t) * (size_t)M), b_cnt = 256, b_ord = up(sizeof(uint32_t) * (size_t)nqb);
  size_t b_otmp = 0;
  {
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const unsigned int *)nullptr, (unsigned int *)nullptr, (const uint32_t *)nullptr,
                                    (uint32_t *)nullptr, nqb, 0, 32, st);
    b_otmp = up(tmp);
  }
  size_t b_sorted = 0, b_keys = 0, b_idx = 0, b_tmp = 0, b_lohi = 0;
  if (!loo) {
    b_sorted = up(sizeof(double) * (size_t)M * d);
    b_keys = up(sizeof(uint64_t) * (size_t)M);
    b_idx = up(sizeof(uint32_t) * (size_t)M);
    b_lohi = 256 + sizeof(double) * 2 * KDEB200_MAX_DIM * PR_BOUNDS_BLOCKS;
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const uint64_t *)nullptr, (uint64_t *)nullptr, (const uint32_t *)nullptr,
                                    (uint32_t *)nullptr, (int)M, 0, 64, st);
    b_tmp = up(tmp);
  }
  char *base = nullptr;
  KDE_CUDA(cudaMallocAsync(&base, b_qbox + b_mask + b_redo + b_cnt + 5 * b_ord + b_otmp + b_sorted + 2 * b_keys + 2 * b_idx + b_tmp + b_lohi, st));
  char *pp = base;
  auto take = [&](size_t b) { char *r = pp; pp += b; return r; };
  double *d_qbox = reinterpret_cast<double *>(take(b_qbox));
  uint32_t *d_mask = reinterpret_cast<uint32_t *>(take(b_mask));
  int64_t *d_redo = reinterpret_cast<int64_t *>(take(b_redo));
  char *d_cnt = take(b_cnt);
  unsigned long long *d_kept = reinterpret_cast<unsigned long long *>(d_cnt);
  unsigned int *d_nredo = reinterpret_cast<unsigned int *>(d_cnt + 16);
  unsigned int *c_in = reinterpret_cast<unsigned int *>(take(b_ord)), *c_out = reinterpret_cast<unsigned int *>(take(b_ord));
  uint32_t *o_in = reinterpret_cast<uint32_t *>(take(b_ord)), *o_out = reinterpret_cast<uint32_t *>(take(b_ord));
  unsigned int *c_key2 = reinterpret_cast<unsigned int *>(take(b_ord));
  void *d_otmp = take(b_otmp);
  KDE_CUDA(cudaMemsetAsync(d_cnt, 0, b_cnt, st));
  KDE_CUDA(cudaMemsetAsync(c_in, 0, b_ord, st));
  P.qidx = nullptr;
  if (loo) {
    E.queries = bd->d_leaf;
    boxes_kernel<<<(unsigned)((nqb + 3) / 4), 128, 0, st>>>(bd->d_leaf + E.q0 * SE, SE, d, M, PR_BQ, d_qbox, nullptr);
    KDE_CUDA(cudaGetLastError());
  } else {
    double *d_sorted = reinterpret_cast<double *>(take(b_sorted));
    uint64_t *k_in = reinterpret_cast<uint64_t *>(take(b_keys)), *k_out = reinterpret_cast<uint64_t *>(take(b_keys));
    uint32_t *i_in = reinterpret_cast<uint32_t *>(take(b_idx)), *i_out = reinterpret_cast<uint32_t *>(take(b_idx));
    void *d_tmp = take(b_tmp);
    double *d_lohi = reinterpret_cast<double *>(take(b_lohi));
    const int bits = 63 / d > 21 ? 21 : 63 / d;
    bounds_kernel<<<PR_BOUNDS_BLOCKS, 256, 0, st>>>(d_pos, d, M, d_lohi + 32);
    bounds_final_kernel<<<1, 32, 0, st>>>(d_lohi + 32, PR_BOUNDS_BLOCKS, d, d_lohi);
    morton_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(d_pos, d, M, d_lohi, bits, k_in, i_in);
    size_t tmp = b_tmp;
    cudaError_t ce = cub::DeviceRadixSort::SortPairs(d_tmp, tmp, k_in, k_out, i_in, i_out, (int)M, 0, bits * d, st);
    if (ce != cudaSuccess) {
      cudaFreeAsync(base, st);
      KDE_FAIL(100 + (int)ce, "eval (pruned): sorting the query points: %s", cudaGetErrorString(ce));
    }
    gather_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(d_pos, d, M, i_out, d_sorted);
    boxes_kernel<<<(unsigned)((nqb + 3) / 4), 128, 0, st>>>(d_sorted, d, d, M, PR_BQ, d_qbox, nullptr);
    KDE_CUDA(cudaGetLastError());
    E.queries = d_sorted;
    P.qidx = i_out;
    if (launches) *launches += 5;
  }
  MP.qbox = d_qbox;
  MP.tbox = bd->d_tilebox;
  MP.mask = d_mask;
  MP.kept = d_kept;
  MP.count = c_in;
  MP.nqb = nqb;
  MP.ntile = ntile;
  MP.words = words;
  MP.d = d;
  MP.cut = PR_CUT;
  MP.sym_bq = 0;
  MP.sym_tn = 0;
  const int64_t units = (int64_t)nqb * words;
  mask_kernel<<<(unsigned)((units + 3) / 4), 128, 0, st>>>(MP);
  KDE_CUDA(cudaGetLastError());
  {  // longest block first
    iota_negate_kernel<<<(unsigned)((nqb + 255) / 256), 256, 0, st>>>(c_in, c_out, o_in, nqb);
    size_t tmp = b_otmp;
    cudaError_t ce = cub::DeviceRadixSort::SortPairs(d_otmp, tmp, c_out, c_key2, o_in, o_out, nqb, 0, 32, st);
    if (ce != cudaSuccess) {
      cudaFreeAsync(base, st);
      KDE_FAIL(100 + (int)ce, "eval (pruned): ordering the query blocks: %s", cudaGetErrorString(ce));
    }
  }
  P.order = o_out;
  P.mask = d_mask;
  P.words = words;
  P.redo = d_redo;
  P.nredo = d_nredo;
  cudaError_t e = launch_pruned(d, P, loo != 0, (unsigned)nqb, (size_t)EV_STAGES * EV_TILE_BYTES, st);
  if (e != cudaSuccess) {
    cudaFreeAsync(base, st);
    KDE_FAIL(100 + (int)e, "eval (pruned) kernel launch: %s", cudaGetErrorString(e));
  }
  const unsigned redo_grid = (unsigned)(M < 4 * c.sm_count ? M : 4 * c.sm_count);
  redo_rows_kernel<<<redo_grid, 256, 0, st>>>(P, d, SE, loo);
  KDE_CUDA(cudaGetLastError());
  if (launches) *launches += 4;
  if (int rc = publish_stats(d_kept, d_nredo, (double)nqb * (double)ntile, st)) return rc;
  KDE_CUDA(cudaFreeAsync(base, st));
  return 0;
}


// ================================================================ symmetric leave-one-out ====================
// The LOO sums S_j = sum_{i != j} w_i K(x_i, x_j) of ALL rows of one density (nLOO_LL, src/CrossValidation.jl:15-24 ->
// evalDirect, src/DualTree01.jl:130-162) evaluate every unordered pair twice in the row-by-row form.  K is symmetric, so
// this kernel evaluates each pair {i, j}, i < j, ONCE and credits both rows: w_j K to row i's register accumulator and
// w_i K to column j.  A CTA owns a block of 128 x Q rows and walks the component tiles at or above its diagonal (and
// inside the pruning mask); per tile the column credits are reduced over the warp with shuffles, over the CTA through
// shared memory, and written to colpart[block][column]; a second kernel adds a row's own sum and its column credits in
// block order (deterministic) and applies the epilogue.  FP64 work per unordered pair: 3d + 9 instead of 2 (3d + 8);
// the order of additions differs from the reference's, which the 1e-12 tolerance of the likelihood allows
//  -- kdeb200_set_pruning(0) keeps the reference-order brute-force kernel.
constexpr int64_t SYM_MAX_N = 262144;  // colpart is (row blocks) x N doubles

struct SymParams {
  EvalParams E;          // comps = queries = leaf records; N = M
  const uint32_t *mask;  // nrb x words (upper triangle inside the pruning mask)
  const uint32_t *order; // row blocks, longest first
  const unsigned int *count;  // kept tiles per row block
  int words;
  int S;                 // CTAs per row block: CTA (b, s) takes the s-th share of the block's kept tiles
  int part, nparts;      // in-process multi-GPU: this device owns the row blocks b with b % nparts == part
  double *rowpart;       // S x N partial row sums (zeroed before the launch)
  double *colpart;       // nrb x N
};

template <int D, int Q>
__global__ void __launch_bounds__(EV_THREADS) loo_sym_kernel(const __grid_constant__ SymParams P) {
  constexpr int SE = Rec<D>::SE;
  constexpr int BQ = EV_THREADS * Q;
  constexpr int TNMAX = EV_TILE_BYTES / (SE * 8);
  constexpr int NW = EV_THREADS / 32;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *tiles = reinterpret_cast<double *>(smem_raw);
  double *cp = tiles + EV_STAGES * (EV_TILE_BYTES / 8);  // [NW][TNMAX] warp partials of the column credits
  __shared__ __align__(16) double tab[KDE_EXP_TAB];
  __shared__ __align__(8) uint64_t bars[EV_STAGES];
  const EvalParams &E = P.E;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < KDE_EXP_TAB; i += EV_THREADS) tab[i] = E.exptab[i];
  if (tid == 0) {
    for (int s = 0; s < EV_STAGES; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  const int TN = E.tile_nodes;
  const int ntile = (int)((E.N + TN - 1) / TN);
  const int blk = (int)P.order[blockIdx.x / P.S], share = (int)(blockIdx.x % P.S);
  if (blk % P.nparts != P.part) return;  // another GPU's row block
  const uint32_t *row = P.mask + (int64_t)blk * P.words;
  // this CTA's share of the block's kept tiles: ordinals [k0, k1)
  const int kept = (int)P.count[blk], per = (kept + P.S - 1) / P.S;
  const int k0 = share * per, k1 = (k0 + per < kept) ? k0 + per : kept;
  if (k0 >= k1) return;  // rowpart is pre-zeroed, colpart is only read where a tile was processed
  const int ntodo = k1 - k0;
  const int t_first = kth_tile(row, P.words, ntile, k0);

  auto issue = [&](int t, int slot) {
    const int64_t a = (int64_t)t * TN;
    const int64_t cnt = (E.N - a < TN) ? (E.N - a) : TN;
    const uint32_t bytes = (uint32_t)(cnt * SE * sizeof(double));
    uint64_t *bar = &bars[slot % EV_STAGES];
    mbar_expect_tx(bar, bytes);
    tma_bulk_g2s(tiles + (size_t)(slot % EV_STAGES) * (EV_TILE_BYTES / 8), E.comps + a * SE, bytes, bar);
  };
  int p_tile = t_first, p_slot = 0;
  if (tid == 0) {
    while (p_slot < ntodo && p_slot < EV_STAGES) {
      issue(p_tile, p_slot);
      ++p_slot;
      p_tile = next_tile(row, P.words, ntile, p_tile + 1);
    }
  }

  const int64_t qbase = (int64_t)blk * BQ;
  double x[Q][D], wq[Q], sum[Q];
  int64_t self[Q];
#pragma unroll
  for (int i = 0; i < Q; ++i) {
    const int64_t qi = qbase + tid + (int64_t)i * EV_THREADS;
    const bool ok = qi < E.N;
    const double *src = E.comps + (ok ? qi : E.N - 1) * (int64_t)SE;
#pragma unroll
    for (int k = 0; k < D; ++k) x[i][k] = src[k];
    wq[i] = ok ? src[D] : 0.0;  // rows past the end credit nothing
    sum[i] = 0.0;
    self[i] = ok ? qi : (int64_t)1 << 62;  // ... and never pass the "column above row" test
  }
  const double *ich = E.ich;
  const int64_t qhi = qbase + BQ;
  double *colout = P.colpart + (int64_t)blk * E.N;

  int slot = 0;
  for (int t = t_first; slot < ntodo; t = next_tile(row, P.words, ntile, t + 1), ++slot) {
    const int64_t a = (int64_t)t * TN;
    const int cnt = (int)((E.N - a < TN) ? (E.N - a) : TN);
    mbar_wait(&bars[slot % EV_STAGES], (uint32_t)((slot / EV_STAGES) & 1));
    const double *rec = tiles + (size_t)(slot % EV_STAGES) * (EV_TILE_BYTES / 8);
    const bool diag = a < qhi;  // the tile reaches into the block's own rows: only columns above the row count
    int c = 0;
    for (; c + 2 <= cnt; c += 2) {
      double r0[SE], r1[SE], e0[Q], e1[Q];
      load_rec<SE>(rec + c * SE, r0);
      load_rec<SE>(rec + (c + 1) * SE, r1);
#pragma unroll
      for (int i = 0; i < Q; ++i) {
        e0[i] = kde_exp_flush(quad<D>(x[i], r0, ich), tab, E.ec);
        e1[i] = kde_exp_flush(quad<D>(x[i], r1, ich), tab, E.ec);
      }
      if (diag) {
#pragma unroll
        for (int i = 0; i < Q; ++i) {
          if (!(a + c > self[i])) e0[i] = 0.0;
          if (!(a + c + 1 > self[i])) e1[i] = 0.0;
        }
      }
      double t0 = 0.0, t1 = 0.0;
#pragma unroll
      for (int i = 0; i < Q; ++i) {
        sum[i] = __fma_rn(e0[i], r0[D], sum[i]);
        sum[i] = __fma_rn(e1[i], r1[D], sum[i]);
        t0 = __fma_rn(e0[i], wq[i], t0);
        t1 = __fma_rn(e1[i], wq[i], t1);
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        t0 += __shfl_xor_sync(0xffffffffu, t0, off);
        t1 += __shfl_xor_sync(0xffffffffu, t1, off);
      }
      if (lane == 0) {
        cp[warp * TNMAX + c] = t0;
        cp[warp * TNMAX + c + 1] = t1;
      }
    }
    for (; c < cnt; ++c) {
      double r0[SE];
      load_rec<SE>(rec + c * SE, r0);
      double t0 = 0.0;
#pragma unroll
      for (int i = 0; i < Q; ++i) {
        double e = kde_exp_flush(quad<D>(x[i], r0, ich), tab, E.ec);
        if (diag && !(a + c > self[i])) e = 0.0;
        sum[i] = __fma_rn(e, r0[D], sum[i]);
        t0 = __fma_rn(e, wq[i], t0);
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) t0 += __shfl_xor_sync(0xffffffffu, t0, off);
      if (lane == 0) cp[warp * TNMAX + c] = t0;
    }
    __syncthreads();  // the stage is free and every warp's column partials are in shared memory
    if (tid == 0 && p_slot < ntodo) {
      issue(p_tile, p_slot);
      ++p_slot;
      p_tile = next_tile(row, P.words, ntile, p_tile + 1);
    }
    for (int cc = tid; cc < cnt; cc += EV_THREADS) {
      double v = cp[cc];
#pragma unroll
      for (int w = 1; w < NW; ++w) v += cp[w * TNMAX + cc];
      colout[a + cc] = v;
    }
    __syncthreads();  // cp is rewritten by the next tile
  }
#pragma unroll
  for (int i = 0; i < Q; ++i) {
    const int64_t qi = qbase + tid + (int64_t)i * EV_THREADS;
    if (qi < E.N) P.rowpart[(int64_t)share * E.N + qi] = sum[i];
  }
}

// S_j = rowsum[j] + sum over the row blocks b (in order) that credited column j; epilogue as eval.cu; rows below the
// bound of the pruning go to the exact pass
__global__ void loo_sym_finalize_kernel(const __grid_constant__ SymParams S, int bq, int SE, int d, double thresh, double *out,
                                        int64_t *redo, unsigned int *nredo) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const EvalParams &E = S.E;
  if (j >= E.N) return;
  const int t = (int)(j / E.tile_nodes);
  const int bmax = (int)(j / bq);  // blocks past j's own never see column j above their rows
  double s = 0.0;
  for (int y = 0; y < S.S; ++y) s += S.rowpart[(int64_t)y * E.N + j];
  for (int b = 0; b <= bmax; ++b)
    if ((S.mask[(int64_t)b * S.words + (t >> 5)] >> (t & 31)) & 1u) s += S.colpart[(int64_t)b * E.N + j];
  if (!(s >= thresh)) {
    const unsigned k = atomicAdd(nredo, 1u);
    redo[k] = j;
    return;
  }
  double v = 0.5 * (s + s) / E.norm;
  v = v / (1.0 - E.comps[j * SE + d]);
  out[j] = v;
}

// multi-GPU: what THIS device contributes to every row j -- its own rows' sums and the column credits of its row blocks
__global__ void loo_sym_partial_kernel(const __grid_constant__ SymParams S, int bq, double *tot) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const EvalParams &E = S.E;
  if (j >= E.N) return;
  const int t = (int)(j / E.tile_nodes);
  const int bmax = (int)(j / bq);
  double s = 0.0;
  for (int y = 0; y < S.S; ++y) s += S.rowpart[(int64_t)y * E.N + j];  // zero unless j is one of this device's rows
  for (int b = S.part; b <= bmax; b += S.nparts)
    if ((S.mask[(int64_t)b * S.words + (t >> 5)] >> (t & 31)) & 1u) s += S.colpart[(int64_t)b * E.N + j];
  tot[j] = s;
}
// ... and the primary folds the devices' contributions in device order, then the epilogue / exact-pass list as above
__global__ void loo_sym_combine_kernel(const __grid_constant__ EvalParams E, const double *tots, int nparts, int SE, int d,
                                       double thresh, double *out, int64_t *redo, unsigned int *nredo) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= E.N) return;
  double s = 0.0;
  for (int g = 0; g < nparts; ++g) s += tots[(int64_t)g * E.N + j];
  if (!(s >= thresh)) {
    const unsigned k = atomicAdd(nredo, 1u);
    redo[k] = j;
    return;
  }
  double v = 0.5 * (s + s) / E.norm;
  v = v / (1.0 - E.comps[j * SE + d]);
  out[j] = v;
}

template <int D, int Q>
static cudaError_t launch_sym_dq(const SymParams &P, unsigned grid, cudaStream_t st) {
  constexpr int SE = Rec<D>::SE;
  constexpr int TNMAX = EV_TILE_BYTES / (SE * 8);
  const size_t smem = (size_t)EV_STAGES * EV_TILE_BYTES + sizeof(double) * (EV_THREADS / 32) * TNMAX;
  auto kern = loo_sym_kernel<D, Q>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  kern<<<grid, EV_THREADS, smem, st>>>(P);
  return cudaGetLastError();
}
#ifndef SYM_Q1
#define SYM_Q1 8
#endif
constexpr int sym_q(int d) { return d == 1 ? SYM_Q1 : (d == 2 ? 4 : (d <= 6 ? 2 : 1)); }
static cudaError_t launch_sym(int d, const SymParams &P, unsigned grid, cudaStream_t st) {
  switch (d) {
    case 1: return launch_sym_dq<1, sym_q(1)>(P, grid, st);
    case 2: return launch_sym_dq<2, sym_q(2)>(P, grid, st);
    case 3: return launch_sym_dq<3, sym_q(3)>(P, grid, st);
    case 4: return launch_sym_dq<4, sym_q(4)>(P, grid, st);
    case 5: return launch_sym_dq<5, sym_q(5)>(P, grid, st);
    case 6: return launch_sym_dq<6, sym_q(6)>(P, grid, st);
    case 7: return launch_sym_dq<7, sym_q(7)>(P, grid, st);
    case 8: return launch_sym_dq<8, sym_q(8)>(P, grid, st);
  }
  return cudaErrorInvalidValue;
}

bool loo_sym_applicable(const kdeb200_tree_s *bd) { return bd->N >= 4096 && bd->N <= SYM_MAX_N; }

static int sym_eval_params(kdeb200_tree_t bd, const double *bw_var, EvalParams &E, double *half_ivar) {
  Context &c = ctx();
  const int d = bd->d, SE = bd->SE;
  int TN = 1;
  while (TN * 2 * SE * 8 <= EV_TILE_BYTES) TN *= 2;
  E.comps = bd->d_leaf;
  E.queries = bd->d_leaf;
  E.N = bd->N;
  E.M = bd->N;
  E.q0 = 0;
  E.qstride = SE;
  E.perm = nullptr;
  E.out = nullptr;
  E.partial = nullptr;
  E.exptab = c.d_exptab;
  E.ec = make_exp_consts();
  E.S = 1;
  E.chunk = 0;
  E.tile_nodes = TN;
  double norm = std::pow(2.0 * M_PI, (double)d / 2.0);
  for (int k = 0; k < d; ++k) {
    const double v = bw_var ? bw_var[k] : bd->hvar[k];
    if (!(v > 0.0) || !std::isfinite(v)) KDE_FAIL(5, "eval: bandwidth variance must be finite and > 0 (dim %d: %g)", k + 1, v);
    E.ich[k] = -0.5 / v;
    if (half_ivar) half_ivar[k] = 0.5 / v;
    norm *= std::sqrt(v);
  }
  E.norm = norm;
  return 0;
}

// multi-GPU, on the primary: fold the nparts contribution vectors (d_tots: nparts x N) into the LOO densities d_L
int loo_sym_combine_device(kdeb200_tree_t bd, const double *bw_var, const double *d_tots, int nparts, double *d_L,
                           cudaStream_t st, int *launches) {
  Context &c = ctx();
  const int d = bd->d, SE = bd->SE;
  const int64_t N = bd->N;
  EvalParams E;
  if (int rc = sym_eval_params(bd, bw_var, E, nullptr)) return rc;
  E.out = d_L;
  const double thresh = PR_DELTA * bd->wtotal / PR_REL;
  char *base = nullptr;
  const size_t b_redo = (sizeof(int64_t) * (size_t)N + 255) & ~(size_t)255;
  KDE_CUDA(cudaMallocAsync(&base, b_redo + 256, st));
  int64_t *d_redo = reinterpret_cast<int64_t *>(base);
  unsigned int *d_nredo = reinterpret_cast<unsigned int *>(base + b_redo);
  KDE_CUDA(cudaMemsetAsync(d_nredo, 0, 256, st));
  loo_sym_combine_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(E, d_tots, nparts, SE, d, thresh, d_L, d_redo, d_nredo);
  KDE_CUDA(cudaGetLastError());
  PrunedParams R;
  R.E = E;
  R.mask = nullptr;
  R.qidx = nullptr;
  R.order = nullptr;
  R.words = 0;
  R.thresh = thresh;
  R.redo = d_redo;
  R.nredo = d_nredo;
  const unsigned redo_grid = (unsigned)(N < 4 * c.sm_count ? N : 4 * c.sm_count);
  redo_rows_kernel<<<redo_grid, 256, 0, st>>>(R, d, SE, 1);
  KDE_CUDA(cudaGetLastError());
  if (launches) *launches += 2;
  KDE_CUDA(cudaFreeAsync(base, st));
  return 0;
}

// LOO densities of ALL leaves of bd, leaf order, into d_L (N doubles).  nparts > 1 (in-process multi-GPU): only the row
// blocks b % nparts == part are evaluated here and d_tot (N doubles) receives this device's contribution to every row;
// the primary adds the contributions up with loo_sym_combine_device.
int loo_sym_device(kdeb200_tree_t bd, const double *bw_var, double *d_L, cudaStream_t st, int *launches, int part,
                   int nparts, double *d_tot) {
  Context &c = ctx();
  const int d = bd->d, SE = bd->SE;
  const int64_t N = bd->N;
  int TN = 1;
  while (TN * 2 * SE * 8 <= EV_TILE_BYTES) TN *= 2;
  const int bq = EV_THREADS * sym_q(d);
  const int ntile = (int)((N + TN - 1) / TN), nrb = (int)((N + bq - 1) / bq), words = (ntile + 31) / 32;
  SymParams S;
  EvalParams &E = S.E;
  E.comps = bd->d_leaf;
  E.queries = bd->d_leaf;
  E.N = N;
  E.M = N;
  E.q0 = 0;
  E.qstride = SE;
  E.perm = nullptr;
  E.out = d_L;
  E.partial = nullptr;
  E.exptab = c.d_exptab;
  E.ec = make_exp_consts();
  E.S = 1;
  E.chunk = 0;
  E.tile_nodes = TN;
  MaskParams MP;
  double norm = std::pow(2.0 * M_PI, (double)d / 2.0);
  for (int k = 0; k < d; ++k) {
    const double v = bw_var ? bw_var[k] : bd->hvar[k];
    if (!(v > 0.0) || !std::isfinite(v)) KDE_FAIL(5, "eval: bandwidth variance must be finite and > 0 (dim %d: %g)", k + 1, v);
    E.ich[k] = -0.5 / v;
    MP.half_ivar[k] = 0.5 / v;
    norm *= std::sqrt(v);
  }
  E.norm = norm;
  if (int rc = ensure_tile_boxes(bd, TN, ntile, launches)) return rc;
  const double thresh = PR_DELTA * bd->wtotal / PR_REL;

  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t b_qbox = up(sizeof(double) * (size_t)nrb * 2 * d), b_mask = up(sizeof(uint32_t) * (size_t)nrb * words),
               b_redo = up(sizeof(int64_t) * (size_t)N), b_cnt = 256, b_ord = up(sizeof(uint32_t) * (size_t)nrb),
               b_col = up(sizeof(double) * (size_t)nrb * N);
  // CTAs per row block: the triangle makes the blocks' work differ by two orders of magnitude, and there are few
  // blocks (98 at N = 100k) -- split every block's kept tiles into S shares so that the launch is >= ~8 waves of
  // similar units, longest first
  const int nown = (nrb + nparts - 1) / nparts;
  int nsplit = (8 * 3 * c.sm_count + nown - 1) / nown;
  if (nsplit > (nparts > 1 ? 64 : 32)) nsplit = nparts > 1 ? 64 : 32;  // a device of a multi-GPU set owns few blocks
  if (nsplit < 1) nsplit = 1;
  const size_t b_row = up(sizeof(double) * (size_t)N * nsplit);
  size_t b_otmp = 0;
  {
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const unsigned int *)nullptr, (unsigned int *)nullptr, (const uint32_t *)nullptr,
                                    (uint32_t *)nullptr, nrb, 0, 32, st);
    b_otmp = up(tmp);
  }
  char *base = nullptr;
  KDE_CUDA(cudaMallocAsync(&base, b_qbox + b_mask + b_redo + b_cnt + 5 * b_ord + b_otmp + b_row + b_col, st));
  char *pp = base;
  auto take = [&](size_t b) { char *r = pp; pp += b; return r; };
  double *d_qbox = reinterpret_cast<double *>(take(b_qbox));
  uint32_t *d_mask = reinterpret_cast<uint32_t *>(take(b_mask));
  int64_t *d_redo = reinterpret_cast<int64_t *>(take(b_redo));
  char *d_cnt = take(b_cnt);
  unsigned int *c_in = reinterpret_cast<unsigned int *>(take(b_ord)), *c_out = reinterpret_cast<unsigned int *>(take(b_ord));
  uint32_t *o_in = reinterpret_cast<uint32_t *>(take(b_ord)), *o_out = reinterpret_cast<uint32_t *>(take(b_ord));
  unsigned int *c_key2 = reinterpret_cast<unsigned int *>(take(b_ord));
  void *d_otmp = take(b_otmp);
  double *d_row = reinterpret_cast<double *>(take(b_row));
  double *d_col = reinterpret_cast<double *>(take(b_col));
  unsigned long long *d_kept = reinterpret_cast<unsigned long long *>(d_cnt);
  unsigned int *d_nredo = reinterpret_cast<unsigned int *>(d_cnt + 16);
  KDE_CUDA(cudaMemsetAsync(d_cnt, 0, b_cnt, st));
  KDE_CUDA(cudaMemsetAsync(c_in, 0, b_ord, st));
  KDE_CUDA(cudaMemsetAsync(d_row, 0, b_row, st));
  boxes_kernel<<<(unsigned)((nrb + 3) / 4), 128, 0, st>>>(bd->d_leaf, SE, d, N, bq, d_qbox, nullptr);
  KDE_CUDA(cudaGetLastError());
  MP.qbox = d_qbox;
  MP.tbox = bd->d_tilebox;
  MP.mask = d_mask;
  MP.kept = d_kept;
  MP.count = c_in;
  MP.nqb = nrb;
  MP.ntile = ntile;
  MP.words = words;
  MP.d = d;
  MP.cut = PR_CUT;
  MP.sym_bq = bq;
  MP.sym_tn = TN;
  const int64_t units = (int64_t)nrb * words;
  mask_kernel<<<(unsigned)((units + 3) / 4), 128, 0, st>>>(MP);
  KDE_CUDA(cudaGetLastError());
  iota_negate_kernel<<<(unsigned)((nrb + 255) / 256), 256, 0, st>>>(c_in, c_out, o_in, nrb);
  {
    size_t tmp = b_otmp;
    cudaError_t ce = cub::DeviceRadixSort::SortPairs(d_otmp, tmp, c_out, c_key2, o_in, o_out, nrb, 0, 32, st);
    if (ce != cudaSuccess) {
      cudaFreeAsync(base, st);
      KDE_FAIL(100 + (int)ce, "loo (symmetric): ordering the row blocks: %s", cudaGetErrorString(ce));
    }
  }
  S.mask = d_mask;
  S.order = o_out;
  S.words = words;
  S.count = c_in;
  S.S = nsplit;
  S.part = part;
  S.nparts = nparts;
  S.rowpart = d_row;
  S.colpart = d_col;
  cudaError_t e = launch_sym(d, S, (unsigned)nrb * (unsigned)nsplit, st);
  if (e != cudaSuccess) {
    cudaFreeAsync(base, st);
    KDE_FAIL(100 + (int)e, "loo (symmetric) kernel launch: %s", cudaGetErrorString(e));
  }
  if (nparts > 1) {
    loo_sym_partial_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(S, bq, d_tot);
    KDE_CUDA(cudaGetLastError());
    if (launches) *launches += 5;
    KDE_CUDA(cudaFreeAsync(base, st));
    return 0;
  }
  loo_sym_finalize_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(S, bq, SE, d, thresh, d_L, d_redo, d_nredo);
  KDE_CUDA(cudaGetLastError());
  PrunedParams R;  // the exact pass of the pruned evaluation serves the rows below the bound
  R.E = E;
  R.mask = d_mask;
  R.qidx = nullptr;
  R.order = nullptr;
  R.words = words;
  R.thresh = thresh;
  R.redo = d_redo;
  R.nredo = d_nredo;
  const unsigned redo_grid = (unsigned)(N < 4 * c.sm_count ? N : 4 * c.sm_count);
  redo_rows_kernel<<<redo_grid, 256, 0, st>>>(R, d, SE, 1);
  KDE_CUDA(cudaGetLastError());
  if (launches) *launches += 6;
  if (int rc = publish_stats(d_kept, d_nredo, (double)nrb * (double)ntile, st)) return rc;
  KDE_CUDA(cudaFreeAsync(base, st));
  return 0;
}


// ================================================================ FP32, pruned ================================
// The FP32 evaluation (eval_f32.cu: packed FADD2 / FFMA2 + MUFU.EX2, contract 1e-5) through the same box-pair pruning.
// Its tiles hold component PAIRS, so it has its own tile boxes; its bound is looser -- a dropped pair's kernel value is
// below 1e-16, rows whose kept sum is below 1e10 x that go to the exact pass (FP64, all components) -- so the window is
// 8.6 instead of 10.9 bandwidths.  Free queries only (the FP32 LOO form has no row-range variant).
struct F32Prune {
  const uint32_t *mask = nullptr, *order = nullptr, *qidx = nullptr;
  int words = 0;
  double thresh = 0.0;
  int64_t *redo = nullptr;
  unsigned int *nredo = nullptr;
};
int f32_tile_pairs(int d);
int f32_queries_per_block();
int eval_device_f32_ex(kdeb200_tree_t bd, const double *d_pos, int64_t M, int loo, double *d_out, cudaStream_t st,
                       int *launches, const F32Prune *pr);
constexpr double PR_CUT32 = 36.84136;  // ln(1e16)
constexpr double PR_DELTA32 = 1e-16;
constexpr double PR_REL32 = 1e-6;

int eval_pruned_f32_device(kdeb200_tree_t bd, const double *d_pos, int64_t M, double *d_out, cudaStream_t st, int *launches) {
  Context &c = ctx();
  if (M <= 0) return 0;
  const int d = bd->d, SE = bd->SE;
  if (M >= (int64_t)1 << 31) KDE_FAIL(3, "eval (pruned): at most 2^31 - 1 query points per call");
  const int TC = 2 * f32_tile_pairs(d);  // components per FP32 tile
  const int BQ = f32_queries_per_block();
  const int ntile = (int)((bd->N + TC - 1) / TC), nqb = (int)((M + BQ - 1) / BQ), words = (ntile + 31) / 32;
  if (!bd->d_tilebox32) {  // boxes of the pair tiles (cached; the weight total comes from the FP64 boxes or from here)
    KDE_CUDA(cudaMallocAsync(&bd->d_tilebox32, sizeof(double) * ((size_t)ntile * 2 * d + ntile), c.stream));
    boxes_kernel<<<(unsigned)((ntile + 3) / 4), 128, 0, c.stream>>>(bd->d_leaf, SE, d, bd->N, TC, bd->d_tilebox32,
                                                                   bd->d_tilebox32 + (size_t)ntile * 2 * d);
    KDE_CUDA(cudaGetLastError());
    std::vector<double> wsum(ntile);
    KDE_CUDA(cudaMemcpyAsync(wsum.data(), bd->d_tilebox32 + (size_t)ntile * 2 * d, sizeof(double) * ntile, cudaMemcpyDeviceToHost, c.stream));
    KDE_CUDA(cudaStreamSynchronize(c.stream));
    double wt = 0.0;
    for (double w : wsum) wt += std::fabs(w);
    bd->wtotal = wt;
    if (launches) *launches += 1;
  }
  MaskParams MP;
  PrunedParams R;  // for the exact FP64 pass
  EvalParams &E = R.E;
  E.comps = bd->d_leaf;
  E.N = bd->N;
  E.M = M;
  E.q0 = 0;
  E.qstride = d;
  E.perm = nullptr;
  E.out = d_out;
  E.partial = nullptr;
  E.exptab = c.d_exptab;
  E.ec = make_exp_consts();
  E.S = 1;
  E.chunk = 0;
  E.tile_nodes = TC;
  double norm = std::pow(2.0 * M_PI, (double)d / 2.0);
  for (int k = 0; k < d; ++k) {
    const double v = bd->hvar[k];
    if (!(v > 0.0) || !std::isfinite(v)) KDE_FAIL(5, "eval: bandwidth variance must be finite and > 0");
    E.ich[k] = -0.5 / v;
    MP.half_ivar[k] = 0.5 / v;
    norm *= std::sqrt(v);
  }
  E.norm = norm;
  const double thresh = PR_DELTA32 * bd->wtotal / PR_REL32;

  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t b_qbox = up(sizeof(double) * (size_t)nqb * 2 * d), b_mask = up(sizeof(uint32_t) * (size_t)nqb * words),
               b_redo = up(sizeof(int64_t) * (size_t)M), b_cnt = 256, b_ord = up(sizeof(uint32_t) * (size_t)nqb),
               b_sorted = up(sizeof(double) * (size_t)M * d), b_keys = up(sizeof(uint64_t) * (size_t)M),
               b_idx = up(sizeof(uint32_t) * (size_t)M), b_lohi = 256 + sizeof(double) * 2 * KDEB200_MAX_DIM * PR_BOUNDS_BLOCKS;
  size_t b_otmp = 0, b_tmp = 0;
  {
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const unsigned int *)nullptr, (unsigned int *)nullptr, (const uint32_t *)nullptr,
                                    (uint32_t *)nullptr, nqb, 0, 32, st);
    b_otmp = up(tmp);
    tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const uint64_t *)nullptr, (uint64_t *)nullptr, (const uint32_t *)nullptr,
                                    (uint32_t *)nullptr, (int)M, 0, 64, st);
    b_tmp = up(tmp);
  }
  char *base = nullptr;
  KDE_CUDA(cudaMallocAsync(&base, b_qbox + b_mask + b_redo + b_cnt + 5 * b_ord + b_otmp + b_sorted + 2 * b_keys + 2 * b_idx + b_tmp + b_lohi, st));
  char *pp = base;
  auto take = [&](size_t b) { char *r = pp; pp += b; return r; };
  double *d_qbox = reinterpret_cast<double *>(take(b_qbox));
  uint32_t *d_mask = reinterpret_cast<uint32_t *>(take(b_mask));
  int64_t *d_redo = reinterpret_cast<int64_t *>(take(b_redo));
  char *d_cnt = take(b_cnt);
  unsigned int *c_in = reinterpret_cast<unsigned int *>(take(b_ord)), *c_out = reinterpret_cast<unsigned int *>(take(b_ord));
  uint32_t *o_in = reinterpret_cast<uint32_t *>(take(b_ord)), *o_out = reinterpret_cast<uint32_t *>(take(b_ord));
  unsigned int *c_key2 = reinterpret_cast<unsigned int *>(take(b_ord));
  void *d_otmp = take(b_otmp);
  double *d_sorted = reinterpret_cast<double *>(take(b_sorted));
  uint64_t *k_in = reinterpret_cast<uint64_t *>(take(b_keys)), *k_out = reinterpret_cast<uint64_t *>(take(b_keys));
  uint32_t *i_in = reinterpret_cast<uint32_t *>(take(b_idx)), *i_out = reinterpret_cast<uint32_t *>(take(b_idx));
  void *d_tmp = take(b_tmp);
  double *d_lohi = reinterpret_cast<double *>(take(b_lohi));
  unsigned long long *d_kept = reinterpret_cast<unsigned long long *>(d_cnt);
  unsigned int *d_nredo = reinterpret_cast<unsigned int *>(d_cnt + 16);
  KDE_CUDA(cudaMemsetAsync(d_cnt, 0, b_cnt, st));
  KDE_CUDA(cudaMemsetAsync(c_in, 0, b_ord, st));
  const int bits = 63 / d > 21 ? 21 : 63 / d;
  bounds_kernel<<<PR_BOUNDS_BLOCKS, 256, 0, st>>>(d_pos, d, M, d_lohi + 32);
  bounds_final_kernel<<<1, 32, 0, st>>>(d_lohi + 32, PR_BOUNDS_BLOCKS, d, d_lohi);
  morton_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(d_pos, d, M, d_lohi, bits, k_in, i_in);
  {
    size_t tmp = b_tmp;
    cudaError_t ce = cub::DeviceRadixSort::SortPairs(d_tmp, tmp, k_in, k_out, i_in, i_out, (int)M, 0, bits * d, st);
    if (ce != cudaSuccess) {
      cudaFreeAsync(base, st);
      KDE_FAIL(100 + (int)ce, "eval (pruned, FP32): sorting the query points: %s", cudaGetErrorString(ce));
    }
  }
  gather_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(d_pos, d, M, i_out, d_sorted);
  boxes_kernel<<<(unsigned)((nqb + 3) / 4), 128, 0, st>>>(d_sorted, d, d, M, BQ, d_qbox, nullptr);
  KDE_CUDA(cudaGetLastError());
  MP.qbox = d_qbox;
  MP.tbox = bd->d_tilebox32;
  MP.mask = d_mask;
  MP.kept = d_kept;
  MP.count = c_in;
  MP.nqb = nqb;
  MP.ntile = ntile;
  MP.words = words;
  MP.d = d;
  MP.cut = PR_CUT32;
  MP.sym_bq = 0;
  MP.sym_tn = 0;
  const int64_t units = (int64_t)nqb * words;
  mask_kernel<<<(unsigned)((units + 3) / 4), 128, 0, st>>>(MP);
  iota_negate_kernel<<<(unsigned)((nqb + 255) / 256), 256, 0, st>>>(c_in, c_out, o_in, nqb);
  {
    size_t tmp = b_otmp;
    cudaError_t ce = cub::DeviceRadixSort::SortPairs(d_otmp, tmp, c_out, c_key2, o_in, o_out, nqb, 0, 32, st);
    if (ce != cudaSuccess) {
      cudaFreeAsync(base, st);
      KDE_FAIL(100 + (int)ce, "eval (pruned, FP32): ordering the query blocks: %s", cudaGetErrorString(ce));
    }
  }
  F32Prune pr;
  pr.mask = d_mask;
  pr.order = o_out;
  pr.qidx = i_out;
  pr.words = words;
  pr.thresh = thresh;
  pr.redo = d_redo;
  pr.nredo = d_nredo;
  if (int rc = eval_device_f32_ex(bd, d_sorted, M, 0, d_out, st, launches, &pr)) {
    cudaFreeAsync(base, st);
    return rc;
  }
  E.queries = d_sorted;  // the exact pass reads the sorted FP64 queries and writes through the same index map
  R.mask = d_mask;
  R.qidx = i_out;
  R.order = nullptr;
  R.words = words;
  R.thresh = thresh;
  R.redo = d_redo;
  R.nredo = d_nredo;
  const unsigned redo_grid = (unsigned)(M < 4 * c.sm_count ? M : 4 * c.sm_count);
  redo_rows_kernel<<<redo_grid, 256, 0, st>>>(R, d, SE, 0);
  KDE_CUDA(cudaGetLastError());
  if (launches) *launches += 9;
  if (int rc = publish_stats(d_kept, d_nredo, (double)nqb * (double)ntile, st)) return rc;
  KDE_CUDA(cudaFreeAsync(base, st));
  return 0;
}

}  // namespace kdeb200
