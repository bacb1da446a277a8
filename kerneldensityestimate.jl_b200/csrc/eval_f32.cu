// eval_f32.cu -- FP32 variant of K2 (tolerance 1e-5 relative, BASELINE.json north_star).
//
// Same sum as eval.cu (evalDirect, src/DualTree01.jl:130-162) with the pair arithmetic in FP32:
// coordinates are centred on the density's mean and pre-scaled by sqrt(0.5*log2(e)/variance_k) in
// FP64 once per tree.  Components are stored in PAIRS, [x_0(a), x_0(b), .., x_{d-1}(a), x_{d-1}(b),
// w(a), w(b)], so that one Blackwell packed-FP32 instruction (FADD2 / FFMA2, PTX add/fma.f32x2)
// serves two components: per pair and query d FADD2 + d FFMA2 + 2 MUFU.EX2 + 1 FFMA2, i.e. 3.5 + 1
// issue slots per evaluation at d = 3 against 8 MUFU cycles per warp -> MUFU-bound (16 ex2/clk/SM).
// Per-tile FP32 partial sums are folded into an FP64 accumulator, so the error does not grow
// with the number of components.
#include <cmath>
#include <vector>

#include "eval_shared.cuh"
#include "tree.cuh"

namespace kdeb200 {

constexpr int F32_THREADS = 128;
#ifndef F32_Q_N
#define F32_Q_N 3
#endif
#ifndef F32_UNR
#define F32_UNR 4
#endif
// swept on the B200 (tools/bench_f32.py, 500k x 500k): 3 queries per thread with 4 component pairs per loop trip
// beats 4 x 2 by 14 % at d = 3; 6 or 8 queries per thread lose 10-20 %
constexpr int F32_Q = F32_Q_N;  // query points per thread
constexpr int F32_UNROLL = F32_UNR;
constexpr int F32_STAGES = 3;
constexpr int F32_TILE_BYTES = 8192;

struct EvalF32Params {
  const float *comps;      // ceil(N/2) pair records, stride SP floats (16-byte multiple)
  const double *queries;   // raw FP64 coordinates, query i at queries + i*qstride
  const double *leafw;     // LOO: FP64 leaf records (for 1 - w_j), stride SE
  const int64_t *perm;
  double *out;
  int64_t N, M;
  int qstride, SE, tile_pairs, loo;
  double ctr[KDEB200_MAX_DIM], scl[KDEB200_MAX_DIM];
  double norm;
  // error-bounded pruned route (eval_pruned.cu): mask != null => only the tiles of the block's mask row are visited,
  // CTAs take blocks in `order`, outputs go through qidx (Morton-sorted free queries), rows whose kept sum is below
  // thresh are listed for the exact FP64 pass
  const uint32_t *mask, *order, *qidx;
  int words;
  double thresh;
  int64_t *redo;
  unsigned int *nredo;
};

template <int D>
struct F32Rec {
  static constexpr int SP = (2 * (D + 1) + 3) & ~3;  // floats per component pair
};

typedef unsigned long long f32x2;  // two packed floats {lo, hi}
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("sub.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float ex2_neg(float a) {  // 2^(-a): the negation is a free MUFU operand modifier
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-a));
  return e;
}

__global__ void prep_f32_kernel(const double *leaf, int SE, int D, int SP, int64_t N, const double *ctr_scl,
                                float *out) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // pair index
  if (2 * p >= N) return;
  for (int h = 0; h < 2; ++h) {
    const int64_t i = 2 * p + h;
    const bool ok = i < N;  // an odd N is padded with a zero-weight component
    for (int k = 0; k < D; ++k) out[p * SP + 2 * k + h] = ok ? (float)((leaf[i * SE + k] - ctr_scl[k]) * ctr_scl[8 + k]) : 0.f;
    out[p * SP + 2 * D + h] = ok ? (float)leaf[i * SE + D] : 0.f;
  }
  for (int k = 2 * (D + 1); k < SP; ++k) out[p * SP + k] = 0.f;
}

template <int D, bool LOO, bool PRUNED>
__global__ void __launch_bounds__(F32_THREADS) eval_f32_kernel(const __grid_constant__ EvalF32Params P) {
  constexpr int SP = F32Rec<D>::SP;
  constexpr int Q = F32_Q;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float *tiles = reinterpret_cast<float *>(smem_raw);
  __shared__ __align__(8) uint64_t bars[F32_STAGES];
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < F32_STAGES; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  const int TP = P.tile_pairs;
  const int64_t npairs = (P.N + 1) / 2;
  const int ntiles = (int)((npairs + TP - 1) / TP);
  const int blk = PRUNED ? (int)P.order[blockIdx.x] : (int)blockIdx.x;
  const uint32_t *row = PRUNED ? P.mask + (int64_t)blk * P.words : nullptr;
  auto nxt = [&](int pos) { return PRUNED ? next_tile(row, P.words, ntiles, pos) : (pos < ntiles ? pos : ntiles); };
  auto issue = [&](int t, int slot) {
    const int64_t a = (int64_t)t * TP;
    const int64_t cnt = (npairs - a < TP) ? (npairs - a) : TP;
    const uint32_t bytes = (uint32_t)(cnt * SP * sizeof(float));
    uint64_t *bar = &bars[slot % F32_STAGES];
    mbar_expect_tx(bar, bytes);
    tma_bulk_g2s(tiles + (size_t)(slot % F32_STAGES) * (F32_TILE_BYTES / 4), P.comps + a * SP, bytes, bar);
  };
  int p_tile = 0, p_slot = 0;  // producer (thread 0): F32_STAGES tiles ahead of the consumers
  if (tid == 0) {
    p_tile = nxt(0);
    while (p_tile < ntiles && p_slot < F32_STAGES) {
      issue(p_tile, p_slot);
      ++p_slot;
      p_tile = nxt(p_tile + 1);
    }
  }

  const int64_t qbase = (int64_t)blk * (F32_THREADS * Q);
  f32x2 x2[Q][D];  // each query coordinate broadcast into both halves
  double sum[Q];
  int64_t self[Q];
#pragma unroll
  for (int i = 0; i < Q; ++i) {
    int64_t qi = qbase + tid + (int64_t)i * F32_THREADS;
    if (qi >= P.M) qi = P.M - 1;
    const double *src = P.queries + qi * (int64_t)P.qstride;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const float v = (float)((src[k] - P.ctr[k]) * P.scl[k]);
      x2[i][k] = pack2(v, v);
    }
    sum[i] = 0.0;
    self[i] = LOO ? qi : -1;
  }
  const int64_t qlo = qbase, qhi = qbase + F32_THREADS * Q;

  int slot = 0;
  for (int t = nxt(0); t < ntiles; t = nxt(t + 1), ++slot) {
    const int64_t a = (int64_t)t * TP;  // first pair of the tile
    const int cnt = (int)((npairs - a < TP) ? (npairs - a) : TP);
    mbar_wait(&bars[slot % F32_STAGES], (uint32_t)((slot / F32_STAGES) & 1));
    const float *rec = tiles + (size_t)(slot % F32_STAGES) * (F32_TILE_BYTES / 4);
    const bool check = LOO && (2 * a < qhi) && (2 * (a + cnt) > qlo);
    f32x2 part[Q];
#pragma unroll
    for (int i = 0; i < Q; ++i) part[i] = 0ull;
    if (!check) {
#pragma unroll F32_UNROLL
      for (int c = 0; c < cnt; ++c) {
        f32x2 r2[SP / 2];
#pragma unroll
        for (int k = 0; k < SP / 2; k += 2) {
          const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(rec + c * SP + 2 * k);
          r2[k] = v.x;
          r2[k + 1] = v.y;
        }
#pragma unroll
        for (int i = 0; i < Q; ++i) {
          f32x2 df = sub2(x2[i][0], r2[0]);
          f32x2 acc = mul2(df, df);
#pragma unroll
          for (int k = 1; k < D; ++k) {
            df = sub2(x2[i][k], r2[k]);
            acc = fma2(df, df, acc);
          }
          float alo, ahi;
          unpack2(acc, alo, ahi);
          part[i] = fma2(pack2(ex2_neg(alo), ex2_neg(ahi)), r2[D], part[i]);
        }
      }
    } else {  // tiles that overlap the CTA's own rows: leave-one-out test per component
      for (int c = 0; c < cnt; ++c) {
        const float *r = rec + c * SP;
#pragma unroll
        for (int i = 0; i < Q; ++i) {
          float xs[D], dummy;
#pragma unroll
          for (int k = 0; k < D; ++k) unpack2(x2[i][k], xs[k], dummy);
          float plo, phi;
          unpack2(part[i], plo, phi);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < D; ++k) {
              const float df = xs[k] - r[2 * k + h];
              acc = __fmaf_rn(df, df, acc);
            }
            const float w = (2 * (a + c) + h == self[i]) ? 0.f : r[2 * D + h];
            plo = __fmaf_rn(ex2_neg(acc), w, plo);
          }
          part[i] = pack2(plo, phi);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < Q; ++i) {
      float plo, phi;
      unpack2(part[i], plo, phi);
      sum[i] += (double)plo + (double)phi;
    }
    __syncthreads();
    if (tid == 0 && p_tile < ntiles) {
      issue(p_tile, p_slot);
      ++p_slot;
      p_tile = nxt(p_tile + 1);
    }
  }
#pragma unroll
  for (int i = 0; i < Q; ++i) {
    const int64_t qi = qbase + tid + (int64_t)i * F32_THREADS;
    if (qi >= P.M) continue;
    if (PRUNED && !(sum[i] >= P.thresh)) {  // too small for the pruning bound: exact FP64 pass (eval_pruned.cu)
      const unsigned k = atomicAdd(P.nredo, 1u);
      P.redo[k] = qi;
      continue;
    }
    double v = sum[i] / P.norm;
    if (LOO) v = v / (1.0 - P.leafw[qi * P.SE + D]);
    int64_t o = (LOO && P.perm) ? P.perm[qi] : qi;
    if (PRUNED && !LOO) o = P.qidx[qi];
    P.out[o] = v;
  }
}

template <int D>
static cudaError_t launch_f32(const EvalF32Params &P, bool loo, unsigned grid, size_t smem, cudaStream_t st) {
  auto go = [&](auto k) {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<grid, F32_THREADS, smem, st>>>(P);
  };
  if (P.mask) go(eval_f32_kernel<D, false, true>);  // pruned route: free queries only
  else if (loo) go(eval_f32_kernel<D, true, false>);
  else go(eval_f32_kernel<D, false, false>);
  return cudaGetLastError();
}

// pruned route: the caller (eval_pruned.cu) supplies the mask / order / query map / exact-pass list
struct F32Prune {
  const uint32_t *mask = nullptr, *order = nullptr, *qidx = nullptr;
  int words = 0;
  double thresh = 0.0;
  int64_t *redo = nullptr;
  unsigned int *nredo = nullptr;
};
int f32_tile_pairs(int d) {
  const int SP = (2 * (d + 1) + 3) & ~3;
  int TP = 1;
  while (TP * 2 * SP * 4 <= F32_TILE_BYTES) TP *= 2;
  return TP;
}
int f32_queries_per_block() { return F32_THREADS * F32_Q; }

int eval_device_f32_ex(kdeb200_tree_t bd, const double *d_pos, int64_t M, int loo, double *d_out, cudaStream_t st,
                       int *launches, const F32Prune *pr);
int eval_device_f32(kdeb200_tree_t bd, const double *d_pos, int64_t M, int loo, double *d_out, cudaStream_t st,
                    int *launches) {
  return eval_device_f32_ex(bd, d_pos, M, loo, d_out, st, launches, nullptr);
}

int eval_device_f32_ex(kdeb200_tree_t bd, const double *d_pos, int64_t M, int loo, double *d_out, cudaStream_t st,
                       int *launches, const F32Prune *pr) {
  if (M <= 0) return 0;
  const int d = bd->d;
  const int SP = (2 * (d + 1) + 3) & ~3;
  const int64_t npairs = (bd->N + 1) / 2;
  EvalF32Params P;
  double norm = std::pow(2.0 * M_PI, (double)d / 2.0);
  double cs[16];
  for (int k = 0; k < d; ++k) {
    const double v = bd->hvar[k];
    if (!(v > 0.0) || !std::isfinite(v)) KDE_FAIL(5, "eval: bandwidth variance must be finite and > 0");
    P.ctr[k] = cs[k] = bd->root_mean[k];
    P.scl[k] = cs[8 + k] = std::sqrt(0.5 * 1.4426950408889634 / v);  // exp(-q/2) = 2^(-q/2 * log2 e)
    norm *= std::sqrt(v);
    // FP32 coordinates are centred on the density mean and scaled to bandwidth units: their rounding error grows
    // with extent / sigma and breaks the 1e-5 contract beyond ~2000 (1.0e-5 at 2000, 2.8e-5 at 5000).  No silent
    // loss of accuracy and no hidden FP64 fallback: refuse, the caller picks KDEB200_F64.
    if (!(bd->extent[k] * P.scl[k] <= 1500.0))
      KDE_FAIL(5, "eval (FP32): data extent / bandwidth = %.3g in dimension %d exceeds the range in which FP32 "
                  "coordinates keep 1e-5 relative accuracy; use KDEB200_F64", bd->extent[k] * P.scl[k], k + 1);
  }
  if (!bd->d_leaf32) {  // one-time FP32 shadow of the leaf records (freed with the tree)
    double *d_cs = nullptr;
    KDE_CUDA(cudaMallocAsync(&bd->d_leaf32, sizeof(float) * (size_t)npairs * SP, st));
    KDE_CUDA(cudaMallocAsync(&d_cs, sizeof(cs), st));
    KDE_CUDA(cudaMemcpyAsync(d_cs, cs, sizeof(cs), cudaMemcpyHostToDevice, st));
    KDE_CUDA(cudaStreamSynchronize(st));
    prep_f32_kernel<<<(unsigned)((npairs + 255) / 256), 256, 0, st>>>(bd->d_leaf, bd->SE, d, SP, bd->N, d_cs,
                                                                      bd->d_leaf32);
    KDE_CUDA(cudaGetLastError());
    KDE_CUDA(cudaFreeAsync(d_cs, st));
    if (launches) *launches += 1;
  }
  P.comps = bd->d_leaf32;
  P.queries = loo ? bd->d_leaf : d_pos;
  P.qstride = loo ? bd->SE : d;
  P.leafw = bd->d_leaf;
  P.SE = bd->SE;
  P.perm = loo ? bd->d_perm : nullptr;
  P.out = d_out;
  P.N = bd->N;
  P.M = M;
  P.loo = loo;
  P.norm = norm;
  P.mask = pr ? pr->mask : nullptr;
  P.order = pr ? pr->order : nullptr;
  P.qidx = pr ? pr->qidx : nullptr;
  P.words = pr ? pr->words : 0;
  P.thresh = pr ? pr->thresh : 0.0;
  P.redo = pr ? pr->redo : nullptr;
  P.nredo = pr ? pr->nredo : nullptr;
  int TP = 1;
  while (TP * 2 * SP * 4 <= F32_TILE_BYTES) TP *= 2;
  P.tile_pairs = TP;
  const size_t smem = F32_STAGES * F32_TILE_BYTES;
  const unsigned grid = (unsigned)((M + F32_THREADS * F32_Q - 1) / (F32_THREADS * F32_Q));
  cudaError_t e = cudaErrorInvalidValue;
  switch (d) {
    case 1: e = launch_f32<1>(P, loo, grid, smem, st); break;
    case 2: e = launch_f32<2>(P, loo, grid, smem, st); break;
    case 3: e = launch_f32<3>(P, loo, grid, smem, st); break;
    case 4: e = launch_f32<4>(P, loo, grid, smem, st); break;
    case 5: e = launch_f32<5>(P, loo, grid, smem, st); break;
    case 6: e = launch_f32<6>(P, loo, grid, smem, st); break;
    case 7: e = launch_f32<7>(P, loo, grid, smem, st); break;
    case 8: e = launch_f32<8>(P, loo, grid, smem, st); break;
  }
  if (e != cudaSuccess) KDE_FAIL(100 + (int)e, "eval_f32 kernel launch: %s", cudaGetErrorString(e));
  if (launches) *launches += 1;
  return 0;
}

}  // namespace kdeb200
