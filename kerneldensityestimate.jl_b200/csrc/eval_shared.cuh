// eval_shared.cuh -- device helpers shared by the evaluation kernels (eval.cu) and the fused
// golden-section LOOCV kernel (lcv.cu).  Both must produce bit-identical likelihoods, so the pieces
// that define the arithmetic live here once.
#pragma once
#include "common.cuh"

namespace kdeb200 {

constexpr int EV_THREADS = 128;
constexpr int EV_STAGES = 3;
constexpr int EV_TILE_BYTES = 8192;

struct EvalParams {
  const double *comps;    // N records, stride SE
  const double *queries;  // query i at queries + i*qstride
  const int64_t *perm;    // LOO: leaf -> original index (output scatter); may be null (leaf order)
  double *out;            // M results (S == 1) ...
  double *partial;        // ... or S x M partial sums
  const double *exptab;
  ExpConsts ec;
  int64_t N, M, q0, chunk;
  int qstride, S, tile_nodes;
  double ich[KDEB200_MAX_DIM];  // -0.5 / variance_k
  double norm;                  // (2 pi)^(d/2) prod sqrt(variance_k)
};

template <int D>
struct Rec {
  static constexpr int SE = (D + 2) & ~1;
};

template <int S>
__device__ __forceinline__ void load_rec(const double *__restrict__ r, double (&rr)[S]) {
#pragma unroll
  for (int k = 0; k < S; k += 2) {
    const double2 v = *reinterpret_cast<const double2 *>(r + k);
    rr[k] = v.x;
    rr[k + 1] = v.y;
  }
}

// -0.5 * sum_k (x_k - mu_k)^2 / var_k   (distGauss! exponent, src/DualTree01.jl:32-44)
template <int D, int S>
__device__ __forceinline__ double quad(const double (&x)[D], const double (&r)[S], const double *__restrict__ ich) {
  double acc = 0.0;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    const double df = __dadd_rn(x[k], -r[k]);
    acc = __fma_rn(__dmul_rn(df, df), ich[k], acc);
  }
  return acc;
}

// next set bit at or after position `pos` of the block's mask row; returns ntile when there is none
__device__ __forceinline__ int next_tile(const uint32_t *__restrict__ row, int words, int ntile, int pos) {
  int w = pos >> 5;
  if (w >= words) return ntile;
  uint32_t bits = row[w] & (0xffffffffu << (pos & 31));
  while (bits == 0) {
    if (++w >= words) return ntile;
    bits = row[w];
  }
  return w * 32 + __ffs(bits) - 1;
}

// Rows whose fast-path total is below EV_TINY (density < 1e-275: the point is > 35 bandwidths away from
// every component) are recomputed here with libdevice exp, sequentially in leaf order, so that
// subnormal values and exact zeros (the likelihood's zero rule) match the reference.
constexpr double EV_TINY = 1e-275;
static __device__ __noinline__ double exact_row(const double *__restrict__ comps, int SE, int D, int64_t N,
                                         const double *__restrict__ xq, const double *__restrict__ ich, int64_t self) {
  double s = 0.0;
  for (int64_t i = 0; i < N; ++i) {
    if (i == self) continue;
    const double *r = comps + i * SE;
    double acc = 0.0;
    for (int k = 0; k < D; ++k) {
      const double df = __dadd_rn(xq[k], -r[k]);
      acc = __fma_rn(__dmul_rn(df, df), ich[k], acc);
    }
    s = __fma_rn(exp(acc), r[D], s);
  }
  return s;
}

// one row of sum_j W_j log L_j with the reference's zero rule (src/DualTree01.jl:460-468):
// L_j == 0 && W_j != 0 => flag (-Inf); L_j == 0 && W_j == 0 contributes log(1) * 0
__device__ __forceinline__ void loglik_term(double l, double w, double &s, int &f) {
  if (l == 0.0) {
    if (w != 0.0) f = 1;
  } else {
    s = __fma_rn(log(l), w, s);
  }
}

// fixed-order tree reduction over the CTA (blockDim.x a power of two <= 1024): result in sh[0], shf[0]
__device__ __forceinline__ void loglik_block_reduce(double s, int f, double *sh, int *shf) {
  sh[threadIdx.x] = s;
  shf[threadIdx.x] = f;
  __syncthreads();
  for (int off = blockDim.x / 2; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) {
      sh[threadIdx.x] += sh[threadIdx.x + off];
      shf[threadIdx.x] |= shf[threadIdx.x + off];
    }
    __syncthreads();
  }
}

}  // namespace kdeb200
