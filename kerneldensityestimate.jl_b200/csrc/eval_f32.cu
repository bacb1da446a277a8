// eval_f32.cu -- FP32 variant of K2 (tolerance 1e-5 relative, BASELINE.json north_star).
//
// Same sum as eval.cu (evalDirect, src/DualTree01.jl:130-162) with the pair arithmetic in FP32:
// coordinates are centred on the density's mean and pre-scaled by sqrt(0.5*log2(e)/variance_k) in
// FP64 once per tree, so a pair costs d FADD + d FFMA + 1 MUFU.EX2 + 1 FFMA (MUFU-bound for
// d <= 3: 16 ex2/clk/SM, the FMA pipe for d >= 4).  Per-tile FP32 partial sums are folded into an
// FP64 accumulator, so the error does not grow with the number of components.
#include <cmath>
#include <vector>

#include "tree.cuh"

namespace kdeb200 {

constexpr int F32_THREADS = 128;
constexpr int F32_Q = 4;
constexpr int F32_STAGES = 3;
constexpr int F32_TILE_BYTES = 8192;

struct EvalF32Params {
  const float *comps;      // N records [x'_0..x'_{d-1}, w], stride SF floats (16-byte multiple)
  const double *queries;   // raw FP64 coordinates, query i at queries + i*qstride
  const double *leafw;     // LOO: FP64 leaf records (for 1 - w_j), stride SE
  const int64_t *perm;
  double *out;
  int64_t N, M;
  int qstride, SE, tile_nodes, loo;
  double ctr[KDEB200_MAX_DIM], scl[KDEB200_MAX_DIM];
  double norm;
};

template <int D>
struct F32Rec {
  static constexpr int SF = (D + 1 + 3) & ~3;
};

__global__ void prep_f32_kernel(const double *leaf, int SE, int D, int SF, int64_t N, const double *ctr_scl,
                                float *out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  for (int k = 0; k < D; ++k) out[i * SF + k] = (float)((leaf[i * SE + k] - ctr_scl[k]) * ctr_scl[8 + k]);
  out[i * SF + D] = (float)leaf[i * SE + D];
  for (int k = D + 1; k < SF; ++k) out[i * SF + k] = 0.f;
}

template <int D, bool LOO>
__global__ void __launch_bounds__(F32_THREADS) eval_f32_kernel(const __grid_constant__ EvalF32Params P) {
  constexpr int SF = F32Rec<D>::SF;
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
  const int TN = P.tile_nodes;
  const int ntiles = (int)((P.N + TN - 1) / TN);
  auto issue = [&](int t) {
    const int64_t a = (int64_t)t * TN;
    const int64_t cnt = (P.N - a < TN) ? (P.N - a) : TN;
    const uint32_t bytes = (uint32_t)(cnt * SF * sizeof(float));
    uint64_t *bar = &bars[t % F32_STAGES];
    mbar_expect_tx(bar, bytes);
    tma_bulk_g2s(tiles + (size_t)(t % F32_STAGES) * (F32_TILE_BYTES / 4), P.comps + a * SF, bytes, bar);
  };
  if (tid == 0)
    for (int t = 0; t < F32_STAGES && t < ntiles; ++t) issue(t);

  const int64_t qbase = (int64_t)blockIdx.x * (F32_THREADS * Q);
  float x[Q][D];
  double sum[Q];
  int64_t self[Q];
#pragma unroll
  for (int i = 0; i < Q; ++i) {
    int64_t qi = qbase + tid + (int64_t)i * F32_THREADS;
    if (qi >= P.M) qi = P.M - 1;
    const double *src = P.queries + qi * (int64_t)P.qstride;
#pragma unroll
    for (int k = 0; k < D; ++k) x[i][k] = (float)((src[k] - P.ctr[k]) * P.scl[k]);
    sum[i] = 0.0;
    self[i] = LOO ? qi : -1;
  }
  const int64_t qlo = qbase, qhi = qbase + F32_THREADS * Q;

  for (int t = 0; t < ntiles; ++t) {
    const int64_t a = (int64_t)t * TN;
    const int cnt = (int)((P.N - a < TN) ? (P.N - a) : TN);
    mbar_wait(&bars[t % F32_STAGES], (uint32_t)((t / F32_STAGES) & 1));
    const float *rec = tiles + (size_t)(t % F32_STAGES) * (F32_TILE_BYTES / 4);
    const bool check = LOO && (a < qhi) && (a + cnt > qlo);
    float part[Q];
#pragma unroll
    for (int i = 0; i < Q; ++i) part[i] = 0.f;
#pragma unroll 4
    for (int c = 0; c < cnt; ++c) {
      float r[SF];
#pragma unroll
      for (int k = 0; k < SF; k += 4) {
        const float4 v = *reinterpret_cast<const float4 *>(rec + c * SF + k);
        r[k] = v.x; r[k + 1] = v.y; r[k + 2] = v.z; r[k + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < Q; ++i) {
        float nacc = 0.f;
#pragma unroll
        for (int k = 0; k < D; ++k) {
          const float df = x[i][k] - r[k];
          nacc = __fmaf_rn(-df, df, nacc);
        }
        float e;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(nacc));
        const float w = (check && a + c == self[i]) ? 0.f : r[D];  // leave-one-out
        part[i] = __fmaf_rn(e, w, part[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < Q; ++i) sum[i] += (double)part[i];
    __syncthreads();
    if (tid == 0 && t + F32_STAGES < ntiles) issue(t + F32_STAGES);
  }
#pragma unroll
  for (int i = 0; i < Q; ++i) {
    const int64_t qi = qbase + tid + (int64_t)i * F32_THREADS;
    if (qi >= P.M) continue;
    double v = sum[i] / P.norm;
    if (LOO) v = v / (1.0 - P.leafw[qi * P.SE + D]);
    const int64_t o = (LOO && P.perm) ? P.perm[qi] : qi;
    P.out[o] = v;
  }
}

template <int D>
static cudaError_t launch_f32(const EvalF32Params &P, bool loo, unsigned grid, size_t smem, cudaStream_t st) {
  if (loo) {
    auto k = eval_f32_kernel<D, true>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<grid, F32_THREADS, smem, st>>>(P);
  } else {
    auto k = eval_f32_kernel<D, false>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<grid, F32_THREADS, smem, st>>>(P);
  }
  return cudaGetLastError();
}

int eval_device_f32(kdeb200_tree_t bd, const double *d_pos, int64_t M, int loo, double *d_out, cudaStream_t st,
                    int *launches) {
  if (M <= 0) return 0;
  const int d = bd->d;
  const int SF = (d + 1 + 3) & ~3;
  EvalF32Params P;
  double norm = std::pow(2.0 * M_PI, (double)d / 2.0);
  double cs[16];
  for (int k = 0; k < d; ++k) {
    const double v = bd->hvar[k];
    if (!(v > 0.0) || !std::isfinite(v)) KDE_FAIL(5, "eval: bandwidth variance must be finite and > 0");
    P.ctr[k] = cs[k] = bd->root_mean[k];
    P.scl[k] = cs[8 + k] = std::sqrt(0.5 * 1.4426950408889634 / v);  // exp(-q/2) = 2^(-q/2 * log2 e)
    norm *= std::sqrt(v);
  }
  if (!bd->d_leaf32) {  // one-time FP32 shadow of the leaf records (freed with the tree)
    double *d_cs = nullptr;
    KDE_CUDA(cudaMallocAsync(&bd->d_leaf32, sizeof(float) * (size_t)bd->N * SF, st));
    KDE_CUDA(cudaMallocAsync(&d_cs, sizeof(cs), st));
    KDE_CUDA(cudaMemcpyAsync(d_cs, cs, sizeof(cs), cudaMemcpyHostToDevice, st));
    KDE_CUDA(cudaStreamSynchronize(st));
    prep_f32_kernel<<<(unsigned)((bd->N + 255) / 256), 256, 0, st>>>(bd->d_leaf, bd->SE, d, SF, bd->N, d_cs,
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
  int TN = 1;
  while (TN * 2 * SF * 4 <= F32_TILE_BYTES) TN *= 2;
  P.tile_nodes = TN;
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
