// gibbs_kernel.cuh -- K1: the multiscale Gibbs KDE-product sampler as one kernel (device code and
// launch template; gibbs.cu holds the host side, gibbs_d<N>.cu one explicit instantiation per d).
//
// Computes what gibbs1 (src/MSGibbs01.jl:527-629) computes for every output sample:
// levelInit!/initIndices! :467-497, samplePoint! :440-463 (gaussianProductMeanCov! :176-216),
// levelDown! :500-523, sampleIndices! :364-385, sampleIndex :404-429 (makeFasterSampleIndex!
// :250-328, selectLabelOnLevel :330-351, updateGlbParticlesVariance! :89-115), labels :612-616.
//
// Mapping (DESIGN.md "K1"): ONE THREAD PER CHAIN.  All chains execute the same static schedule
// of label draws (level, pass, density), so a CTA streams each level's node records once through
// a 3-stage shared-memory ring filled by 1-D TMA bulk copies and every lane reads the same record
// (broadcast LDS).  A draw is two passes: pass 1 accumulates the unnormalised weights p[z]
// SEQUENTIALLY in the reference's node order (identical summation order => identical pT and CDF
// up to exp rounding) and checkpoints the running sum every G nodes (<= 64 checkpoints, local
// memory); pass 2 re-evaluates only the chunk that contains u * pT, from global memory.
// The arithmetic of one kernel evaluation is restructured:
//   leaf levels (uniform bandwidth): sqrt(0.5/c_k) and the normaliser hoisted out of the node loop,
//     ln w folded into the exponent                       -> 3d + 8 FP64-pipe instr / node
//   internal levels, sampleIndices!: -0.5/b_k and ln w - 0.5 sum ln b_k precomputed per node
//   internal levels, sampleIndex   : c_k = b_k + Calmost_k, sum_k ln c_k -> one rsqrt(prod c_k), which also
//     yields the d reciprocals at d = 3                   -> 34 FP64-pipe instr / node
#pragma once
#include <cmath>
#include <cstring>
#include <vector>

#include "tree.cuh"

namespace kdeb200 {

#ifndef GB_THREADS_N
#define GB_THREADS_N 128
#endif
#ifndef GB_STAGES_N
#define GB_STAGES_N 3
#endif
#ifndef GB_TILE_N
#define GB_TILE_N 8192
#endif
constexpr int GB_THREADS = GB_THREADS_N;  // chains per CTA
constexpr int GB_STAGES = GB_STAGES_N;    // ring depth
constexpr int GB_TILE_BYTES = GB_TILE_N;  // bytes per ring stage
#ifndef GB_MAXCK_N
#define GB_MAXCK_N 64
#endif
constexpr int GB_MAXCK = GB_MAXCK_N;  // checkpoints per draw (per-thread local memory: 8 bytes each)
#ifndef GB_UNROLL_A
#define GB_UNROLL_A 4
#endif
#ifndef GB_UNROLL_C
#define GB_UNROLL_C 2
#endif
#ifndef GB_MINBLOCKS
#define GB_MINBLOCKS 4
#endif

enum : int { VAR_A = 0, VAR_B = 1, VAR_C = 2 };

// nodes per software-pipelined group / double group of pass 1 (host and device must agree: the host
// never picks a checkpoint chunk smaller than one double group once a level has two of them)
#ifndef GB_UNROLL_D6
#define GB_UNROLL_D6 1
#endif
__host__ __device__ constexpr int gb_unr(int D, int var) { return (D <= 4) ? ((var == VAR_C) ? GB_UNROLL_C : GB_UNROLL_A) : (D <= 5 ? 2 : (D == 6 ? GB_UNROLL_D6 : 1)); }
__host__ __device__ constexpr int gb_dg(int D, int var) { return 2 * gb_unr(D, var); }

struct alignas(16) Draw {
  const double *rec;        // evaluation records of this level (variant-specific layout)
  const double *rec_state;  // records the chain state is refreshed from ([m.., lnw] or [m.., b.., lnw])
  const double *wts;        // raw weights (fallback path)
  const int64_t *levperm;   // permutation of the level's nodes (label recording)
  int n;                    // nodes on the level
  int stride;               // doubles per evaluation record
  int state_stride;
  int G;                    // checkpoint chunk (power of two)
  int nchunks;
  int tnodes;               // nodes per tile (power of two)
  int ntiles;
  int tile0;                // first tile of this draw in the per-sample tile stream
  short j;                  // density
  signed char variant;      // VAR_A / VAR_B / VAR_C
  signed char kind;         // 0: sampleIndices! (against X), 1: sampleIndex (leave-one-out product)
  signed char state_has_bw; // rec_state carries per-node variances
  signed char new_level;    // first draw of a level: samplePoint! comes first
  short level;              // 1-based level
};

struct alignas(16) TileDesc {
  const double *src;
  uint32_t bytes;
  uint32_t pad;
};

struct GibbsParams {
  const Draw *draws;
  const TileDesc *tiles;
  int *counter;  // batch counter of this launch (zeroed before the launch)
  const double *exptab;
  ExpConsts ec;
  const double *randU, *randN;  // injected streams or null (Philox)
  double *points;               // d x (s1-s0)
  int64_t *indices;             // M x (s1-s0)
  int64_t *level_labels;        // optional [(s1-s0)][M][L]: labelsChoosen (src/MSGibbs01.jl:109-112)
  int64_t s0, s1, perU, perN;
  uint64_t seed;
  int ndraws, ntiles, M, L, T, add_entropy, nbatches;
  int literal;  // K1w only: the reference's arithmetic verbatim (divide, log, NaN rules) for degenerate bandwidths
  const double *root_rec[KDEB200_MAX_DENS];  // [m.., b.., lnw] of node 1
  const int64_t *labels[KDEB200_MAX_DENS];   // deepest level: permutation + 1
  double hvar[KDEB200_MAX_DENS][KDEB200_MAX_DIM];
  unsigned char mask[KDEB200_MAX_DENS][KDEB200_MAX_DIM];   // partialDimMask[j][k]
  unsigned char other[KDEB200_MAX_DENS][KDEB200_MAX_DIM];  // OR_{i != j} mask[i][k]
  // K1f (gibbs_f32.cu) only: the affine map of its FP32 records, x' = (x - nctr) * nisig, and the slow-draw counter
  double nctr[KDEB200_MAX_DIM], nisig[KDEB200_MAX_DIM];
  unsigned long long *slow_draws;
};

// ---- one kernel evaluation, three record layouts -----------------------------------------
// Explicit rn intrinsics: pass 1 and pass 2 must produce bit-identical p[z].
template <int D, bool MASK>
struct Hoist {
  double mu[D];    // X or Malmost
  double ich[D];   // variant A: sqrt(0.5 / (h_k + Calmost_k))
  double cadd[D];  // variant C: Calmost_k
  bool act[D];     // dimension participates (partialDimMask logic, :270-285)
};

// record -> registers with 16-byte loads (records are 16-byte aligned, strides are even)
template <int S>
__device__ __forceinline__ void load_rec(const double *__restrict__ r, double (&rr)[S]) {
#pragma unroll
  for (int k = 0; k < S; k += 2) {
    const double2 v = *reinterpret_cast<const double2 *>(r + k);
    rr[k] = v.x;
    rr[k + 1] = v.y;
  }
}

// the same through the read-only path (LDG.NC, L1-cached): GLOBAL pointers only -- the warp-per-chain kernel, whose
// four chains per CTA and T + 1 passes per level re-read the same records
template <int S>
__device__ __forceinline__ void load_rec_nc(const double *__restrict__ r, double (&rr)[S]) {
#pragma unroll
  for (int k = 0; k < S; k += 2) {
    const double2 v = __ldg(reinterpret_cast<const double2 *>(r + k));
    rr[k] = v.x;
    rr[k + 1] = v.y;
  }
}

// One evaluation is split in two stages so that pass 1 can software-pipeline them (stage 1 of
// node group g overlaps the exp chains of group g-1 in one basic block):
//   pre_*: record -> exponent argument (and, variant C, the normaliser rsqrt(prod c_k))
//   fin  : p = exp(arg) [* scale]
template <int D, bool MASK, bool NC = false>
__device__ __forceinline__ void pre_A(const double *__restrict__ r, const Hoist<D, MASK> &h, double &arg, double &sc) {
  constexpr int S = (D + 2) & ~1;
  double rr[S];
  if (NC) load_rec_nc<S>(r, rr); else load_rec<S>(r, rr);
  double acc = rr[D];
#pragma unroll
  for (int k = 0; k < D; ++k) {
    if (MASK && !h.act[k]) continue;
    // e = (m - mu) * sqrt(0.5/c); acc -= e*e : the FMA reads e twice => two distinct register sources
    const double e = __dmul_rn(__dadd_rn(rr[k], -h.mu[k]), h.ich[k]);
    acc = __fma_rn(-e, e, acc);
  }
  arg = acc;
  sc = 1.0;
}

template <int D, bool MASK, bool NC = false>
__device__ __forceinline__ void pre_B(const double *__restrict__ r, const Hoist<D, MASK> &h, double &arg, double &sc) {
  constexpr int S = (2 * D + 2) & ~1;
  double rr[S];
  if (NC) load_rec_nc<S>(r, rr); else load_rec<S>(r, rr);
  double acc = rr[2 * D];
#pragma unroll
  for (int k = 0; k < D; ++k) {
    const double df = __dadd_rn(rr[k], -h.mu[k]);
    acc = __fma_rn(__dmul_rn(df, df), rr[D + k], acc);
  }
  arg = acc;
  sc = 1.0;
}

template <int D, bool MASK, bool NC = false>
__device__ __forceinline__ void pre_C(const double *__restrict__ r, const Hoist<D, MASK> &h, double &arg, double &sc) {
  constexpr int S = (2 * D + 2) & ~1;
  double rr[S];
  if (NC) load_rec_nc<S>(r, rr); else load_rec<S>(r, rr);
  if (D == 3 && !MASK) {
    // one MUFU.RSQ64H for the normaliser AND the three reciprocals:
    //   rs = rsqrt(c0 c1 c2), R = rs^2 = 1/(c0 c1 c2), 1/c_k = R * prod_{i != k} c_i   (32 FP64 instr / node)
    const double c0 = __dadd_rn(rr[3], h.cadd[0]), c1 = __dadd_rn(rr[4], h.cadd[1]), c2 = __dadd_rn(rr[5], h.cadd[2]);
    const double d0 = __dadd_rn(rr[0], -h.mu[0]), d1 = __dadd_rn(rr[1], -h.mu[1]), d2 = __dadd_rn(rr[2], -h.mu[2]);
    const double c01 = __dmul_rn(c0, c1);
    const double rs = kde_rsqrt(__dmul_rn(c01, c2));
    const double Rv = __dmul_rn(rs, rs);
    // sum_k d_k^2 / c_k = R (d2^2 c0 c1 + c2 (d0^2 c1 + d1^2 c0)): 8 operations (the per-dimension form R c_i c_j d_k^2 took 10)
    double a = __dmul_rn(__dmul_rn(d0, d0), c1);
    a = __fma_rn(__dmul_rn(d1, d1), c0, a);
    a = __dmul_rn(a, c2);
    a = __fma_rn(__dmul_rn(d2, d2), c01, a);
    const double quad = __dmul_rn(a, Rv);
    arg = __fma_rn(quad, -0.5, rr[6]);
    sc = rs;
    return;
  }
  if (D == 1 && !MASK) {  // d = 1: 1/c = rsqrt(c)^2
    const double c0 = __dadd_rn(rr[1], h.cadd[0]);
    const double d0 = __dadd_rn(rr[0], -h.mu[0]);
    const double rs = kde_rsqrt(c0);
    const double quad = __dmul_rn(__dmul_rn(d0, d0), __dmul_rn(rs, rs));
    arg = __fma_rn(quad, -0.5, rr[2]);
    sc = rs;
    return;
  }
  if (D == 2 && !MASK) {  // the same idea at d = 2: 1/c0 = c1 R, 1/c1 = c0 R
    const double c0 = __dadd_rn(rr[2], h.cadd[0]), c1 = __dadd_rn(rr[3], h.cadd[1]);
    const double d0 = __dadd_rn(rr[0], -h.mu[0]), d1 = __dadd_rn(rr[1], -h.mu[1]);
    const double rs = kde_rsqrt(__dmul_rn(c0, c1));
    const double Rv = __dmul_rn(rs, rs);
    double a = __dmul_rn(__dmul_rn(d0, d0), c1);
    a = __fma_rn(__dmul_rn(d1, d1), c0, a);
    const double quad = __dmul_rn(a, Rv);
    arg = __fma_rn(quad, -0.5, rr[4]);
    sc = rs;
    return;
  }
  if (D == 4 && !MASK) {  // ... and at d = 4 with two pair products: 1/c0 = c1 (c2 c3 R), 1/c2 = c3 (c0 c1 R), ...
    const double c0 = __dadd_rn(rr[4], h.cadd[0]), c1 = __dadd_rn(rr[5], h.cadd[1]);
    const double c2 = __dadd_rn(rr[6], h.cadd[2]), c3 = __dadd_rn(rr[7], h.cadd[3]);
    const double d0 = __dadd_rn(rr[0], -h.mu[0]), d1 = __dadd_rn(rr[1], -h.mu[1]);
    const double d2 = __dadd_rn(rr[2], -h.mu[2]), d3 = __dadd_rn(rr[3], -h.mu[3]);
    const double c01 = __dmul_rn(c0, c1), c23 = __dmul_rn(c2, c3);
    const double rs = kde_rsqrt(__dmul_rn(c01, c23));
    const double Rv = __dmul_rn(rs, rs);
    double a = __dmul_rn(__dmul_rn(d0, d0), c1);
    a = __fma_rn(__dmul_rn(d1, d1), c0, a);
    double b = __dmul_rn(__dmul_rn(d2, d2), c3);
    b = __fma_rn(__dmul_rn(d3, d3), c2, b);
    const double quad = __dmul_rn(__fma_rn(b, c01, __dmul_rn(a, c23)), Rv);
    arg = __fma_rn(quad, -0.5, rr[8]);
    sc = rs;
    return;
  }
  double quad = 0.0;
  double prod = 1.0;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    if (MASK && !h.act[k]) continue;
    const double c = __dadd_rn(rr[D + k], h.cadd[k]);
    const double df = __dadd_rn(rr[k], -h.mu[k]);
    quad = __fma_rn(__dmul_rn(df, df), kde_rcp(c), quad);
    prod = __dmul_rn(prod, c);
  }
  arg = __fma_rn(quad, -0.5, rr[2 * D]);
  sc = kde_rsqrt(prod);
}

template <int D, bool MASK, int VAR, bool NC = false>
__device__ __forceinline__ void pre_node(const double *__restrict__ r, const Hoist<D, MASK> &h, double &arg,
                                         double &sc) {
  if (VAR == VAR_A) pre_A<D, MASK, NC>(r, h, arg, sc);
  else if (VAR == VAR_B) pre_B<D, MASK, NC>(r, h, arg, sc);
  else pre_C<D, MASK, NC>(r, h, arg, sc);
}

template <int VAR>
__device__ __forceinline__ double fin_node(double arg, double sc, const double *__restrict__ tab, const ExpConsts &ec) {
  const double e = kde_exp_flush(arg, tab, ec);
  return (VAR == VAR_C) ? __dmul_rn(e, sc) : e;
}

template <int D, bool MASK, int VAR, bool NC = false>
__device__ __forceinline__ double eval_node(const double *__restrict__ r, const Hoist<D, MASK> &h,
                                            const double *__restrict__ tab, const ExpConsts &ec) {
  double arg, sc;
  pre_node<D, MASK, VAR, NC>(r, h, arg, sc);
  return fin_node<VAR>(arg, sc, tab, ec);
}

// pass 1 over the tiles of one draw: sequential sum, checkpoints every G nodes
struct Ring {
  double *tiles;
  uint64_t *bars;
  const TileDesc *descs;
  int64_t known;   // tiles of the batches this CTA has claimed so far (current + one look-ahead): the producer
                   // may prefetch up to here; grows by ntiles whenever another batch is claimed
  int64_t issued;  // tiles handed to the TMA unit so far (meaningful in thread 0 only)
  int ntiles;      // tiles per sample schedule
  TileDesc nxt;    // descriptor of tile `issued`, prefetched (thread 0)
  int nxt_idx;     // issued % ntiles
};

__device__ __forceinline__ void ring_issue(const Ring &R, int64_t q, const TileDesc &td) {
  uint64_t *bar = &R.bars[q % GB_STAGES];
  mbar_expect_tx(bar, td.bytes);
  tma_bulk_g2s(R.tiles + (size_t)(q % GB_STAGES) * (GB_TILE_BYTES / 8), td.src, td.bytes, bar);
}

// thread 0: keep the ring GB_STAGES tiles ahead of the `consumed` tiles, never past the claimed batches.
// Stage issued % GB_STAGES is free once tile issued - GB_STAGES has been consumed.  The descriptor of the NEXT tile is
// loaded right after an issue and stays in flight in registers until the next call: the producer's global load used to
// sit on warp 0's critical path once per tile (ncu: a quarter of warp 0's stall samples in the FP32 kernel).
__device__ __forceinline__ void ring_fill(Ring &R, int64_t consumed) {
  while (R.issued < R.known && R.issued < consumed + GB_STAGES) {
    ring_issue(R, R.issued, R.nxt);
    ++R.issued;
    if (++R.nxt_idx == R.ntiles) R.nxt_idx = 0;  // == issued % ntiles without the 64-bit division
    R.nxt = R.descs[R.nxt_idx];
  }
}

// consume one pipelined group: p = exp(arg)[*sc], sequential adds
template <int VAR, int UNR>
__device__ __forceinline__ void consume_group(const double (&arg)[UNR], const double (&sc)[UNR],
                                              const double *__restrict__ tab, const ExpConsts &ec, double &S) {
  double p[UNR];
#pragma unroll
  for (int u = 0; u < UNR; ++u) p[u] = fin_node<VAR>(arg[u], sc[u], tab, ec);
#pragma unroll
  for (int u = 0; u < UNR; ++u) S = __dadd_rn(S, p[u]);
}

template <int D, bool MASK, int VAR>
__device__ __forceinline__ double pass1(const Draw &dr, const Hoist<D, MASK> &h, const double *__restrict__ tab,
                                        const ExpConsts &ec, Ring &R, int64_t &q, double *__restrict__ ck) {
  // UNR nodes per group.  Within a tile the groups are software-pipelined with two register sets
  // (A/B ping-pong, no rotation moves): stage 1 (record -> exponent) of group g is issued in the
  // same basic block as the exp chains of group g-1.  Fewer nodes per group at high d, where ptxas
  // 12.9 segfaults on wide unrolls.
  constexpr int UNR = gb_unr(D, VAR);
  constexpr int DG = 2 * UNR;
  constexpr int stride = (VAR == VAR_A) ? ((D + 2) & ~1) : ((2 * D + 2) & ~1);  // == dr.stride
  const int n = dr.n, G = dr.G;
  double S = 0.0;
  int c = 0;
  int consumed = 0;
  for (int t = 0; t < dr.ntiles; ++t, ++q) {
    const int cnt = (n - consumed < dr.tnodes) ? (n - consumed) : dr.tnodes;
    mbar_wait(&R.bars[q % GB_STAGES], (uint32_t)((q / GB_STAGES) & 1));
    const double *rec = R.tiles + (size_t)(q % GB_STAGES) * (GB_TILE_BYTES / 8);
    int z = 0;
    if (D <= 6 && G >= DG && cnt >= DG) {  // checkpoints fall on double-group boundaries (the host never picks a
                                           // smaller chunk for n >= 2 DG); d >= 7: ptxas 12.9 segfaults on this body
      const int full = cnt & ~(DG - 1);
      double a0[UNR], s0[UNR], a1[UNR], s1[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) pre_node<D, MASK, VAR>(rec + (size_t)u * stride, h, a0[u], s0[u]);
      for (; z + DG < full; z += DG) {
#pragma unroll
        for (int u = 0; u < UNR; ++u) pre_node<D, MASK, VAR>(rec + (size_t)(z + UNR + u) * stride, h, a1[u], s1[u]);
        consume_group<VAR, UNR>(a0, s0, tab, ec, S);
#pragma unroll
        for (int u = 0; u < UNR; ++u) pre_node<D, MASK, VAR>(rec + (size_t)(z + DG + u) * stride, h, a0[u], s0[u]);
        consume_group<VAR, UNR>(a1, s1, tab, ec, S);
        consumed += DG;
        if ((consumed & (G - 1)) == 0) ck[c++] = S;
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) pre_node<D, MASK, VAR>(rec + (size_t)(z + UNR + u) * stride, h, a1[u], s1[u]);
      consume_group<VAR, UNR>(a0, s0, tab, ec, S);
      consume_group<VAR, UNR>(a1, s1, tab, ec, S);
      consumed += DG;
      if ((consumed & (G - 1)) == 0 || consumed == n) ck[c++] = S;
      z = full;
    }
    for (; z < cnt; ++z) {  // small levels and the ragged end of the last tile: one node at a time
      S = __dadd_rn(S, eval_node<D, MASK, VAR>(rec + (size_t)z * stride, h, tab, ec));
      consumed += 1;
      if ((consumed & (G - 1)) == 0 || consumed == n) ck[c++] = S;
    }
    __syncthreads();  // stage free again
    if (threadIdx.x == 0) ring_fill(R, q + 1);
  }
  return S;
}

// pass 2: locate the first z with target <= prefix(z) inside chunk cs (per-lane global loads)
template <int D, bool MASK, int VAR>
__device__ __forceinline__ int pass2(const Draw &dr, const Hoist<D, MASK> &h, const double *__restrict__ tab,
                                     const ExpConsts &ec, int cs, double S, double target) {
  constexpr int stride = (VAR == VAR_A) ? ((D + 2) & ~1) : ((2 * D + 2) & ~1);  // == dr.stride
  const int z0 = cs * dr.G;
  const int z1 = (z0 + dr.G < dr.n) ? z0 + dr.G : dr.n;
  const double *r = dr.rec + (size_t)z0 * stride;
  int zs = z1 - 1;
  bool found = false;
  for (int z = z0; z < z1; ++z) {
    S = __dadd_rn(S, eval_node<D, MASK, VAR>(r, h, tab, ec));
    if (!found && target <= S) {
      zs = z;
      found = true;
    }
    r += stride;
  }
  return zs;
}

// MD = capacity of the per-chain state in densities (4, 8 or 16 >= M): the state lives in per-thread local memory, and
// sizing it by the call instead of by KDEB200_MAX_DENS keeps the chip-wide local-memory footprint (threads x bytes)
// inside the L2 next to the trees (ncu: DRAM traffic of the C4 launch, profiles/ncu_issued.json).
template <int D, bool MASK, int MD>
__global__ void __launch_bounds__(GB_THREADS, GB_MINBLOCKS) gibbs_kernel(const __grid_constant__ GibbsParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *tiles = reinterpret_cast<double *>(smem_raw);
  __shared__ __align__(16) double tab[KDE_EXP_TAB];
  __shared__ __align__(8) uint64_t bars[GB_STAGES];
  const int tid = threadIdx.x;
  const int M = P.M;

  for (int i = tid; i < KDE_EXP_TAB; i += GB_THREADS) tab[i] = P.exptab[i];
  if (tid == 0) {
    for (int s = 0; s < GB_STAGES; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  // Dynamic batch scheduling: CTAs claim batches of 128 chains from a global counter.  The next batch is
  // claimed at the start of the LAST draw of the current one (late, so that no CTA hoards work, but early
  // enough for the tile ring to prefetch across the batch boundary).  Chains are addressed by sample index,
  // so the result does not depend on which CTA runs which batch.
  // Every CTA ends on exactly one failing claim, so a launch performs nbatches + gridDim.x claims in total: the thread
  // that draws the last ticket puts the counter back to zero for the next launch that is handed this slot (gibbs.cu).
  __shared__ int claim;
  const int last_ticket = P.nbatches + (int)gridDim.x - 1;
  if (tid == 0) {
    claim = atomicAdd(P.counter, 1);
    if (claim == last_ticket) *P.counter = 0;
  }
  __syncthreads();
  int batch = claim, batch_next = P.nbatches;
  Ring R;
  R.tiles = tiles;
  R.bars = bars;
  R.descs = P.tiles;
  R.ntiles = P.ntiles;
  R.known = (batch < P.nbatches) ? P.ntiles : 0;
  R.issued = 0;
  R.nxt = P.tiles[0];
  R.nxt_idx = 0;
  if (tid == 0) ring_fill(R, 0);
  int64_t q = 0;

  // chain state: lambda = 1/variance and lambda*mu of the currently selected node of each density
  double lam[MD * D];
  double lmu[MD * D];
  double ck[GB_MAXCK];
  int selpos[MD];

  while (batch < P.nbatches) {
    int64_t s = P.s0 + (int64_t)batch * GB_THREADS + tid;
    const bool live = s < P.s1;
    if (!live) s = P.s1 - 1;  // idle lanes replay the last chain (keeps the CTA in lock-step)

    // levelInit!/initIndices!/calcIndices!: every density starts at the root
    for (int j = 0; j < M; ++j) {
      const double *rr = P.root_rec[j];
#pragma unroll
      for (int k = 0; k < D; ++k) {
        if (MASK && !P.mask[j][k]) {
          lam[j * D + k] = 0.0;
          lmu[j * D + k] = 0.0;
        } else {
          const double l = 1.0 / rr[D + k];
          lam[j * D + k] = l;
          lmu[j * D + k] = rr[k] * l;
        }
      }
      selpos[j] = 0;
    }

    double X[D];
    for (int di = 0; di < P.ndraws; ++di) {
      const Draw dr = P.draws[di];
      const int j = dr.j;
      if (di == P.ndraws - 1) {  // claim the next batch (uniform branch; two barriers around the shared word)
        __syncthreads();
        if (tid == 0) {
          claim = atomicAdd(P.counter, 1);
          if (claim == last_ticket) *P.counter = 0;
        }
        __syncthreads();
        batch_next = claim;
        if (batch_next < P.nbatches) R.known += P.ntiles;
      }

      if (dr.new_level) {  // samplePoint!(addEntropy = true) with normals g(level-1, k)
#pragma unroll
        for (int k = 0; k < D; ++k) {
          double Lm = 0.0, Hm = 0.0;
          bool any = !MASK;
          for (int i = 0; i < M; ++i) {
            if (MASK && P.mask[i][k]) any = true;
            Lm += lam[i * D + k];
            Hm += lmu[i * D + k];
          }
          const uint32_t slot = (uint32_t)((dr.level - 1) * D + k);
          const double g = P.randN ? P.randN[s * P.perN + slot] : philox_normal(P.seed, (uint64_t)s, slot);
          if (any) {
            const double cov = 1.0 / Lm;
            X[k] = cov * Hm + sqrt(cov) * g;
          } else {
            X[k] = 0.0;
          }
        }
      }

      Hoist<D, MASK> h;
      double scale = 1.0;  // normaliser common to the whole level (variant A), for the 1e-99 test
      if (dr.kind == 0) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
          h.mu[k] = X[k];
          h.cadd[k] = 0.0;
          h.act[k] = MASK ? (P.mask[j][k] && P.other[j][k]) : true;
        }
      } else {  // leave-one-out product of the other densities' selected kernels
#pragma unroll
        for (int k = 0; k < D; ++k) {
          double Lm = 0.0, Hm = 0.0;
          for (int i = 0; i < M; ++i) {
            if (i == j) continue;
            Lm += lam[i * D + k];
            Hm += lmu[i * D + k];
          }
          const bool oth = MASK ? (P.other[j][k] != 0) : (M > 1);
          if (oth) {
            const double cov = 1.0 / Lm;
            h.cadd[k] = cov;
            h.mu[k] = cov * Hm;
          } else {
            h.cadd[k] = 0.0;
            h.mu[k] = 0.0;
          }
          h.act[k] = (MASK ? (P.mask[j][k] != 0) : true) && oth;
        }
      }
      if (dr.variant == VAR_A) {
        double prod = 1.0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
          const double c = P.hvar[j][k] + h.cadd[k];
          h.ich[k] = sqrt(0.5 / c);
          if (!MASK || h.act[k]) prod *= c;
        }
        scale = kde_rsqrt(prod);
      }

      double pT;
      if (dr.variant == VAR_A)
        pT = pass1<D, MASK, VAR_A>(dr, h, tab, P.ec, R, q, ck);
      else if (dr.variant == VAR_B)
        pT = pass1<D, MASK, VAR_B>(dr, h, tab, P.ec, R, q, ck);
      else
        pT = pass1<D, MASK, VAR_C>(dr, h, tab, P.ec, R, q, ck);

      // selectLabelOnLevel: the c-th call of this chain reads randU[(s*perU + c) - 1]
      int zs = 0;
      if (dr.n > 1) {
        const uint32_t c = (uint32_t)(M + di);
        const double u = P.randU ? P.randU[s * P.perU + c - 1] : philox_uniform(P.seed, (uint64_t)s, c);
        if (pT * scale < 1e-99) {  // :311-315: all p[z] = weight(last node)
          const double w = dr.wts[dr.n - 1];
          double tot = 0.0;
          for (int z = 0; z < dr.n; ++z) tot += w;
          const double qv = w / tot;
          double cdf = 0.0;
          zs = dr.n - 1;
          bool found = false;
          for (int z = 0; z < dr.n - 1; ++z) {
            cdf += qv;
            if (!found && u <= cdf) {
              zs = z;
              found = true;
            }
          }
        } else {
          const double target = u * pT;
          int lo = 0, hi = dr.nchunks;  // first chunk whose end-prefix reaches the target
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (target <= ck[mid]) hi = mid; else lo = mid + 1;
          }
          if (lo >= dr.nchunks) {
            zs = dr.n - 1;
          } else if (dr.G == 1) {
            zs = lo;
          } else {
            const double S0 = lo > 0 ? ck[lo - 1] : 0.0;
            if (dr.variant == VAR_A)
              zs = pass2<D, MASK, VAR_A>(dr, h, tab, P.ec, lo, S0, target);
            else if (dr.variant == VAR_B)
              zs = pass2<D, MASK, VAR_B>(dr, h, tab, P.ec, lo, S0, target);
            else
              zs = pass2<D, MASK, VAR_C>(dr, h, tab, P.ec, lo, S0, target);
          }
        }
      }
      selpos[j] = zs;
      if (P.level_labels && dr.kind == 1 && live)  // sampleIndex records permutation[ind[j]] per (sample, density, level)
        P.level_labels[((s - P.s0) * M + j) * P.L + (dr.level - 1)] = dr.levperm[zs];

      // updateGlbParticlesVariance!(j)
      {
        const double *rs = dr.rec_state + (size_t)zs * dr.state_stride;
#pragma unroll
        for (int k = 0; k < D; ++k) {
          if (MASK && !P.mask[j][k]) {
            lam[j * D + k] = 0.0;
            lmu[j * D + k] = 0.0;
          } else {
            const double var = dr.state_has_bw ? rs[D + k] : P.hvar[j][k];
            const double l = 1.0 / var;
            lam[j * D + k] = l;
            lmu[j * D + k] = rs[k] * l;
          }
        }
      }
    }

    // labels (:612-616) and the final samplePoint! (:625)
    if (live) {
      const int64_t o = s - P.s0;
      for (int j = 0; j < M; ++j) P.indices[o * M + j] = P.labels[j][selpos[j]];
#pragma unroll
      for (int k = 0; k < D; ++k) {
        double Lm = 0.0, Hm = 0.0;
        bool any = !MASK;
        for (int i = 0; i < M; ++i) {
          if (MASK && P.mask[i][k]) any = true;
          Lm += lam[i * D + k];
          Hm += lmu[i * D + k];
        }
        double v = 0.0;
        if (any) {
          const double cov = 1.0 / Lm;
          v = cov * Hm;
          if (P.add_entropy) {
            const uint32_t slot = (uint32_t)(P.L * D + k);
            const double g = P.randN ? P.randN[s * P.perN + slot] : philox_normal(P.seed, (uint64_t)s, slot);
            v = v + sqrt(cov) * g;
          }
        }
        P.points[o * D + k] = v;
      }
    }

    batch = batch_next;
  }
}

template <int D>
cudaError_t launch_gibbs_d(const GibbsParams &P, bool masked, int grid_cap, size_t smem, cudaStream_t st,
                                  int sm_count) {
  auto launch = [&](auto kern) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, GB_THREADS, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    int grid = per_sm * sm_count;
    if (grid > grid_cap) grid = grid_cap;
    kern<<<grid, GB_THREADS, smem, st>>>(P);
    return cudaGetLastError();
  };
  if (masked) return launch(gibbs_kernel<D, true, KDEB200_MAX_DENS>);
  if (P.M <= 4) return launch(gibbs_kernel<D, false, 4>);
  if (P.M <= 8) return launch(gibbs_kernel<D, false, 8>);
  return launch(gibbs_kernel<D, false, KDEB200_MAX_DENS>);
}



// ================================================================ K1w: one WARP per chain ======================
// The small configurations the reference is actually used for (products of a few densities of ~100 components, 100 to
// a few thousand samples: SURVEY.md C1 / C2) cannot fill the chip with one thread per chain -- 100 samples are one CTA
// on one SM, and the call is the latency of one thread walking 2 712 dependent kernel evaluations.  Here a warp owns a
// chain: the lanes evaluate the nodes of a level in parallel (coalesced record reads, p[z] to shared memory), ONE lane
// then adds them up sequentially in node order -- the same additions in the same order as the one-thread kernel and
// the reference, so pT, the CDF and hence the labels are bit-identical -- and the lanes search the prefix sums for
// u * pT in parallel.  Same schedule, same records, same eval_node arithmetic, same RNG addressing as gibbs_kernel;
// used when the call has too few chains for the thread-per-chain kernel (gibbs.cu picks).
constexpr int GW_WARPS = 4;  // chains per CTA

// makeFasterSampleIndex! verbatim (src/MSGibbs01.jl:287-303) for one node: per active dimension
//   distr = (mean - mu)^2 / c;  if !isnan(distr): acc += distr; acc += log(c)      p = exp(-0.5 acc) * w;  NaN -> 0
// Used for densities the restructured arithmetic cannot take (zero / huge / non-finite variances): IEEE divides and
// libdevice log / exp produce the Inf / NaN pattern the reference's rules are written for.
template <int D, bool MASK>
__device__ __forceinline__ double eval_node_literal(const double *__restrict__ r, bool has_bw, const double *__restrict__ hvar,
                                                    const Hoist<D, MASK> &h, double w) {
  double acc = 0.0;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    if (!h.act[k]) continue;
    const double c = __dadd_rn(has_bw ? __ldg(r + D + k) : hvar[k], h.cadd[k]);
    const double m = __dadd_rn(__ldg(r + k), -h.mu[k]);
    const double distr = __ddiv_rn(__dmul_rn(m, m), c);
    if (!isnan(distr)) {
      acc = __dadd_rn(acc, distr);
      acc = __dadd_rn(acc, log(c));
    }
  }
  double p = __dmul_rn(exp(__dmul_rn(-0.5, acc)), w);
  if (isnan(p)) p = 0.0;
  return p;
}

template <int D, bool MASK>
__global__ void __launch_bounds__(GW_WARPS * 32) gibbs_warp_kernel(const __grid_constant__ GibbsParams P, int nmax) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(16) double tab[KDE_EXP_TAB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int M = P.M;
  for (int i = tid; i < KDE_EXP_TAB; i += GW_WARPS * 32) tab[i] = P.exptab[i];
  __syncthreads();
  // per-warp shared memory: prefix sums | lam | lmu | X | selpos
  const size_t per_warp = (size_t)nmax + 2 * KDEB200_MAX_DENS * D + D + KDEB200_MAX_DENS / 2 + 2;
  double *pbuf = reinterpret_cast<double *>(smem_raw) + (size_t)warp * per_warp;
  double *lam = pbuf + nmax, *lmu = lam + KDEB200_MAX_DENS * D, *X = lmu + KDEB200_MAX_DENS * D;
  int *selpos = reinterpret_cast<int *>(X + D + 1);
  const unsigned FULL = 0xffffffffu;

  const int64_t s = P.s0 + (int64_t)blockIdx.x * GW_WARPS + warp;
  if (s >= P.s1) return;  // whole warp; no block-level barrier below

  for (int i = lane; i < M * D; i += 32) {  // levelInit!/initIndices!/calcIndices!: every density starts at the root
    const int j = i / D, k = i % D;
    const double *rr = P.root_rec[j];
    if (MASK && !P.mask[j][k]) {
      lam[i] = 0.0;
      lmu[i] = 0.0;
    } else {
      const double l = 1.0 / rr[D + k];
      lam[i] = l;
      lmu[i] = rr[k] * l;
    }
  }
  if (lane < M) selpos[lane] = 0;
  __syncwarp();

  Draw dr_next = P.draws[0];
  for (int di = 0; di < P.ndraws; ++di) {
    const Draw dr = dr_next;
    if (di + 1 < P.ndraws) dr_next = P.draws[di + 1];  // in flight while this draw computes
    const int j = dr.j;
    if (dr.new_level) {  // samplePoint!(addEntropy = true): lane k owns dimension k
      if (lane < D) {
        const int k = lane;
        double Lm = 0.0, Hm = 0.0;
        bool any = !MASK;
        for (int i = 0; i < M; ++i) {
          if (MASK && P.mask[i][k]) any = true;
          Lm += lam[i * D + k];
          Hm += lmu[i * D + k];
        }
        const uint32_t slot = (uint32_t)((dr.level - 1) * D + k);
        const double g = P.randN ? P.randN[s * P.perN + slot] : philox_normal(P.seed, (uint64_t)s, slot);
        double v = 0.0;
        if (any) {
          const double cov = 1.0 / Lm;
          v = cov * Hm + sqrt(cov) * g;
        }
        X[k] = v;
      }
      __syncwarp();
    }

    // the per-draw constants: lane k computes dimension k, then everybody gets a copy
    double mu_k = 0.0, cadd_k = 0.0, ich_k = 0.0, c_k = 1.0;
    bool act_k = true;
    if (lane < D) {
      const int k = lane;
      if (dr.kind == 0) {
        mu_k = X[k];
        cadd_k = 0.0;
        act_k = MASK ? (P.mask[j][k] && P.other[j][k]) : true;
      } else {
        double Lm = 0.0, Hm = 0.0;
        for (int i = 0; i < M; ++i) {
          if (i == j) continue;
          Lm += lam[i * D + k];
          Hm += lmu[i * D + k];
        }
        const bool oth = MASK ? (P.other[j][k] != 0) : (M > 1);
        if (oth) {
          const double cov = 1.0 / Lm;
          cadd_k = cov;
          mu_k = cov * Hm;
        }
        act_k = (MASK ? (P.mask[j][k] != 0) : true) && oth;
      }
      if (dr.variant == VAR_A) {
        c_k = P.hvar[j][k] + cadd_k;
        ich_k = sqrt(0.5 / c_k);
      }
    }
    Hoist<D, MASK> h;
    double scale = 1.0;
    {
      double prod = 1.0;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        h.mu[k] = __shfl_sync(FULL, mu_k, k);
        h.cadd[k] = __shfl_sync(FULL, cadd_k, k);
        h.ich[k] = __shfl_sync(FULL, ich_k, k);
        h.act[k] = __shfl_sync(FULL, (int)act_k, k) != 0;
        const double c = __shfl_sync(FULL, c_k, k);
        if (!MASK || h.act[k]) prod *= c;  // same order of multiplications as gibbs_kernel
      }
      if (dr.variant == VAR_A && !P.literal) scale = kde_rsqrt(prod);
    }

    // pass 1, parallel part: p[z]
    const int n = dr.n;
    for (int z = lane; z < n; z += 32) {
      const double *r = dr.rec + (size_t)z * dr.stride;
      double p;
      if (P.literal) p = eval_node_literal<D, MASK>(r, dr.variant != VAR_A, P.hvar[j], h, __ldg(dr.wts + z));
      else if (dr.variant == VAR_A) p = eval_node<D, MASK, VAR_A, true>(r, h, tab, P.ec);
      else if (dr.variant == VAR_B) p = eval_node<D, MASK, VAR_B, true>(r, h, tab, P.ec);
      else p = eval_node<D, MASK, VAR_C, true>(r, h, tab, P.ec);
      pbuf[z] = p;
    }
    __syncwarp();
    // pass 1, sequential part: running sums in node order (one lane; 4 loads in flight ahead of the adds)
    double pT = 0.0;
    if (P.literal) {
      // :305-327 verbatim: pT, the pT < 1e-99 rule, p[z] /= pT, running sum -- so that Inf / NaN totals select what the
      // reference's comparisons `randU <= p[z]` select (a NaN entry is never chosen, the walk ends on the last node)
      if (lane == 0) {
        double S = 0.0;
        for (int z = 0; z < n; ++z) S = __dadd_rn(S, pbuf[z]);
        if (S < 1e-99) {
          const double w = dr.wts[n - 1];
          S = 0.0;
          for (int z = 0; z < n; ++z) {
            pbuf[z] = w;
            S = __dadd_rn(S, w);
          }
        }
        double c = 0.0;
        for (int z = 0; z < n; ++z) {
          const double q = __ddiv_rn(pbuf[z], S);
          c = z ? __dadd_rn(q, c) : q;
          pbuf[z] = c;
        }
      }
      pT = 1.0;  // the search below compares u itself with the normalised running sums
    } else if (lane == 0) {
      double S = 0.0;
      int z = 0;
      for (; z + 4 <= n; z += 4) {
        const double a0 = pbuf[z], a1 = pbuf[z + 1], a2 = pbuf[z + 2], a3 = pbuf[z + 3];
        S = __dadd_rn(S, a0); pbuf[z] = S;
        S = __dadd_rn(S, a1); pbuf[z + 1] = S;
        S = __dadd_rn(S, a2); pbuf[z + 2] = S;
        S = __dadd_rn(S, a3); pbuf[z + 3] = S;
      }
      for (; z < n; ++z) {
        S = __dadd_rn(S, pbuf[z]);
        pbuf[z] = S;
      }
      pT = S;
    }
    pT = __shfl_sync(FULL, pT, 0);
    __syncwarp();

    // selectLabelOnLevel
    int zs = 0;
    if (n > 1) {
      const uint32_t c = (uint32_t)(M + di);
      const double u = P.randU ? P.randU[s * P.perU + c - 1] : philox_uniform(P.seed, (uint64_t)s, c);
      if (pT * scale < 1e-99) {  // :311-315: all p[z] = weight(last node); rare, one lane
        if (lane == 0) {
          const double w = dr.wts[n - 1];
          double tot = 0.0;
          for (int z = 0; z < n; ++z) tot += w;
          const double qv = w / tot;
          double cdf = 0.0;
          zs = n - 1;
          bool found = false;
          for (int z = 0; z < n - 1; ++z) {
            cdf += qv;
            if (!found && u <= cdf) {
              zs = z;
              found = true;
            }
          }
        }
        zs = __shfl_sync(FULL, zs, 0);
      } else {
        const double target = u * pT;
        unsigned best = (unsigned)(n - 1);
        for (int z = lane; z < n - 1; z += 32)
          if (target <= pbuf[z]) {
            best = (unsigned)z;
            break;
          }
        zs = (int)__reduce_min_sync(FULL, best);
      }
    }
    __syncwarp();  // pbuf is rewritten by the next draw
    if (lane == 0) {
      selpos[j] = zs;
      if (P.level_labels && dr.kind == 1) P.level_labels[((s - P.s0) * M + j) * P.L + (dr.level - 1)] = dr.levperm[zs];
    }
    if (lane < D) {  // updateGlbParticlesVariance!(j)
      const int k = lane;
      const double *rs = dr.rec_state + (size_t)zs * dr.state_stride;
      if (MASK && !P.mask[j][k]) {
        lam[j * D + k] = 0.0;
        lmu[j * D + k] = 0.0;
      } else {
        const double var = dr.state_has_bw ? __ldg(rs + D + k) : P.hvar[j][k];
        const double l = 1.0 / var;
        lam[j * D + k] = l;
        lmu[j * D + k] = __ldg(rs + k) * l;
      }
    }
    __syncwarp();
  }

  // labels (:612-616) and the final samplePoint! (:625)
  const int64_t o = s - P.s0;
  if (lane < M) P.indices[o * M + lane] = P.labels[lane][selpos[lane]];
  if (lane < D) {
    const int k = lane;
    double Lm = 0.0, Hm = 0.0;
    bool any = !MASK;
    for (int i = 0; i < M; ++i) {
      if (MASK && P.mask[i][k]) any = true;
      Lm += lam[i * D + k];
      Hm += lmu[i * D + k];
    }
    double v = 0.0;
    if (any) {
      const double cov = 1.0 / Lm;
      v = cov * Hm;
      if (P.add_entropy) {
        const uint32_t slot = (uint32_t)(P.L * D + k);
        const double g = P.randN ? P.randN[s * P.perN + slot] : philox_normal(P.seed, (uint64_t)s, slot);
        v = v + sqrt(cov) * g;
      }
    }
    P.points[o * D + k] = v;
  }
}

inline size_t gibbs_warp_smem(int d, int nmax) {
  return sizeof(double) * GW_WARPS * ((size_t)nmax + 2 * KDEB200_MAX_DENS * d + d + KDEB200_MAX_DENS / 2 + 2);
}

template <int D>
cudaError_t launch_gibbs_warp_d(const GibbsParams &P, bool masked, int nmax, cudaStream_t st) {
  const size_t smem = gibbs_warp_smem(D, nmax);
  const int64_t n = P.s1 - P.s0;
  const unsigned grid = (unsigned)((n + GW_WARPS - 1) / GW_WARPS);
  auto launch = [&](auto kern) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, GW_WARPS * 32, smem, st>>>(P, nmax);
    return cudaGetLastError();
  };
  return masked ? launch(gibbs_warp_kernel<D, true>) : launch(gibbs_warp_kernel<D, false>);
}

}  // namespace kdeb200
