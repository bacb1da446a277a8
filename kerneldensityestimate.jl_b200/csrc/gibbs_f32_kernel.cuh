// gibbs_f32_kernel.cuh -- K1f (device code and launch template; gibbs_f32.cu holds the host side, gibbs_f32_d<N>.cu one
// explicit instantiation per d): the multiscale Gibbs sampler with the kernel evaluations in packed FP32 (MUFU-bound).
//
// Mode (b) of the north star only (free-running RNG, statistical parity): the label probabilities are evaluated in
// FP32 (2^-22 relative per term), so labels are NOT bit-exact against the reference -- with the same Philox streams a
// chain follows the FP64 kernel until the first draw whose uniform lands within ~1e-6 of a CDF edge.  Everything that
// decides WHERE a product point lies stays FP64: the chain state (lambda, lambda mu of the selected kernels, read from
// the FP64 records), samplePoint!, the final point.  Selected with kdeb200_set_gibbs_precision(KDEB200_F32).
//
// Same structure as gibbs_kernel (one thread per chain, static schedule, TMA tile ring, checkpointed two-pass draw;
// src/MSGibbs01.jl:250-351,527-629), different arithmetic:
//   * records are FP32 and hold PAIRS of nodes, [m_0(a), m_0(b), .., m_{d-1}(a), m_{d-1}(b), ..], so that Blackwell's
//     packed instructions (FFMA2 / FMUL2, PTX fma.rn.f32x2) serve two nodes each; coordinates go through one affine map
//     shared by all densities of the call (centre of the root means, pooled root spread), so FP32 sees O(1) numbers;
//   * leaf levels:  2^-(sum_k t_k^2 - log2 w), t_k = m'_k s_k - mu'_k s_k  -> 2d FFMA2 + 2 MUFU.EX2 per node pair;
//   * internal levels, sampleIndices!: records carry 0.5 log2e / b_k and log2 w - 0.5 sum log2 b_k;
//   * internal levels, sampleIndex: c_k = b_k + Calmost_k, ONE MUFU.RSQ per group of <= 4 dimensions gives the
//     normaliser and (times the other c_i) the reciprocals;
//   * a chunk's terms are summed in FP32 from zero, chunk totals are folded into an FP64 running sum (the checkpoints),
//     pass 2 repeats the chunk's additions bit for bit, so the two passes agree exactly;
//   * a draw whose FP32 total under- or overflows (< 1e-25 or > 1e30 or NaN) is redone by that lane in FP64 with the
//     reference's literal arithmetic from the FP64 records, pT < 1e-99 rule included (counted: kdeb200_gibbs_f32_slow_draws).
#pragma once
#include "gibbs_kernel.cuh"

namespace kdeb200 {



typedef unsigned long long f32x2;  // two packed floats {lo, hi}
__device__ __forceinline__ f32x2 gf_pack(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void gf_unpack(f32x2 v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 gf_fma(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 gf_mul(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float gf_ex2_neg(float a) {  // 2^(-a); the negation is a free MUFU operand modifier
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-a));
  return e;
}
__device__ __forceinline__ float gf_rsqrt(float a) {
  float e;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a));
  return e;
}

// resident CTAs per SM the register budget is cut for: 8 (64 registers) at d <= 3, where ptxas fits with <= 64 bytes of
// spill and the extra warps pay (measured: 349 / 337 / 305 ms per 1M C4 samples at 5 / 6 / 8); fewer at higher d
#ifdef GF_MINBLOCKS
__host__ __device__ constexpr int gf_minblocks(int) { return GF_MINBLOCKS; }
#else
__host__ __device__ constexpr int gf_minblocks(int D) { return D <= 3 ? 8 : (D == 4 ? 6 : 5); }
#endif
#ifndef GF_UNR_A
#define GF_UNR_A 4
#endif
#ifndef GF_UNR_C
#define GF_UNR_C 2
#endif
constexpr double GF_LOG2E = 1.4426950408889634;
constexpr double GF_HL2E = 0.7213475204444817;  // 0.5 log2 e

// floats per node PAIR (16-byte multiples)
__host__ __device__ constexpr int gf_stride(int D, int var) {
  return (var == VAR_A) ? ((2 * (D + 1) + 3) & ~3) : ((2 * (2 * D + 1) + 3) & ~3);
}
__host__ __device__ constexpr int gf_unr(int D, int var) { return (D > 4) ? 1 : ((var == VAR_A) ? GF_UNR_A : GF_UNR_C); }

template <int D>
struct Hoist32 {
  f32x2 a[D];  // A: s_k = sqrt(0.5 log2e / c'_k) (0 on inactive dimensions);  B, C: 1 / 0 activity
  f32x2 b[D];  // A: -mu'_k s_k;  B, C: -mu'_k (0 on inactive dimensions)
  f32x2 c[D];  // C: Calmost'_k (1 on inactive dimensions: no contribution to the normaliser)
};

template <int S, bool NC>
__device__ __forceinline__ void gf_load(const float *__restrict__ r, f32x2 (&v)[S / 2]) {
  if (NC && S % 8 == 0) {
    // pass 2: every lane reads ITS OWN chunk, so a warp-wide load touches 32 different lines and its cost in the LSU data
    // pipe is per instruction: one 256-bit load (LDG.E.ENL2.256, sm_100) instead of two 128-bit ones per 32 bytes
#pragma unroll
    for (int k = 0; k < S / 8; ++k) {
      asm volatile("ld.global.nc.v4.u64 {%0, %1, %2, %3}, [%4];"
                   : "=l"(v[4 * k]), "=l"(v[4 * k + 1]), "=l"(v[4 * k + 2]), "=l"(v[4 * k + 3])
                   : "l"(r + 8 * k));
    }
    return;
  }
#pragma unroll
  for (int k = 0; k < S / 4; ++k) {
    const ulonglong2 q = NC ? __ldg(reinterpret_cast<const ulonglong2 *>(r) + k) : reinterpret_cast<const ulonglong2 *>(r)[k];
    v[2 * k] = q.x;
    v[2 * k + 1] = q.y;
  }
}

// record pair -> the two exponents (negated, base 2) and, variant C, the two normalisers
template <int D, int VAR, bool NC>
__device__ __forceinline__ void gf_pre(const float *__restrict__ r, const Hoist32<D> &h, f32x2 &acc, f32x2 &sc) {
  constexpr int S = gf_stride(D, VAR);
  f32x2 v[S / 2];
  gf_load<S, NC>(r, v);
  if (VAR == VAR_A) {
    acc = v[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const f32x2 t = gf_fma(v[k], h.a[k], h.b[k]);
      acc = gf_fma(t, t, acc);
    }
    sc = 0;
  } else if (VAR == VAR_B) {
    acc = v[2 * D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const f32x2 dl = gf_fma(v[k], h.a[k], h.b[k]);
      acc = gf_fma(gf_mul(dl, dl), v[D + k], acc);
    }
    sc = 0;
  } else {
    f32x2 c[D], dl[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
      c[k] = gf_fma(v[D + k], h.a[k], h.c[k]);
      dl[k] = gf_fma(v[k], h.a[k], h.b[k]);
    }
    f32x2 quad = 0;  // +0.0f, +0.0f
    sc = 0;
#pragma unroll
    for (int g0 = 0; g0 < D; g0 += 4) {
      const int gn = (D - g0 < 4) ? D - g0 : 4;
      // rs = rsqrt(prod c), R = rs^2 = 1 / prod c, 1/c_k = R * (product of the group's other c_i), shared sub-products
      f32x2 o[4], prod, c01 = 0, c23 = 0;
      if (gn == 1) prod = c[g0];
      else if (gn == 2) prod = gf_mul(c[g0], c[g0 + 1]);
      else if (gn == 3) c01 = gf_mul(c[g0], c[g0 + 1]), prod = gf_mul(c01, c[g0 + 2]);
      else c01 = gf_mul(c[g0], c[g0 + 1]), c23 = gf_mul(c[g0 + 2], c[g0 + 3]), prod = gf_mul(c01, c23);
      float pl, ph;
      gf_unpack(prod, pl, ph);
      const f32x2 rs = gf_pack(gf_rsqrt(pl), gf_rsqrt(ph));
      const f32x2 Rv = gf_mul(rs, rs);
      sc = (g0 == 0) ? rs : gf_mul(sc, rs);
      if (gn == 1) {
        o[0] = Rv;
      } else if (gn == 2) {
        o[0] = gf_mul(Rv, c[g0 + 1]);
        o[1] = gf_mul(Rv, c[g0]);
      } else if (gn == 3) {
        const f32x2 t = gf_mul(Rv, c[g0 + 2]);
        o[0] = gf_mul(t, c[g0 + 1]);
        o[1] = gf_mul(t, c[g0]);
        o[2] = gf_mul(Rv, c01);
      } else {
        const f32x2 t01 = gf_mul(Rv, c23), t23 = gf_mul(Rv, c01);
        o[0] = gf_mul(t01, c[g0 + 1]);
        o[1] = gf_mul(t01, c[g0]);
        o[2] = gf_mul(t23, c[g0 + 3]);
        o[3] = gf_mul(t23, c[g0 + 2]);
      }
#pragma unroll
      for (int k = 0; k < gn; ++k) quad = gf_fma(gf_mul(dl[g0 + k], dl[g0 + k]), o[k], quad);
    }
    const float hl = (float)GF_HL2E;
    acc = gf_fma(quad, gf_pack(hl, hl), v[2 * D]);
  }
}

template <int VAR>
__device__ __forceinline__ void gf_fin(f32x2 acc, f32x2 sc, float &pa, float &pb) {
  float al, ah;
  gf_unpack(acc, al, ah);
  pa = gf_ex2_neg(al);
  pb = gf_ex2_neg(ah);
  if (VAR == VAR_C) {
    float sl, sh;
    gf_unpack(sc, sl, sh);
    pa = __fmul_rn(pa, sl);
    pb = __fmul_rn(pb, sh);
  }
}

// one trip: UNR record pairs -> exponents (gf_trip_pre), exponents -> 2 UNR terms added to s in node order (gf_trip_fin)
template <int D, int VAR, int UNR>
__device__ __forceinline__ void gf_trip_pre(const float *__restrict__ rec, const Hoist32<D> &h, f32x2 (&acc)[UNR], f32x2 (&sc)[UNR]) {
  constexpr int stride = gf_stride(D, VAR);
#pragma unroll
  for (int u = 0; u < UNR; ++u) gf_pre<D, VAR, false>(rec + (size_t)u * stride, h, acc[u], sc[u]);
}
template <int VAR, int UNR>
__device__ __forceinline__ void gf_trip_fin(const f32x2 (&acc)[UNR], const f32x2 (&sc)[UNR], float &s) {
  float p[2 * UNR];
#pragma unroll
  for (int u = 0; u < UNR; ++u) gf_fin<VAR>(acc[u], sc[u], p[2 * u], p[2 * u + 1]);
#pragma unroll
  for (int u = 0; u < 2 * UNR; ++u) s = __fadd_rn(s, p[u]);
}

// pass 1: FP32 sums per checkpoint chunk, FP64 running total.  Chunk-outer: the trip loop of a run (the part of a chunk
// inside one tile) carries no checkpoint test.  (An explicit two-register-set software pipeline, as in gibbs_kernel, was
// measured and dropped: at the 64-register budget that 8 CTAs per SM need it loses 3 %)
template <int D, int VAR>
__device__ __forceinline__ double gf_pass1(const Draw &dr, const Hoist32<D> &h, Ring &R, int64_t &q, double *__restrict__ ck) {
  constexpr int UNR = gf_unr(D, VAR);
  constexpr int stride = gf_stride(D, VAR);
  const int npt = (dr.n + 1) >> 1;  // pairs on the level
  const int Gp = dr.G >> 1;         // pairs per chunk
  const int tp = dr.tnodes >> 1;    // pairs per tile
  double S = 0.0;
  float s = 0.f;
  int c = 0, done = 0, left = Gp;
  for (int t = 0; t < dr.ntiles; ++t, ++q) {
    const int np = (npt - done < tp) ? (npt - done) : tp;
    mbar_wait(&R.bars[q % GB_STAGES], (uint32_t)((q / GB_STAGES) & 1));
    const float *rec = reinterpret_cast<const float *>(R.tiles + (size_t)(q % GB_STAGES) * (GB_TILE_BYTES / 8));
    int zp = 0;
    while (zp < np) {
      const int run = (left < np - zp) ? left : (np - zp);
      const float *r = rec + (size_t)zp * stride;
      int i = 0;
      for (; i + UNR <= run; i += UNR) {
        f32x2 acc[UNR], sc[UNR];
        gf_trip_pre<D, VAR, UNR>(r + (size_t)i * stride, h, acc, sc);
        gf_trip_fin<VAR, UNR>(acc, sc, s);
      }
      for (; i < run; ++i) {
        f32x2 acc[1], sc[1];
        gf_trip_pre<D, VAR, 1>(r + (size_t)i * stride, h, acc, sc);
        gf_trip_fin<VAR, 1>(acc, sc, s);
      }
      zp += run;
      done += run;
      left -= run;
      if (left == 0 || done == npt) {
        S += (double)s;
        ck[c++] = S;
        s = 0.f;
        left = Gp;
      }
    }
    __syncthreads();  // stage free again
    if (threadIdx.x == 0) ring_fill(R, q + 1);
  }
  return S;
}

// pass 2: the same additions over chunk cs, from global memory; first node whose running sum reaches t32
template <int D, int VAR>
__device__ __forceinline__ int gf_pass2(const Draw &dr, const Hoist32<D> &h, int cs, float t32) {
  constexpr int stride = gf_stride(D, VAR);
  const int z0 = cs * dr.G;
  const int z1 = (z0 + dr.G < dr.n) ? z0 + dr.G : dr.n;
  const float *r = reinterpret_cast<const float *>(dr.rec) + (size_t)(z0 >> 1) * stride;
  int zs = z1 - 1;
  bool found = false;
  float s = 0.f;
  for (int z = z0; z < z1; z += 2) {
    f32x2 acc, sc;
    float pa, pb;
    gf_pre<D, VAR, true>(r, h, acc, sc);
    gf_fin<VAR>(acc, sc, pa, pb);
    s = __fadd_rn(s, pa);
    if (!found && t32 <= s) {
      zs = z;
      found = true;
    }
    s = __fadd_rn(s, pb);
    if (!found && t32 <= s) {
      zs = z + 1;
      found = true;
    }
    r += stride;
  }
  return zs < z1 - 1 ? zs : z1 - 1;  // the zero-probability pad of an odd level is never a label
}

// A lane whose FP32 total is unusable redoes its draw alone: the reference's arithmetic verbatim in FP64 from the FP64
// records ([m.., lnw] + the uniform variances on leaf levels, [m.., b.., lnw] elsewhere), two sequential sweeps.
template <int D>
__device__ __noinline__ int gf_slow_draw(const Draw &dr, const double *__restrict__ hvar, const Hoist<D, true> &h, double u) {
  const bool has_bw = dr.state_has_bw != 0;
  const int n = dr.n;
  double pT = 0.0;
  for (int z = 0; z < n; ++z)
    pT += eval_node_literal<D, true>(dr.rec_state + (size_t)z * dr.state_stride, has_bw, hvar, h, __ldg(dr.wts + z));
  if (pT < 1e-99) {  // :311-315
    const double w = dr.wts[n - 1];
    double tot = 0.0;
    for (int z = 0; z < n; ++z) tot += w;
    const double qv = w / tot;
    double cdf = 0.0;
    for (int z = 0; z < n - 1; ++z) {
      cdf += qv;
      if (u <= cdf) return z;
    }
    return n - 1;
  }
  const double target = u * pT;
  double S = 0.0;
  for (int z = 0; z < n - 1; ++z) {
    S += eval_node_literal<D, true>(dr.rec_state + (size_t)z * dr.state_stride, has_bw, hvar, h, __ldg(dr.wts + z));
    if (target <= S) return z;
  }
  return n - 1;
}

template <int D, int MD>
__global__ void __launch_bounds__(GB_THREADS, gf_minblocks(D)) gibbs_f32_kernel(const __grid_constant__ GibbsParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *tiles = reinterpret_cast<double *>(smem_raw);
  __shared__ __align__(8) uint64_t bars[GB_STAGES];
  __shared__ int claim;
  const int tid = threadIdx.x;
  const int M = P.M;
  if (tid == 0) {
    for (int s = 0; s < GB_STAGES; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  // dynamic batch scheduling exactly as in gibbs_kernel
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

  double lam[MD * D];
  double lmu[MD * D];
  double ck[GB_MAXCK];
  int selpos[MD];
  unsigned slow = 0;

  while (batch < P.nbatches) {
    int64_t s = P.s0 + (int64_t)batch * GB_THREADS + tid;
    const bool live = s < P.s1;
    if (!live) s = P.s1 - 1;

    for (int j = 0; j < M; ++j) {
      const double *rr = P.root_rec[j];
#pragma unroll
      for (int k = 0; k < D; ++k) {
        if (!P.mask[j][k]) {
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
      if (di == P.ndraws - 1) {
        __syncthreads();
        if (tid == 0) {
          claim = atomicAdd(P.counter, 1);
          if (claim == last_ticket) *P.counter = 0;
        }
        __syncthreads();
        batch_next = claim;
        if (batch_next < P.nbatches) R.known += P.ntiles;
      }

      if (dr.new_level) {  // samplePoint!(addEntropy = true), FP64
#pragma unroll
        for (int k = 0; k < D; ++k) {
          double Lm = 0.0, Hm = 0.0;
          bool any = false;
          for (int i = 0; i < M; ++i) {
            if (P.mask[i][k]) any = true;
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

      // the conditional this draw evaluates against (FP64), then its FP32 image under the call's affine map
      Hoist<D, true> h;
      if (dr.kind == 0) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
          h.mu[k] = X[k];
          h.cadd[k] = 0.0;
          h.act[k] = P.mask[j][k] && P.other[j][k];
        }
      } else {
#pragma unroll
        for (int k = 0; k < D; ++k) {
          double Lm = 0.0, Hm = 0.0;
          for (int i = 0; i < M; ++i) {
            if (i == j) continue;
            Lm += lam[i * D + k];
            Hm += lmu[i * D + k];
          }
          const bool oth = P.other[j][k] != 0;
          if (oth) {
            const double cov = 1.0 / Lm;
            h.cadd[k] = cov;
            h.mu[k] = cov * Hm;
          } else {
            h.cadd[k] = 0.0;
            h.mu[k] = 0.0;
          }
          h.act[k] = (P.mask[j][k] != 0) && oth;
        }
      }
      Hoist32<D> g;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        const double is = P.nisig[k];
        const double mu = (h.mu[k] - P.nctr[k]) * is;
        const double ca = h.cadd[k] * is * is;
        float fa, fb, fc = 0.f;
        if (dr.variant == VAR_A) {
          const double sk = h.act[k] ? sqrt(GF_HL2E / (P.hvar[j][k] * is * is + ca)) : 0.0;
          fa = (float)sk;
          fb = (float)(-mu * sk);
        } else {
          fa = h.act[k] ? 1.f : 0.f;
          fb = h.act[k] ? (float)(-mu) : 0.f;
          fc = h.act[k] ? (float)ca : 1.f;
        }
        g.a[k] = gf_pack(fa, fa);
        g.b[k] = gf_pack(fb, fb);
        g.c[k] = gf_pack(fc, fc);
      }

      double pT;
      if (dr.variant == VAR_A)
        pT = gf_pass1<D, VAR_A>(dr, g, R, q, ck);
      else if (dr.variant == VAR_B)
        pT = gf_pass1<D, VAR_B>(dr, g, R, q, ck);
      else
        pT = gf_pass1<D, VAR_C>(dr, g, R, q, ck);

      int zs = 0;
      if (dr.n > 1) {
        const uint32_t c = (uint32_t)(M + di);
        const double u = P.randU ? P.randU[s * P.perU + c - 1] : philox_uniform(P.seed, (uint64_t)s, c);
        if (!(pT >= 1e-25 && pT <= 1e30)) {
          zs = gf_slow_draw<D>(dr, P.hvar[j], h, u);
          if (live) ++slow;
        } else {
          const double target = u * pT;
          int lo = 0, hi = dr.nchunks;
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (target <= ck[mid]) hi = mid; else lo = mid + 1;
          }
          if (lo >= dr.nchunks) {
            zs = dr.n - 1;
          } else {
            const float t32 = (float)(target - (lo > 0 ? ck[lo - 1] : 0.0));
            if (dr.variant == VAR_A)
              zs = gf_pass2<D, VAR_A>(dr, g, lo, t32);
            else if (dr.variant == VAR_B)
              zs = gf_pass2<D, VAR_B>(dr, g, lo, t32);
            else
              zs = gf_pass2<D, VAR_C>(dr, g, lo, t32);
          }
        }
      }
      selpos[j] = zs;
      if (P.level_labels && dr.kind == 1 && live)
        P.level_labels[((s - P.s0) * M + j) * P.L + (dr.level - 1)] = dr.levperm[zs];

      {  // updateGlbParticlesVariance!(j), from the FP64 records
        const double *rs = dr.rec_state + (size_t)zs * dr.state_stride;
#pragma unroll
        for (int k = 0; k < D; ++k) {
          if (!P.mask[j][k]) {
            lam[j * D + k] = 0.0;
            lmu[j * D + k] = 0.0;
          } else {
            const double var = dr.state_has_bw ? rs[D + k] : P.hvar[j][k];
            const double l = 1.0 / var, lm = rs[k] * l;
            lam[j * D + k] = l;
            lmu[j * D + k] = lm;
          }
        }
      }
    }

    if (live) {
      const int64_t o = s - P.s0;
      for (int j = 0; j < M; ++j) P.indices[o * M + j] = P.labels[j][selpos[j]];
#pragma unroll
      for (int k = 0; k < D; ++k) {
        double Lm = 0.0, Hm = 0.0;
        bool any = false;
        for (int i = 0; i < M; ++i) {
          if (P.mask[i][k]) any = true;
          Lm += lam[i * D + k];
          Hm += lmu[i * D + k];
        }
        double v = 0.0;
        if (any) {
          const double cov = 1.0 / Lm;
          v = cov * Hm;
          if (P.add_entropy) {
            const uint32_t slot = (uint32_t)(P.L * D + k);
            const double gn = P.randN ? P.randN[s * P.perN + slot] : philox_normal(P.seed, (uint64_t)s, slot);
            v = v + sqrt(cov) * gn;
          }
        }
        P.points[o * D + k] = v;
      }
    }
    batch = batch_next;
  }
  if (slow) atomicAdd(P.slow_draws, (unsigned long long)slow);
}

template <int D>
cudaError_t launch_gibbs_f32_d(const GibbsParams &P, int grid_cap, size_t smem, cudaStream_t st, int sm_count) {
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
#ifndef GF_ONLY_D3
  if (P.M <= 4) return launch(gibbs_f32_kernel<D, 4>);
#endif
  if (P.M <= 8) return launch(gibbs_f32_kernel<D, 8>);
  return launch(gibbs_f32_kernel<D, KDEB200_MAX_DENS>);
}

}  // namespace kdeb200
