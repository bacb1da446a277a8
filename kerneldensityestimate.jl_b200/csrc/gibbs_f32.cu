// gibbs_f32.cu -- K1f: the multiscale Gibbs sampler with the kernel evaluations in packed FP32 (MUFU-bound).
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
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <list>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "gibbs_kernel.cuh"

namespace kdeb200 {

extern std::atomic<uint64_t> g_tree_epoch;

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

#ifndef GF_MINBLOCKS
#define GF_MINBLOCKS 5
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
      const int g1 = (g0 + 4 < D) ? g0 + 4 : D;
      f32x2 prod = c[g0];
#pragma unroll
      for (int k = g0 + 1; k < g1; ++k) prod = gf_mul(prod, c[k]);
      float pl, ph;
      gf_unpack(prod, pl, ph);
      const f32x2 rs = gf_pack(gf_rsqrt(pl), gf_rsqrt(ph));
      const f32x2 Rv = gf_mul(rs, rs);
      sc = (g0 == 0) ? rs : gf_mul(sc, rs);
#pragma unroll
      for (int k = g0; k < g1; ++k) {
        f32x2 o = Rv;  // 1/c_k = R * prod_{i in group, i != k} c_i
#pragma unroll
        for (int i = g0; i < g1; ++i)
          if (i != k) o = gf_mul(o, c[i]);
        quad = gf_fma(gf_mul(dl[k], dl[k]), o, quad);
      }
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

// pass 1: FP32 sums per checkpoint chunk, FP64 running total
template <int D, int VAR>
__device__ __forceinline__ double gf_pass1(const Draw &dr, const Hoist32<D> &h, Ring &R, int64_t &q, double *__restrict__ ck) {
  constexpr int UNR = gf_unr(D, VAR);
  constexpr int stride = gf_stride(D, VAR);
  const int n = dr.n, G = dr.G;
  double S = 0.0;
  float s = 0.f;
  int c = 0, consumed = 0;
  for (int t = 0; t < dr.ntiles; ++t, ++q) {
    const int cnt = (n - consumed < dr.tnodes) ? (n - consumed) : dr.tnodes;
    const int np = (cnt + 1) >> 1;
    mbar_wait(&R.bars[q % GB_STAGES], (uint32_t)((q / GB_STAGES) & 1));
    const float *rec = reinterpret_cast<const float *>(R.tiles + (size_t)(q % GB_STAGES) * (GB_TILE_BYTES / 8));
    int zp = 0;
    if (G >= 2 * UNR) {  // chunk ends fall on trip ends (G and the tile size are powers of two)
      for (; zp + UNR <= np; zp += UNR) {
        f32x2 acc[UNR], sc[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) gf_pre<D, VAR, false>(rec + (size_t)(zp + u) * stride, h, acc[u], sc[u]);
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          float pa, pb;
          gf_fin<VAR>(acc[u], sc[u], pa, pb);
          s = __fadd_rn(s, pa);
          s = __fadd_rn(s, pb);
        }
        consumed += 2 * UNR;
        if ((consumed & (G - 1)) == 0 || consumed >= n) {
          S += (double)s;
          ck[c++] = S;
          s = 0.f;
        }
      }
    }
    for (; zp < np; ++zp) {
      f32x2 acc, sc;
      float pa, pb;
      gf_pre<D, VAR, false>(rec + (size_t)zp * stride, h, acc, sc);
      gf_fin<VAR>(acc, sc, pa, pb);
      s = __fadd_rn(s, pa);
      s = __fadd_rn(s, pb);
      consumed += 2;
      if ((consumed & (G - 1)) == 0 || consumed >= n) {
        S += (double)s;
        ck[c++] = S;
        s = 0.f;
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
__global__ void __launch_bounds__(GB_THREADS, GF_MINBLOCKS) gibbs_f32_kernel(const __grid_constant__ GibbsParams P) {
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
            const double l = 1.0 / var;
            lam[j * D + k] = l;
            lmu[j * D + k] = rs[k] * l;
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

// ---- FP64 level records -> FP32 pair records ----------------------------------------------------------------
struct PrepDesc {
  const double *src;  // variant A: [m.., lnw] records; B, C: [m.., b.., lnw] records
  float *dst;
  int n, variant, src_stride, dst_stride;
};
struct PrepMap {
  double ctr[KDEB200_MAX_DIM], isig[KDEB200_MAX_DIM];
};

__global__ void gibbs_f32_prep_kernel(const PrepDesc *descs, int D, PrepMap mp) {
  const PrepDesc pd = descs[blockIdx.y];
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (2 * p >= pd.n) return;
  float *o = pd.dst + (size_t)p * pd.dst_stride;
  const float inf = __int_as_float(0x7f800000);
  for (int hh = 0; hh < 2; ++hh) {
    const int z = 2 * p + hh;
    const bool ok = z < pd.n;  // an odd level is padded with a node of probability zero
    const double *r = pd.src + (size_t)(ok ? z : 0) * pd.src_stride;
    // (clamped: a dimension masked out for this density may hold anything, and inf * 0 would poison the sum)
    for (int k = 0; k < D; ++k) o[2 * k + hh] = ok ? (float)fmin(fmax((r[k] - mp.ctr[k]) * mp.isig[k], -1e30), 1e30) : 0.f;
    if (pd.variant == VAR_A) {
      o[2 * D + hh] = ok ? (float)(-r[D] * GF_LOG2E) : inf;  // -log2 w  (w = 0: +inf -> p = 0)
    } else {
      double sl = 0.0;
      for (int k = 0; k < D; ++k) {
        const double b = fmin(fmax(r[D + k] * mp.isig[k] * mp.isig[k], 1e-30), 1e30);
        sl += log2(b);
        o[2 * (D + k) + hh] = ok ? (float)(pd.variant == VAR_B ? GF_HL2E / b : b) : (pd.variant == VAR_B ? 0.f : 1.f);
      }
      const double lw2 = r[2 * D] * GF_LOG2E;
      o[4 * D + hh] = ok ? (float)(pd.variant == VAR_B ? -(lw2 - 0.5 * sl) : -lw2) : inf;
    }
  }
  const int used = (pd.variant == VAR_A) ? 2 * (D + 1) : 2 * (2 * D + 1);
  for (int k = used; k < pd.dst_stride; ++k) o[k] = 0.f;
}

// ---- host side ---------------------------------------------------------------------------------------------------
namespace {
struct Sched32 {
  uint64_t epoch = 0;
  int slot = 0, M = 0, T = 0;
  bool masked = false;
  kdeb200_tree_t trees[KDEB200_MAX_DENS] = {nullptr};
  char *d_base = nullptr;  // draws | tiles | counters | FP32 records
  Draw *d_draws = nullptr;
  TileDesc *d_tiles = nullptr;
  int *d_counters = nullptr;
  int ndraws = 0, ntiles = 0;
  unsigned next_counter = 0;
  PrepMap map;
  unsigned char mask[KDEB200_MAX_DENS][KDEB200_MAX_DIM];
};
constexpr int S32_COUNTERS = 256;
constexpr size_t S32_KEEP = 4;
std::mutex g32_mu;
std::list<Sched32> g32_sched[KDEB200_MAX_GPUS];
unsigned long long *g32_slow[KDEB200_MAX_GPUS] = {nullptr};
std::atomic<int> g_gibbs_precision{KDEB200_F64};

void s32_free(Sched32 &e, cudaStream_t st) {
  if (e.d_base) cudaFreeAsync(e.d_base, st);
  e.d_base = nullptr;
}
}  // namespace

void gibbs32_drop_schedules(int slot) {
  std::lock_guard<std::mutex> lk(g32_mu);
  for (auto &e : g32_sched[slot]) s32_free(e, ctx_at(slot).stream);
  g32_sched[slot].clear();
  if (g32_slow[slot]) cudaFreeAsync(g32_slow[slot], ctx_at(slot).stream);
  g32_slow[slot] = nullptr;
}

int gibbs_precision() {
  if (const char *ev = getenv("KDEB200_GIBBS_F32")) return atoi(ev) ? KDEB200_F32 : KDEB200_F64;
  return g_gibbs_precision.load();
}
int gibbs_set_precision(int p) {
  if (p != KDEB200_F64 && p != KDEB200_F32) KDE_FAIL(3, "set_gibbs_precision: KDEB200_F64 or KDEB200_F32");
  g_gibbs_precision.store(p);
  return 0;
}

// The affine map shared by the call's densities: centre of the root means; per dimension the largest root standard
// deviation (data spread + bandwidth) pooled with the spread of the centres.  Refuses bandwidths FP32 cannot normalise.
static int make_map(const kdeb200_tree_t *trees, int M, int d, const unsigned char (*mask)[KDEB200_MAX_DIM], PrepMap &mp) {
  for (int k = 0; k < d; ++k) {
    double c = 0.0;
    int na = 0;
    for (int j = 0; j < M; ++j)
      if (mask[j][k]) c += trees[j]->root_mean[k], ++na;
    if (na == 0) {  // nobody reads this dimension (partialDimMask): any finite map will do
      mp.ctr[k] = 0.0;
      mp.isig[k] = 1.0;
      continue;
    }
    c /= na;
    double v = 0.0;
    for (int j = 0; j < M; ++j) {
      if (!mask[j][k]) continue;
      const double dv = trees[j]->root_mean[k] - c;
      const double vj = trees[j]->root_var[k] + dv * dv;
      if (vj > v) v = vj;
    }
    if (!(v > 0.0) || !std::isfinite(v) || !std::isfinite(c)) KDE_FAIL(8, "gibbs (FP32): degenerate spread in dimension %d", k + 1);
    mp.ctr[k] = c;
    mp.isig[k] = 1.0 / std::sqrt(v);
    for (int j = 0; j < M; ++j) {
      if (!mask[j][k]) continue;
      const double hn = trees[j]->hvar[k] / v;
      if (!(hn >= 1e-8))
        KDE_FAIL(8, "gibbs (FP32): the bandwidth of density %d in dimension %d is %.3g of the data spread; the FP32 sampler "
                    "needs >= 1e-4 (products of up to four normalised variances must stay normal floats); use KDEB200_F64",
                 j + 1, k + 1, std::sqrt(hn));
    }
  }
  for (int k = d; k < KDEB200_MAX_DIM; ++k) mp.ctr[k] = 0.0, mp.isig[k] = 1.0;
  return 0;
}

static int s32_get(const kdeb200_tree_t *trees, int M, int L, int T, bool masked,
                   const unsigned char (*mask)[KDEB200_MAX_DIM], Sched32 *out, int **counter) {
  Context &c = ctx();
  std::lock_guard<std::mutex> lk(g32_mu);
  auto &lst = g32_sched[c.slot];
  const uint64_t epoch = g_tree_epoch.load();
  for (auto it = lst.begin(); it != lst.end();) {
    if (it->epoch != epoch) {
      s32_free(*it, c.stream);
      it = lst.erase(it);
    } else {
      ++it;
    }
  }
  for (auto it = lst.begin(); it != lst.end(); ++it) {
    if (it->M != M || it->T != T || it->masked != masked) continue;
    bool same = std::memcmp(it->mask, mask, sizeof(it->mask)) == 0;  // the affine map depends on the mask
    for (int j = 0; j < M; ++j) same = same && it->trees[j] == trees[j];
    if (!same) continue;
    lst.splice(lst.begin(), lst, it);
    *counter = it->d_counters + (it->next_counter++ % S32_COUNTERS);
    *out = *it;
    return 0;
  }
  const int d = trees[0]->d;
  Sched32 e;
  if (int rc = make_map(trees, M, d, mask, e.map)) return rc;
  std::memcpy(e.mask, mask, sizeof(e.mask));

  // FP32 record blocks, one per (density slot, level, variant) the schedule touches
  std::map<std::tuple<int, int, int>, size_t> block;  // -> offset in floats
  std::vector<PrepDesc> preps;
  std::vector<size_t> prep_off;
  size_t nfloats = 0;
  auto block_of = [&](int j, int li, int var) -> size_t {
    const auto key = std::make_tuple(j, li, var);
    auto it = block.find(key);
    if (it != block.end()) return it->second;
    const kdeb200_tree_s *t = trees[j];
    const Level &lv = t->levels[li];
    PrepDesc pd;
    pd.src = t->d_buf + (var == VAR_A ? lv.offA : lv.offC);
    pd.dst = nullptr;
    pd.n = (int)lv.n;
    pd.variant = var;
    pd.src_stride = (var == VAR_A) ? t->SA : t->SC;
    pd.dst_stride = gf_stride(d, var);
    const size_t off = nfloats;
    nfloats += (size_t)((lv.n + 1) / 2) * pd.dst_stride;
    preps.push_back(pd);
    prep_off.push_back(off);
    block[key] = off;
    return off;
  };
  std::vector<Draw> draws;
  struct TileRef { size_t off; uint32_t bytes; };
  std::vector<TileRef> tiles;
  std::vector<size_t> draw_off;
  for (int l = 1; l <= L; ++l)
    for (int pass = 0; pass <= T; ++pass)
      for (int j = 0; j < M; ++j) {
        const kdeb200_tree_s *t = trees[j];
        const int li = l < t->depth ? l : t->depth;
        const Level &lv = t->levels[li];
        Draw dr;
        std::memset(&dr, 0, sizeof(dr));
        dr.j = (short)j;
        dr.kind = pass == 0 ? 0 : 1;
        dr.level = (short)l;
        dr.new_level = (pass == 0 && j == 0) ? 1 : 0;
        dr.n = (int)lv.n;
        dr.wts = t->d_buf + lv.offW;
        dr.levperm = t->d_levperm + lv.offP;
        if (lv.cls == 0) {
          dr.variant = VAR_A;
          dr.rec_state = t->d_buf + lv.offA;
          dr.state_stride = t->SA;
          dr.state_has_bw = 0;
        } else {
          dr.variant = (pass == 0 && !masked) ? VAR_B : VAR_C;
          dr.rec_state = t->d_buf + lv.offC;
          dr.state_stride = t->SC;
          dr.state_has_bw = 1;
        }
        dr.stride = gf_stride(d, dr.variant);
        const size_t off = block_of(j, li, dr.variant);
        int G = 2;
        while ((int64_t)G * GB_MAXCK < lv.n) G *= 2;
        const int trip = 2 * gf_unr(d, dr.variant);
        if (G < trip && lv.n >= 2 * trip) G = trip;
        dr.G = G;
        dr.nchunks = (int)((lv.n + G - 1) / G);
        int tn = 2;  // nodes per tile: the largest power of two whose pairs fit
        while ((size_t)tn * dr.stride * sizeof(float) <= GB_TILE_BYTES) tn *= 2;
        dr.tnodes = tn;
        dr.ntiles = (int)((lv.n + tn - 1) / tn);
        dr.tile0 = (int)tiles.size();
        for (int qq = 0; qq < dr.ntiles; ++qq) {
          const int64_t a = (int64_t)qq * tn;
          const int64_t cnt = (lv.n - a < tn) ? (lv.n - a) : tn;
          tiles.push_back(TileRef{off + (size_t)(a / 2) * dr.stride, (uint32_t)(((cnt + 1) / 2) * dr.stride * sizeof(float))});
        }
        draws.push_back(dr);
        draw_off.push_back(off);
      }

  e.epoch = epoch;
  e.slot = c.slot;
  e.M = M;
  e.T = T;
  e.masked = masked;
  for (int j = 0; j < M; ++j) e.trees[j] = trees[j];
  e.ndraws = (int)draws.size();
  e.ntiles = (int)tiles.size();
  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t b_draws = up(sizeof(Draw) * draws.size()), b_tiles = up(sizeof(TileDesc) * tiles.size()),
               b_cnt = up(sizeof(int) * S32_COUNTERS), b_prep = up(sizeof(PrepDesc) * preps.size()),
               b_rec = up(sizeof(float) * nfloats);
  KDE_CUDA(cudaMallocAsync(&e.d_base, b_draws + b_tiles + b_cnt + b_prep + b_rec, c.stream));
  e.d_draws = reinterpret_cast<Draw *>(e.d_base);
  e.d_tiles = reinterpret_cast<TileDesc *>(e.d_base + b_draws);
  e.d_counters = reinterpret_cast<int *>(e.d_base + b_draws + b_tiles);
  PrepDesc *d_prep = reinterpret_cast<PrepDesc *>(e.d_base + b_draws + b_tiles + b_cnt);
  float *d_rec = reinterpret_cast<float *>(e.d_base + b_draws + b_tiles + b_cnt + b_prep);
  std::vector<TileDesc> tds(tiles.size());
  for (size_t i = 0; i < tiles.size(); ++i) {
    tds[i].src = reinterpret_cast<const double *>(d_rec + tiles[i].off);
    tds[i].bytes = tiles[i].bytes;
    tds[i].pad = 0;
  }
  for (size_t i = 0; i < draws.size(); ++i) draws[i].rec = reinterpret_cast<const double *>(d_rec + draw_off[i]);
  int maxpairs = 1;
  for (size_t i = 0; i < preps.size(); ++i) {
    preps[i].dst = d_rec + prep_off[i];
    if ((preps[i].n + 1) / 2 > maxpairs) maxpairs = (preps[i].n + 1) / 2;
  }
  cudaError_t err = cudaMemcpyAsync(e.d_draws, draws.data(), sizeof(Draw) * draws.size(), cudaMemcpyHostToDevice, c.stream);
  if (err == cudaSuccess) err = cudaMemcpyAsync(e.d_tiles, tds.data(), sizeof(TileDesc) * tds.size(), cudaMemcpyHostToDevice, c.stream);
  if (err == cudaSuccess) err = cudaMemcpyAsync(d_prep, preps.data(), sizeof(PrepDesc) * preps.size(), cudaMemcpyHostToDevice, c.stream);
  if (err == cudaSuccess) err = cudaMemsetAsync(e.d_counters, 0, b_cnt, c.stream);
  if (err == cudaSuccess) {
    gibbs_f32_prep_kernel<<<dim3((unsigned)((maxpairs + 127) / 128), (unsigned)preps.size()), 128, 0, c.stream>>>(d_prep, d, e.map);
    err = cudaGetLastError();
  }
  if (err == cudaSuccess) err = cudaStreamSynchronize(c.stream);
  if (err != cudaSuccess) {
    cudaFreeAsync(e.d_base, c.stream);
    KDE_FAIL(100 + (int)err, "gibbs (FP32): building the schedule: %s", cudaGetErrorString(err));
  }
  lst.push_front(e);
  while (lst.size() > S32_KEEP) {
    cudaDeviceSynchronize();
    s32_free(lst.back(), c.stream);
    lst.pop_back();
  }
  *counter = lst.front().d_counters + (lst.front().next_counter++ % S32_COUNTERS);
  *out = lst.front();
  return 0;
}

static int slow_counter(unsigned long long **out) {
  Context &c = ctx();
  std::lock_guard<std::mutex> lk(g32_mu);
  if (!g32_slow[c.slot]) {
    KDE_CUDA(cudaMallocAsync(&g32_slow[c.slot], sizeof(unsigned long long), c.stream));
    KDE_CUDA(cudaMemsetAsync(g32_slow[c.slot], 0, sizeof(unsigned long long), c.stream));
    KDE_CUDA(cudaStreamSynchronize(c.stream));
  }
  *out = g32_slow[c.slot];
  return 0;
}

// draws redone in FP64 on the calling context's device since the last call of this function (synchronises the device)
int gibbs32_slow_draws(unsigned long long *out) {
  unsigned long long *d = nullptr;
  if (int rc = slow_counter(&d)) return rc;
  KDE_CUDA(cudaDeviceSynchronize());
  KDE_CUDA(cudaMemcpy(out, d, sizeof(*out), cudaMemcpyDeviceToHost));
  KDE_CUDA(cudaMemset(d, 0, sizeof(*out)));
  return 0;
}

template <int D>
static cudaError_t launch32(const GibbsParams &P, int grid_cap, size_t smem, cudaStream_t st, int sm_count) {
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
  if (P.M <= 4) return launch(gibbs_f32_kernel<D, 4>);
  if (P.M <= 8) return launch(gibbs_f32_kernel<D, 8>);
  return launch(gibbs_f32_kernel<D, KDEB200_MAX_DENS>);
}

// P arrives filled by gibbs_device (streams, outputs, masks, root records, labels, sizes); this adds the FP32 schedule
int gibbs32_launch(const kdeb200_tree_t *trees, int ndens, int L, int Niter, bool masked, GibbsParams &P, cudaStream_t st) {
  Context &c = ctx();
  const int d = trees[0]->d;
  Sched32 E;
  int *d_counter = nullptr;
  if (int rc = s32_get(trees, ndens, L, Niter, masked, P.mask, &E, &d_counter)) return rc;
  unsigned long long *d_slow = nullptr;
  if (int rc = slow_counter(&d_slow)) return rc;
  P.draws = E.d_draws;
  P.tiles = E.d_tiles;
  P.counter = d_counter;
  P.ndraws = E.ndraws;
  P.ntiles = E.ntiles;
  P.slow_draws = d_slow;
  for (int k = 0; k < KDEB200_MAX_DIM; ++k) {
    P.nctr[k] = E.map.ctr[k];
    P.nisig[k] = E.map.isig[k];
  }
  const size_t smem = GB_STAGES * GB_TILE_BYTES;
  cudaError_t e = cudaErrorInvalidValue;
  switch (d) {
    case 1: e = launch32<1>(P, P.nbatches, smem, st, c.sm_count); break;
    case 2: e = launch32<2>(P, P.nbatches, smem, st, c.sm_count); break;
    case 3: e = launch32<3>(P, P.nbatches, smem, st, c.sm_count); break;
    case 4: e = launch32<4>(P, P.nbatches, smem, st, c.sm_count); break;
    case 5: e = launch32<5>(P, P.nbatches, smem, st, c.sm_count); break;
    case 6: e = launch32<6>(P, P.nbatches, smem, st, c.sm_count); break;
    case 7: e = launch32<7>(P, P.nbatches, smem, st, c.sm_count); break;
    case 8: e = launch32<8>(P, P.nbatches, smem, st, c.sm_count); break;
  }
  if (e != cudaSuccess) KDE_FAIL(100 + (int)e, "gibbs (FP32) kernel launch: %s", cudaGetErrorString(e));
  return 0;
}

}  // namespace kdeb200
