// common.cuh -- shared host/device helpers of libkdeb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/kdeb200.h"

namespace kdeb200 {

// ---------------------------------------------------------------- errors ----------------
void set_error(const char *fmt, ...);
const char *get_error();

#define KDE_CUDA(call)                                                                             \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      ::kdeb200::set_error("%s: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return 100 + (int)e__;                                                                       \
    }                                                                                              \
  } while (0)

#define KDE_FAIL(code, ...)             \
  do {                                  \
    ::kdeb200::set_error(__VA_ARGS__); \
    return (code);                      \
  } while (0)

// ---------------------------------------------------------------- context ---------------
struct Context {
  bool ready = false;
  int device = 0;
  int slot = 0;  // 0 = primary, 1.. = the other GPUs of the in-process multi-GPU set
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  double *d_exptab = nullptr;  // KDE_EXP_TAB entries of 2^(j/TAB), high word biased (kde_exp_core)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double last_ms = 0.0;
  int last_launches = 0;
};
Context &ctx();           // the context bound to this host thread (ScopedDevice), else the primary
Context &ctx_at(int slot);
int multi_count();        // GPUs in the in-process set (1 unless kdeb200_init_multi was called)
int ensure_init();
// Binds the calling host thread to the context of `slot` (cudaSetDevice + ctx()) for the scope's lifetime.
struct ScopedDevice {
  explicit ScopedDevice(int slot);
  ~ScopedDevice();
  ScopedDevice(const ScopedDevice &) = delete;
  ScopedDevice &operator=(const ScopedDevice &) = delete;
 private:
  void *prev_;
  int prev_dev_ = 0;
};

// ---------------------------------------------------------------- Philox4x32-10 ---------
// Counter-based generator (Salmon et al., SC'11).  Streams are addressed as
//   uniform(sample s, draw c)  : counter = (c_lo, c_hi | 0<<31.., s_lo, s_hi), stream tag 0
//   normal (sample s, slot q)  : stream tag 1
// so the variates a chain consumes do not depend on how samples are sharded across GPUs.
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                        uint32_t k0, uint32_t k1, uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c0;
    const uint64_t p1 = (uint64_t)M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0;
    k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// uniform in [0,1) with 53 random bits, like Julia's rand()
__host__ __device__ __forceinline__ double philox_uniform(uint64_t seed, uint64_t sample, uint32_t draw) {
  uint32_t o[4];
  philox4x32_10(draw, 0u, (uint32_t)sample, (uint32_t)(sample >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), o);
  const uint64_t bits = ((uint64_t)o[0] << 32) | o[1];
  return (double)(bits >> 11) * 0x1.0p-53;
}

#ifdef __CUDACC__
// standard normal by Box-Muller (device only: the host never generates normals itself --
// kdeb200_philox_streams runs this same code on the GPU so tests see bit-identical values)
__device__ __forceinline__ double philox_normal(uint64_t seed, uint64_t sample, uint32_t slot) {
  uint32_t o[4];
  philox4x32_10(slot, 1u, (uint32_t)sample, (uint32_t)(sample >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), o);
  const uint64_t b1 = ((uint64_t)o[0] << 32) | o[1];
  const uint64_t b2 = ((uint64_t)o[2] << 32) | o[3];
  const double u1 = ((double)(b1 >> 11) + 0.5) * 0x1.0p-53;  // (0,1)
  const double u2 = (double)(b2 >> 11) * 0x1.0p-53;          // [0,1)
  return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

// ---------------------------------------------------------------- FP64 exp --------------
#ifndef KDE_EXP_TAB
#define KDE_EXP_TAB 2048
#endif
#if KDE_EXP_TAB == 2048
#define KDE_EXP_K 2954.639443740597       /* 2048/ln2 */
#define KDE_EXP_C1 -0.0003384507717577858   /* -ln2/2048 (rounded) */
#define KDE_EXP_SHL 9                    /* (n >> 11) << 20 == (n & ~2047) << 9 */
#elif KDE_EXP_TAB == 256
#define KDE_EXP_K 369.3299304675746
#define KDE_EXP_C1 -0.0027076061740622863
#define KDE_EXP_SHL 12
#else
#error "KDE_EXP_TAB must be 256 or 2048"
#endif
// The constants travel in the kernel-parameter struct (ExpConsts) so that ptxas can use them as
// c[0x0][..] / uniform-register operands: a DFMA with three distinct REGISTER sources issues at 2/3 rate on
// this part (register-bank limit, tools/micro/ops.cu), one with a constant or immediate operand at full rate.
struct ExpConsts {
  double k;      // TAB/ln2
  double c1;     // -ln2/TAB (rounded)
  double sixth;  // 1/6
};
inline ExpConsts make_exp_consts() { return ExpConsts{KDE_EXP_K, KDE_EXP_C1, 1.6666666666666666e-01}; }

// exp(x) = 2^(n / TAB) * e^r = 2^(n >> log2 TAB) * T[n mod TAB] * e^r,  n = rint(x*TAB/ln2), |r| <= ln2/(2 TAB).
//   TAB = 2048 (16 KB shared memory): degree-3 Taylor (truncation r^4/24 <= 3.4e-17)  -> 7 FP64 instructions
//   TAB = 256                       : degree-4 Taylor (r^5/120 <= 3.8e-17)            -> 8 FP64 instructions
// (libdevice exp(): 17) + 3 integer + 1 LDS instructions.  The reduction uses ONE constant: r = x - n*fl(ln2/TAB)
// is exact up to its own rounding (FMA) and differs from the true remainder by n*(ln2/TAB - fl(.)), a relative
// error of 3.4e-17*|x| in the result (< 1e-15 for every term that can matter in a sum, 2.4e-14 at the
// clamp).  Valid for |x| <= 700 (normal results); anything else must be fixed up by the caller.  Branch-free
// so that independent evaluations interleave in one basic block.  Every DFMA has at most two distinct
// register sources: p = r + (q r) r instead of r + q r^2.
__device__ __forceinline__ double kde_exp_core(double x, const double *__restrict__ tab, const ExpConsts &ec) {
  const double t = __fma_rn(x, ec.k, 6755399441055744.0);  // 1.5*2^52 has zero low word: an FP64 immediate
  const int n = __double2loint(t);
  const double nf = __dadd_rn(t, -6755399441055744.0);
  const double r = __fma_rn(nf, ec.c1, x);
#if KDE_EXP_TAB == 2048
  const double q = __fma_rn(r, ec.sixth, 0.5);
#else
  const double q = __fma_rn(__fma_rn(r, 4.1666666666666664e-02, ec.sixth), r, 0.5);
#endif
  const double p = __fma_rn(__dmul_rn(q, r), r, r);
  // The table stores 2^(j/TAB) with (j << SHL) subtracted from its high word (context.cu), so ONE integer
  // multiply-add, hi + n * 2^SHL, yields 2^(n >> log2 TAB) * 2^(j/TAB) without masking n: the j bits cancel.
  // Scaling T before the FMA instead of y after it is exact (power of two, normal range) and off the critical path.
  const double Tb = tab[n & (KDE_EXP_TAB - 1)];
  const double T = __hiloint2double(__double2hiint(Tb) + n * (1 << KDE_EXP_SHL), __double2loint(Tb));
  return __fma_rn(T, p, T);
}

// Clamped flavour: x < -700 is replaced by -700 with ONE integer instruction on the high word (negative NaN
// becomes tiny too; the reference maps NaN weights to 0, src/MSGibbs01.jl:302).  A term below e^-700 ~ 1e-304
// can never matter in the Gibbs kernel: either pT >= 1e-99 and the term is < 1e-205 of it, or every term is
// that small and the pT < 1e-99 fallback (:311) fires with or without it.  The evaluation kernel uses the
// same clamp and recomputes exactly any row whose total is tiny.  x <= 700 is the caller's contract
// (exponents are ln w plus non-positive terms, plus at most -0.5 sum ln b_k).
__device__ __forceinline__ double kde_exp_flush(double x, const double *__restrict__ tab, const ExpConsts &ec) {
  const unsigned hx = min((unsigned)__double2hiint(x), 0xC085E000u);
  return kde_exp_core(__hiloint2double((int)hx, __double2loint(x)), tab, ec);
}

// reciprocal / reciprocal square root of a normal, positive double without the libdevice
// special-case branches: MUFU seed (2^-22) + one cubic step -> < 1 ulp + rounding
__device__ __forceinline__ double kde_rcp(double c) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(c));
  const double e = __fma_rn(-c, y, 1.0);
  const double t = __fma_rn(e, e, e);
  return __fma_rn(y, t, y);
}
__device__ __forceinline__ double kde_rsqrt(double c) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(c));
  const double e = __fma_rn(-c, __dmul_rn(y, y), 1.0);  // 1 - c y^2
  double u = __fma_rn(e, 0.375, 0.5);
  u = __dmul_rn(u, e);
  return __fma_rn(y, u, y);
}

// ---------------------------------------------------------------- TMA bulk + mbarrier ---
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D bulk copy global -> shared through the TMA unit (SASS: UBLKCP), completion on an mbarrier.
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
#endif  // __CUDACC__

}  // namespace kdeb200
