// peaks.cu -- pipe-rate microbenchmarks for the roofline denominators:
// dependency-free DFMA, FFMA and MUFU.EX2 streams on every SM, timed with CUDA events.
#include "common.cuh"

namespace kdeb200 {

constexpr int PK_ILP = 8;

template <int WHICH>
__global__ void __launch_bounds__(256) peak_kernel(int iters, double *sink) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (WHICH == 0) {
    double a[PK_ILP];
#pragma unroll
    for (int i = 0; i < PK_ILP; ++i) a[i] = 1.0 + 1e-9 * (tid + i);
    const double b = 0.9999999, c = 1e-7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < PK_ILP; ++i) a[i] = __fma_rn(a[i], b, c);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < PK_ILP; ++i) s += a[i];
    if (s == 123.456) sink[0] = s;
  } else if (WHICH == 1) {
    float a[PK_ILP];
#pragma unroll
    for (int i = 0; i < PK_ILP; ++i) a[i] = 1.0f + 1e-6f * (tid + i);
    const float b = 0.99999f, c = 1e-5f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < PK_ILP; ++i) a[i] = __fmaf_rn(a[i], b, c);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < PK_ILP; ++i) s += a[i];
    if (s == 123.456f) sink[0] = s;
  } else if (WHICH == 2) {
    float a[PK_ILP];
#pragma unroll
    for (int i = 0; i < PK_ILP; ++i) a[i] = -1e-3f * (float)((tid + i) & 1023);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < PK_ILP; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < PK_ILP; ++i) s += a[i];
    if (s == 123.456f) sink[0] = s;
  } else if (WHICH == 3 || WHICH == 4) {  // MUFU.RCP64H / MUFU.RSQ64H seeds
    double a[PK_ILP];
#pragma unroll
    for (int i = 0; i < PK_ILP; ++i) a[i] = 1.0 + 1e-3 * ((tid + i) & 1023);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < PK_ILP; ++i) {
        if (WHICH == 3) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(a[i]));
        else asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(a[i]));
      }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < PK_ILP; ++i) s += a[i];
    if (s == 123.456) sink[0] = s;
  } else {  // DFMA with three distinct register-pair sources (register-bank pressure)
    double a[PK_ILP], b[PK_ILP], c[PK_ILP];
#pragma unroll
    for (int i = 0; i < PK_ILP; ++i) {
      a[i] = 1.0 + 1e-9 * (tid + i);
      b[i] = 0.9999999 + 1e-12 * (tid + i);
      c[i] = 1e-7 + 1e-15 * tid;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < PK_ILP; ++i) a[i] = __fma_rn(a[i], b[i], c[(i + 3) % PK_ILP]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < PK_ILP; ++i) s += a[i] + b[i] + c[i];
    if (s == 123.456) sink[0] = s;
  }
}

// DFMA latency/throughput probe: ILP independent dependent-chains per thread, chosen occupancy
template <int ILP>
__global__ void dfma_probe_kernel(int iters, double *sink) {
  double a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
  const double b = 0.9999999, c = 1e-7;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = __fma_rn(a[i], b, c);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i];
  if (s == 123.456) sink[0] = s;
}

int dfma_probe(int ilp, int blocks_per_sm, int threads, int iters, double *lane_ops_per_s) {
  if (int rc = ensure_init()) return rc;
  Context &c = ctx();
  double *sink = nullptr;
  KDE_CUDA(cudaMalloc(&sink, 8));
  const int blocks = c.sm_count * blocks_per_sm;
  auto run = [&](int n) {
    switch (ilp) {
      case 1: dfma_probe_kernel<1><<<blocks, threads, 0, c.stream>>>(n, sink); break;
      case 2: dfma_probe_kernel<2><<<blocks, threads, 0, c.stream>>>(n, sink); break;
      case 4: dfma_probe_kernel<4><<<blocks, threads, 0, c.stream>>>(n, sink); break;
      case 8: dfma_probe_kernel<8><<<blocks, threads, 0, c.stream>>>(n, sink); break;
      default: dfma_probe_kernel<16><<<blocks, threads, 0, c.stream>>>(n, sink); break;
    }
  };
  run(iters / 8 + 1);
  KDE_CUDA(cudaEventRecord(c.ev0, c.stream));
  run(iters);
  KDE_CUDA(cudaEventRecord(c.ev1, c.stream));
  KDE_CUDA(cudaEventSynchronize(c.ev1));
  KDE_CUDA(cudaGetLastError());
  float ms = 0;
  KDE_CUDA(cudaEventElapsedTime(&ms, c.ev0, c.ev1));
  if (lane_ops_per_s) *lane_ops_per_s = (double)blocks * threads * (double)iters * (ilp > 8 ? 16 : ilp) / (ms * 1e-3);
  cudaFree(sink);
  return 0;
}

int pipe_peak(int which, int iters, double *lane_ops_per_s, double *ms_out) {
  if (int rc = ensure_init()) return rc;
  Context &c = ctx();
  if (which < 0 || which > 5) KDE_FAIL(3, "pipe_peak: which must be 0 (DFMA), 1 (FFMA), 2 (MUFU.EX2), 3 (MUFU.RCP64H), 4 (MUFU.RSQ64H) or 5 (DFMA, 3 register sources)");
  if (iters < 1) iters = 1;
  double *sink = nullptr;
  KDE_CUDA(cudaMalloc(&sink, 8));
  const int blocks = c.sm_count * 8, threads = 256;
  auto run = [&](int n) {
    if (which == 0) peak_kernel<0><<<blocks, threads, 0, c.stream>>>(n, sink);
    else if (which == 1) peak_kernel<1><<<blocks, threads, 0, c.stream>>>(n, sink);
    else if (which == 2) peak_kernel<2><<<blocks, threads, 0, c.stream>>>(n, sink);
    else if (which == 3) peak_kernel<3><<<blocks, threads, 0, c.stream>>>(n, sink);
    else if (which == 4) peak_kernel<4><<<blocks, threads, 0, c.stream>>>(n, sink);
    else peak_kernel<5><<<blocks, threads, 0, c.stream>>>(n, sink);
  };
  run(iters / 8 + 1);  // warm-up
  KDE_CUDA(cudaEventRecord(c.ev0, c.stream));
  run(iters);
  KDE_CUDA(cudaEventRecord(c.ev1, c.stream));
  KDE_CUDA(cudaEventSynchronize(c.ev1));
  KDE_CUDA(cudaGetLastError());
  float ms = 0;
  KDE_CUDA(cudaEventElapsedTime(&ms, c.ev0, c.ev1));
  const double ops = (double)blocks * threads * (double)iters * PK_ILP;
  if (lane_ops_per_s) *lane_ops_per_s = ops / (ms * 1e-3);
  if (ms_out) *ms_out = ms;
  cudaFree(sink);
  return 0;
}

}  // namespace kdeb200
