// context.cu -- process-wide state: device binding, stream, exp table, error strings.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace kdeb200 {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char *get_error() { return g_err; }

Context &ctx() {
  static Context c;
  return c;
}

static std::mutex g_mu;

static int init_device(int device) {
  Context &c = ctx();
  int count = 0;
  KDE_CUDA(cudaGetDeviceCount(&count));
  if (count <= 0) KDE_FAIL(20, "no CUDA device visible: libkdeb200 has no CPU fallback");
  if (device < 0 || device >= count) KDE_FAIL(21, "kdeb200_init: device %d out of range (0..%d)", device, count - 1);
  KDE_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  KDE_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) KDE_FAIL(22, "libkdeb200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
  if (c.ready && c.device == device) return 0;
  if (c.ready) {  // re-bind
    cudaSetDevice(c.device);
    cudaFree(c.d_exptab);
    cudaEventDestroy(c.ev0);
    cudaEventDestroy(c.ev1);
    cudaStreamDestroy(c.stream);
    c.ready = false;
    KDE_CUDA(cudaSetDevice(device));
  }
  c.device = device;
  c.sm_count = prop.multiProcessorCount;
  KDE_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  {  // keep up to 1 GiB of freed blocks in the stream-ordered pool (the default returns everything to the driver
     // at every sync, which made tree create / destroy cost tens of milliseconds)
    cudaMemPool_t pool;
    KDE_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t keep = 1ull << 30;  // bounded: multi-GB staging buffers (injected randU) go back to the driver
    KDE_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  }
  KDE_CUDA(cudaEventCreate(&c.ev0));
  KDE_CUDA(cudaEventCreate(&c.ev1));
  double tab[KDE_EXP_TAB];
  for (int j = 0; j < KDE_EXP_TAB; ++j) {  // 2^(j/TAB), high word biased by -(j << SHL): see kde_exp_core
    const double v = (double)exp2l((long double)j / (long double)KDE_EXP_TAB);
    uint64_t bits;
    std::memcpy(&bits, &v, sizeof(bits));
    bits -= (uint64_t)((uint32_t)j << KDE_EXP_SHL) << 32;
    std::memcpy(&tab[j], &bits, sizeof(bits));
  }
  KDE_CUDA(cudaMalloc(&c.d_exptab, sizeof(tab)));
  KDE_CUDA(cudaMemcpy(c.d_exptab, tab, sizeof(tab), cudaMemcpyHostToDevice));
  c.ready = true;
  return 0;
}

int ensure_init() {
  std::lock_guard<std::mutex> lk(g_mu);
  Context &c = ctx();
  if (c.ready) {
    KDE_CUDA(cudaSetDevice(c.device));
    return 0;
  }
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  return init_device(dev);
}

}  // namespace kdeb200

using namespace kdeb200;

extern "C" {

const char *kdeb200_last_error(void) { return get_error(); }
int kdeb200_version(void) { return 100; }

int kdeb200_device_count(int *count) {
  if (!count) KDE_FAIL(2, "device_count: NULL");
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess) {
    *count = 0;
    KDE_FAIL(100 + (int)e, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
  }
  return 0;
}

int kdeb200_init(int device) {
  std::lock_guard<std::mutex> lk(g_mu);
  return init_device(device);
}

int kdeb200_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  Context &c = ctx();
  if (!c.ready) return 0;
  cudaSetDevice(c.device);
  cudaStreamSynchronize(c.stream);
  cudaFree(c.d_exptab);
  cudaEventDestroy(c.ev0);
  cudaEventDestroy(c.ev1);
  cudaStreamDestroy(c.stream);
  c = Context();
  return 0;
}

int kdeb200_device_props(int *sm_count, int *cc_major, int *cc_minor, int *clock_khz, size_t *free_bytes) {
  if (int rc = ensure_init()) return rc;
  cudaDeviceProp prop;
  KDE_CUDA(cudaGetDeviceProperties(&prop, ctx().device));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  if (clock_khz) {
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx().device);
    *clock_khz = khz;
  }
  if (free_bytes) {
    size_t fr = 0, tot = 0;
    KDE_CUDA(cudaMemGetInfo(&fr, &tot));
    *free_bytes = fr;
  }
  return 0;
}

int kdeb200_last_kernel_ms(double *ms, int *launches) {
  if (ms) *ms = ctx().last_ms;
  if (launches) *launches = ctx().last_launches;
  return 0;
}

}  // extern "C"
