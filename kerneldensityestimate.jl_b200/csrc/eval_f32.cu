// eval_f32.cu -- FP32 variant of K2 (placeholder until the FP64 path is parity-green on hardware).
#include "tree.cuh"
namespace kdeb200 {
int eval_device_f32(kdeb200_tree_t, const double *, int64_t, int, double *, cudaStream_t, int *) {
  KDE_FAIL(9, "eval: the FP32 variant is not built yet");
}
}  // namespace kdeb200
