// gibbs_f32_d4.cu -- explicit instantiation of the FP32 Gibbs kernel for d = 4.
#include "gibbs_f32_kernel.cuh"
namespace kdeb200 {
template cudaError_t launch_gibbs_f32_d<4>(const GibbsParams &, int, size_t, cudaStream_t, int);
}
