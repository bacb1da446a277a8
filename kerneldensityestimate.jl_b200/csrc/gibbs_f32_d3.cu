// gibbs_f32_d3.cu -- explicit instantiation of the FP32 Gibbs kernel for d = 3.
#include "gibbs_f32_kernel.cuh"
namespace kdeb200 {
template cudaError_t launch_gibbs_f32_d<3>(const GibbsParams &, int, size_t, cudaStream_t, int);
}
