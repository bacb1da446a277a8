// gibbs_d2.cu -- explicit instantiation of the Gibbs kernel for d = 2 (both mask variants).
#include "gibbs_kernel.cuh"
namespace kdeb200 {
template cudaError_t launch_gibbs_d<2>(const GibbsParams &, bool, int, size_t, cudaStream_t, int);
template cudaError_t launch_gibbs_warp_d<2>(const GibbsParams &, bool, int, cudaStream_t);
}
