// MBSV_MODE_GIBBS: the heat-bath sampler of the statistical tier (one warp per chain, options across lanes).
#include "gibbs_kernel.cuh"
namespace mbsv {
KernelFn get_kernel_gibbs() { return mbsv_gibbs_kernel; }
}  // namespace mbsv
