// MBSV_MODE_GIBBS: the heat-bath sampler of the statistical tier (G lanes per chain, options across those lanes).
#include "gibbs_kernel.cuh"
namespace mbsv {
KernelFn get_kernel_gibbs(int group, bool global_state) {
    if (global_state) return mbsv_gibbs_kernel<32, true>;
    return group == 4 ? mbsv_gibbs_kernel<4> : group == 8 ? mbsv_gibbs_kernel<8> : mbsv_gibbs_kernel<32>;
}
}  // namespace mbsv
