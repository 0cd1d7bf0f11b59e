// Global-workspace kernels for warps of ONE chain (a giant component run with few chains): contiguous per-chain state.
// No trace variants: a traced run of such a batch uses the 32-wide layout.
#include "sampler_kernel.cuh"
namespace mbsv {
KernelFn get_kernel_packed(int mode, bool cc) {
    if (mode == 0) return cc ? mbsv_chain_kernel<0, false, false, 4, true, true> : mbsv_chain_kernel<0, false, false, 4, false, true>;
    return cc ? mbsv_chain_kernel<1, false, false, 4, true, true> : mbsv_chain_kernel<1, false, false, 4, false, true>;
}
}  // namespace mbsv
