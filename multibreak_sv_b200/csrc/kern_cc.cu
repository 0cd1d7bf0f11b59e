// Kernels for batches that contain connected-component clusters (GASV localization -1): greedy set cover on
// the device.  One instantiation per (mode, trace, state placement); MAXE = 4 covers every experiment count.
#include "sampler_kernel.cuh"
namespace mbsv {
KernelFn get_kernel_cc(int mode, bool trace, bool smem) {
    if (mode == 0) {
        if (trace) return smem ? mbsv_chain_kernel<0, true, true, 4, true> : mbsv_chain_kernel<0, true, false, 4, true>;
        return smem ? mbsv_chain_kernel<0, false, true, 4, true> : mbsv_chain_kernel<0, false, false, 4, true>;
    }
    if (trace) return smem ? mbsv_chain_kernel<1, true, true, 4, true> : mbsv_chain_kernel<1, true, false, 4, true>;
    return smem ? mbsv_chain_kernel<1, false, true, 4, true> : mbsv_chain_kernel<1, false, false, 4, true>;
}
}  // namespace mbsv
