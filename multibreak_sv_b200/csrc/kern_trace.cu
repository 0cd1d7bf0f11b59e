// Both modes with the per-iteration trace (mcmc.txt columns + accepted-move log); parity tests and
// non---savespace runs.  One instantiation per mode and state placement (MAXE = 4 covers every E).
#include "sampler_kernel.cuh"
namespace mbsv {
KernelFn get_kernel_trace(int mode, bool smem, int E) {
    (void)E;
    if (mode == 0) return smem ? mbsv_chain_kernel<0, true, true, 4> : mbsv_chain_kernel<0, true, false, 4>;
    return smem ? mbsv_chain_kernel<1, true, true, 4> : mbsv_chain_kernel<1, true, false, 4>;
}
}  // namespace mbsv
