// MODE_REPLAY_EXACT without trace.
#include "sampler_kernel.cuh"
namespace mbsv {
KernelFn get_kernel_replay(bool smem, int E) {
    if (smem) return E == 1 ? mbsv_chain_kernel<0, false, true, 1> : E == 2 ? mbsv_chain_kernel<0, false, true, 2> : mbsv_chain_kernel<0, false, true, 4>;
    return E == 1 ? mbsv_chain_kernel<0, false, false, 1> : E == 2 ? mbsv_chain_kernel<0, false, false, 2> : mbsv_chain_kernel<0, false, false, 4>;
}
}  // namespace mbsv
