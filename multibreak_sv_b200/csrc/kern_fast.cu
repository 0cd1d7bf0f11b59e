// MODE_FAST without trace: the throughput kernels (one instantiation per experiment-count bucket and
// state placement).
#include "sampler_kernel.cuh"
namespace mbsv {
KernelFn get_kernel_fast(bool smem, int E) {
    if (smem) return E == 1 ? mbsv_chain_kernel<1, false, true, 1> : E == 2 ? mbsv_chain_kernel<1, false, true, 2> : mbsv_chain_kernel<1, false, true, 4>;
    return E == 1 ? mbsv_chain_kernel<1, false, false, 1> : E == 2 ? mbsv_chain_kernel<1, false, false, 2> : mbsv_chain_kernel<1, false, false, 4>;
}
}  // namespace mbsv
