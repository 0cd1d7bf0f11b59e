// MODE_FAST without trace: the throughput kernels (one instantiation per experiment-count bucket, state placement and,
// for the shared-memory classes, register budget: see MINB / SMTOT in sampler_kernel.cuh).
#include "sampler_kernel.cuh"
namespace mbsv {
#ifndef MBSV_SMTOT
#define MBSV_SMTOT 1
#endif
template <int MINB, bool SMTOT>
static KernelFn smem_kernel(int E, size_t* extra) {
    *extra = SMTOT ? (size_t)4096 * (E <= 2 ? E : 4) + 2560 : 0;
    return E == 1 ? mbsv_chain_kernel<1, false, true, 1, false, false, MINB, SMTOT>
         : E == 2 ? mbsv_chain_kernel<1, false, true, 2, false, false, MINB, SMTOT>
                  : mbsv_chain_kernel<1, false, true, 4, false, false, MINB, SMTOT>;
}
// extra_smem = bytes of dynamic shared memory the kernel wants behind the state tiles (SMTOT: counters [5][128] ints, then
// totals and logBinomial terms [4 * MAXE][128] 8-byte words)
KernelFn get_kernel_fast(bool smem, int E, int rows, size_t* extra_smem) {
    constexpr bool ST = MBSV_SMTOT != 0;
    if (smem) return rows >= 128 ? smem_kernel<3, false>(E, extra_smem) : rows >= 64 ? smem_kernel<4, false>(E, extra_smem) : smem_kernel<0, ST>(E, extra_smem);
    *extra_smem = ST ? (size_t)4096 * (E <= 2 ? E : 4) + 2560 : 0;
    return E == 1 ? mbsv_chain_kernel<1, false, false, 1, false, false, 0, ST> : E == 2 ? mbsv_chain_kernel<1, false, false, 2, false, false, 0, ST>
                                                                                        : mbsv_chain_kernel<1, false, false, 4, false, false, 0, ST>;
}
}  // namespace mbsv
