// The MODE_FAST, no-trace instantiations of the lane-per-chain kernel, shared by kern_fast.cu (32-bit state words) and
// kern_fast8.cu (8-bit state words): one per experiment-count bucket, state placement and, for the shared-memory classes,
// register budget (MINB / SMTOT in sampler_kernel.cuh).
#pragma once
#include "sampler_kernel.cuh"
namespace mbsv {
#ifndef MBSV_SMTOT
#define MBSV_SMTOT 1
#endif
// extra = bytes of dynamic shared memory the kernel wants behind the state tiles (SMTOT: counters [5][128] ints, then totals and
// logBinomial terms [4 * MAXE][128] 8-byte words)
template <int MINB, bool SMTOT, typename ST_T>
static KernelFn fast_smem_kernel(int E, size_t* extra) {
    *extra = SMTOT ? (size_t)4096 * (E <= 2 ? E : 4) + 2560 : 0;
    return E == 1 ? mbsv_chain_kernel<1, false, true, 1, false, false, MINB, SMTOT, ST_T>
         : E == 2 ? mbsv_chain_kernel<1, false, true, 2, false, false, MINB, SMTOT, ST_T>
                  : mbsv_chain_kernel<1, false, true, 4, false, false, MINB, SMTOT, ST_T>;
}
}  // namespace mbsv
