// MODE_FAST without trace, 8-bit state words (ST_T in sampler_kernel.cuh) for the shared-memory classes of 64 rows and more of
// problems whose fragments have at most 127 alignments and whose clusters at most 127 supporting ESPs (see kern_fast.cu).
#include "kern_fast_select.cuh"
namespace mbsv {
KernelFn get_kernel_fast8(int E, size_t* extra_smem) { return fast_smem_kernel<0, MBSV_SMTOT != 0, signed char>(E, extra_smem); }
}  // namespace mbsv
