// Memoised kernels (memo_kernel.cuh): one instantiation without and one with the per-iteration trace.
#include "memo_kernel.cuh"
namespace mbsv {
KernelFn get_kernel_memo(bool trace) { return trace ? mbsv_memo_kernel<true> : mbsv_memo_kernel<false>; }
}  // namespace mbsv
