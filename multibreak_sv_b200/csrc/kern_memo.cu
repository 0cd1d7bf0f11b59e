// Memoised kernels (memo_kernel.cuh): one instantiation without and one with the per-iteration trace.
#include "memo_kernel.cuh"
namespace mbsv {
KernelFn get_kernel_memo(bool trace, bool mixed) {
    return mixed ? (trace ? mbsv_memo_kernel<true, true> : mbsv_memo_kernel<false, true>)
                 : (trace ? mbsv_memo_kernel<true, false> : mbsv_memo_kernel<false, false>);
}
cudaError_t launch_memo_build(const DevProblem& P, const DevRun& R, int n_tab, const long long* state_ptr, const int* tab_sub,
                              long long n_states, int scratch_rows, cudaStream_t stream) {
    if (n_tab <= 0 || n_states <= 0) return cudaSuccess;
    const size_t shm = (size_t)scratch_rows * 128 * sizeof(int);
    if (shm > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute((const void*)mbsv_memo_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
        if (e != cudaSuccess) return e;
    }
    const long long grid = (n_states + 127) / 128;
    mbsv_memo_build_kernel<<<(unsigned)grid, 128, shm, stream>>>(P, R, n_tab, state_ptr, tab_sub, scratch_rows);
    return cudaGetLastError();
}
}  // namespace mbsv
