// Speculative-window kernel: one CTA per (subproblem, chain) (giant_kernel.cuh).
#include "giant_kernel.cuh"
namespace mbsv {
KernelFn get_kernel_giant(int n_exp, int* threads, size_t* smem_bytes) {
    constexpr int NT = 512;
    *threads = NT;
    if (n_exp <= 2) { *smem_bytes = GiantLayout<NT, 2>::bytes; return mbsv_giant_kernel<NT, 2>; }
    *smem_bytes = GiantLayout<NT, 4>::bytes;
    return mbsv_giant_kernel<NT, 4>;
}
int giant_max_entries() { return GIANT_MAXENT < GIANT_MAXMV ? GIANT_MAXENT : GIANT_MAXMV; }
}  // namespace mbsv
