// MODE_FAST without trace: the throughput kernels and their selection per launch.
#include "kern_fast_select.cuh"
namespace mbsv {
KernelFn get_kernel_fast8(int E, size_t* extra_smem);
// rows = state rows of the shared-memory class (ignored for the global workspace).  narrow_ok = every assignment and count of
// the problem fits a signed char (the host checks).  *state_bytes = size of a state word of the kernel returned.
//
// Measured per class on B200 (profiles/r02_general_kernel_ab.json):
//  * classes of 64 / 96 / 128 rows: with 32-bit words their tiles allow 3-4 CTAs per SM, so they get the registers of 3-4 CTAs
//    per SM (no spills); with 8-bit words the tiles are a quarter, they run 5 CTAs per SM with the totals in shared memory and
//    more L1 is left for the CSR: 74 / 104 / 110 -> 62 / 94 / 98 ms;
//  * classes up to 48 rows: 5 CTAs per SM, totals in shared memory, 32-bit words (byte accesses share banks: 8-bit is 11 % slower);
//  * the global workspace: 32-bit words (8-bit: no gain, 338 -> 342 ms), and 8-bit shared-memory classes of 192..512 rows are
//    slower than leaving those subproblems there (338 -> 375 ms).
KernelFn get_kernel_fast(bool smem, int E, int rows, bool narrow_ok, size_t* extra_smem, int* state_bytes) {
    constexpr bool ST = MBSV_SMTOT != 0;
    *state_bytes = 4;
    if (smem && narrow_ok && rows >= 64) { *state_bytes = 1; return get_kernel_fast8(E, extra_smem); }
    if (smem) return rows >= 128 ? fast_smem_kernel<3, false, int>(E, extra_smem) : rows >= 64 ? fast_smem_kernel<4, false, int>(E, extra_smem)
                                                                                               : fast_smem_kernel<0, ST, int>(E, extra_smem);
    *extra_smem = ST ? (size_t)4096 * (E <= 2 ? E : 4) + 2560 : 0;
    return E == 1 ? mbsv_chain_kernel<1, false, false, 1, false, false, 0, ST> : E == 2 ? mbsv_chain_kernel<1, false, false, 2, false, false, 0, ST>
                                                                                        : mbsv_chain_kernel<1, false, false, 4, false, false, 0, ST>;
}
}  // namespace mbsv
