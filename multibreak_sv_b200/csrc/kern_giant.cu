// Speculative-window kernel: one CTA per (subproblem, chain) (giant_kernel.cuh).
// logBinomial / exp are inlined here: a proposal evaluates two or three of them back to back, independent of each other,
// and only a handful of warps are active, so instruction-level overlap is worth more than code size.
#define MBSV_INLINE_MATH 1
#include "giant_kernel.cuh"
namespace mbsv {

// aln_ent4[a] = the counted slots (aln_esp_slot >= 0) of alignment a in ESP order, -1 padded; {-3,-1,-1,-1} if more than four
__global__ void mbsv_giant_prep_kernel(const __grid_constant__ DevProblem P, int4* __restrict__ out) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= P.n_aln) return;
    const int4 rec = P.aln_rec[a];
    const int ne = rec.w >> 8;
    int s[GIANT_ALN_SLOTS] = {-1, -1, -1, -1};
    int n = 0;
    for (int j = 0; j < ne; ++j) {
        const int slot = P.aln_esp_slot[rec.z + j];
        if (slot >= 0) { if (n < GIANT_ALN_SLOTS) s[n] = slot; ++n; }
    }
    // more counted slots than the record holds: marked, the general kernel then walks aln_esp_slot (the speculative-window
    // kernel never sees such an alignment: sub_maxpos <= giant_max_aln_entries() is a condition of its launch)
    out[a] = n > GIANT_ALN_SLOTS ? make_int4(-3, -1, -1, -1) : make_int4(s[0], s[1], s[2], s[3]);
}

cudaError_t launch_giant_prep(const DevProblem& P, int4* aln_ent4, cudaStream_t stream) {
    mbsv_giant_prep_kernel<<<(P.n_aln + 255) / 256, 256, 0, stream>>>(P, aln_ent4);
    return cudaGetLastError();
}

// cluster = CTAs per chain (1, 2, 4 or 8): launched as a thread-block cluster of that size (cudaLaunchKernelEx)
KernelFn get_kernel_giant(int n_exp, int cluster, int* threads, size_t* smem_bytes) {
    constexpr int NT = 512;
    *threads = NT;
    *smem_bytes = n_exp <= 2 ? GiantLayout<NT, 2>::bytes : GiantLayout<NT, 4>::bytes;
    if (n_exp <= 2) {
        switch (cluster) {
            case 1: return mbsv_giant_kernel<NT, 2, 1>;
            case 2: return mbsv_giant_kernel<NT, 2, 2>;
            case 4: return mbsv_giant_kernel<NT, 2, 4>;
            case 8: return mbsv_giant_kernel<NT, 2, 8>;
        }
    } else {
        switch (cluster) {
            case 1: return mbsv_giant_kernel<NT, 4, 1>;
            case 2: return mbsv_giant_kernel<NT, 4, 2>;
            case 4: return mbsv_giant_kernel<NT, 4, 4>;
            case 8: return mbsv_giant_kernel<NT, 4, 8>;
        }
    }
    return nullptr;
}
int giant_window_offsets() { return 512; }
int giant_max_entries() { return GIANT_MAXENT < GIANT_MAXMV ? GIANT_MAXENT : GIANT_MAXMV; }
int giant_max_aln_entries() { return GIANT_ALN_SLOTS; }
}  // namespace mbsv
