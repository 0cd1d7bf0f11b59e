// Speculative-window sampler for chains that are too few to fill the GPU lane-per-chain: one CTA runs ONE
// (subproblem, chain) — the giant connected component of BASELINE configs[4], or any large subproblem of a small batch.
//
// The reference chain (MultiBreakSV.java:1101-1182) is strictly serial: iteration i+1 starts where iteration i stopped
// drawing from java.util.Random, and how many nextDouble()s an iteration consumes (1 lazy, 4 AlterSV, 5 or 6 Naive)
// depends on the state of the fragment it picked.  A serial walk therefore pays one dependent memory round trip per
// iteration (~2 us on a 500k-fragment state).  Here a CTA of NT threads evaluates a WINDOW of NT consecutive RNG
// offsets at once and stays bit-identical to the serial chain:
//
//   P1  thread j assumes "an iteration starts at offset base + j" (LCG jump-ahead), draws its coins, and for a Naive
//       proposal loads the picked fragment's alignment count and current assignment: that fixes the number of draws
//       c_j of the iteration, i.e. where the next iteration would start (j + c_j).
//   P2  the iterations that really happen are the offsets reachable from offset 0 by j -> j + c_j: warp-level
//       pointer doubling with __shfl, eight-entry transfer tables composed across the warps.  The reachable non-lazy
//       offsets (the proposals) are compacted onto the first warps.
//   P3  every proposal gathers what the Delta-log-posterior needs (alignment records, count slots, counts, logprobcov)
//       with all its loads in flight together, and evaluates its acceptance against the totals at the window start.
//   P4  the per-experiment error/length totals couple every proposal to every earlier accepted move, so the tentative
//       accept flags are turned into exact ones by a prefix sum of the accepted (dE, dL) and ONE re-evaluation with
//       the exact totals each proposal would see; logBinomial(E, L) is a pure function of the totals, so the cached
//       "old" value of the serial chain is reproduced bit for bit.  A proposal whose exact decision differs from its
//       tentative one (its ratio and its accept coin agree to ~1e-9) cuts the window there.
//   P5  a proposal that reads a state row (fragment assignment, cluster count) written by an earlier accepted
//       proposal of the same window was evaluated on stale data: a shared-memory hash set of the written rows finds
//       the first such proposal and cuts the window in front of it.
//   P6  the iterations in front of the cut are committed (state rows, difference tallies, counters), the RNG base moves
//       to the cut (or to the window's exit offset), and the next window starts from the committed state.
//
// Every cut only shortens a window; what is committed is exactly what the serial chain does.  Arithmetic = the FAST
// arithmetic of sampler_kernel.cuh (same expressions, same order).  Not handled here (the planner keeps those on the
// lane-per-chain kernels): REPLAY_EXACT / memo-eligible subproblems, traces, connected components, proposals with more
// than GIANT_MAXENT count entries or GIANT_MAXMV toggled fragments.
//
// initializeMatrix (MultiBreakSV.java:999-1015) is the same trick on single LCG steps: a fragment consumes 2 steps
// (nextDouble() < perr) or 3 (nextDouble, nextInt), so the fragment starts are the offsets reachable by +2 / +3.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "sampler_common.cuh"

namespace mbsv {

constexpr int GIANT_MAXENT = 32;    // count-slot entries of one proposal
constexpr int GIANT_MAXMV = 32;     // fragments toggled by one AlterSV proposal
constexpr int GIANT_ALN_SLOTS = 4;  // counted slots of one alignment (aln_ent4)
constexpr int GIANT_HASH = 4096;    // written-row hash set (power of two)
constexpr int GIANT_TMAX = 2048;    // logprobcov doubles staged in shared memory (E * (max_support + 1) <= this)
constexpr int GIANT_LPCHUNK = 2048; // clusters per chunk of the serial log-posterior sum
constexpr int GIANT_NOCUT = 0x7fffffff;

// State rows of the chain: a contiguous column in the global workspace, read through L2 (ld.global.cg): rows are written
// by other threads of the CTA between windows, and nothing is gained from L1 on random rows.
struct GiantStateGlobal {
    int* st;
    int rows;
#ifdef MBSV_BOUNDS_CHECK
    __device__ __noinline__ void chk(int i) const {
        if (i < 0 || i >= rows) { printf("MBSV_ST(giant): row %d outside [0,%d) (block %d thread %d)\n", i, rows, blockIdx.x, threadIdx.x); __trap(); }
    }
#else
    __device__ __forceinline__ void chk(int) const {}
#endif
    __device__ __forceinline__ int ld(int i) const { chk(i); return __ldcg(st + i); }
    __device__ __forceinline__ void stw(int i, int v) const { chk(i); __stcg(st + i, v); }
    __device__ __forceinline__ void inc(int i) const { chk(i); atomicAdd(st + i, 1); }
};

// 1-D bulk copy global -> shared through the TMA unit (cp.async.bulk, completion on an mbarrier); bytes % 16 == 0.
__device__ __forceinline__ void giant_bulk_load(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(gsrc), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void giant_bar_init(unsigned long long* bar) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void giant_bar_wait(unsigned long long* bar, unsigned phase) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    unsigned done = 0;
    while (!done) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(b), "r"(phase) : "memory");
    }
}

struct GiantReach { bool reach; int rank, drank; };

// Offsets 0..NT-1 of this CTA, one per thread: an item starting at offset j ends at j + c (1 <= c <= 8).  Which offsets are
// reachable?  Per warp, pointer doubling gives for every lane the set of in-warp offsets visited from it and where the walk
// leaves the warp; the first eight lanes publish that as the warp's transfer table (entry offset -> exit offset, items,
// dense items).  giant_reach_tables ends with a __syncthreads; giant_reach_walk composes the tables of the warps in front
// of a thread's warp, starting from the offset `entry` at which the walk enters this CTA (0 for the first CTA of a window)
// with `rank0` items / `drank0` dense items in front of it.
template <int NT>
__device__ __forceinline__ unsigned giant_reach_tables(const int c, const bool dense, unsigned* tab, unsigned* msk) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned S = 1u << lane;
    int J = lane + c;
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        const int src = J < 32 ? J : lane;
        const unsigned S2 = __shfl_sync(0xffffffffu, S, src);
        const int J2 = __shfl_sync(0xffffffffu, J, src);
        if (J < 32) { S |= S2; J = J2; }
    }
    const unsigned dmask = __ballot_sync(0xffffffffu, dense);
    if (lane < 8) {
        tab[w * 8 + lane] = (unsigned)(J - 32) | ((unsigned)__popc(S) << 3) | ((unsigned)__popc(S & dmask) << 9);
        msk[w * 8 + lane] = S;
    }
    __syncthreads();
    return dmask;
}
// the whole CTA as one transfer table entry: entered at offset e, where does the walk leave, how many items / dense items
template <int NT>
__device__ __forceinline__ unsigned giant_reach_cta_entry(int e, const unsigned* tab) {
    int rank = 0, drank = 0;
    for (int k = 0; k < NT / 32; ++k) {
        const unsigned t = tab[k * 8 + e];
        rank += (t >> 3) & 63;
        drank += (t >> 9) & 63;
        e = t & 7;
    }
    return (unsigned)e | ((unsigned)rank << 3) | ((unsigned)drank << 16);      // rank <= NT (13 bits), drank <= NT / 4
}
template <int NT>
__device__ __forceinline__ GiantReach giant_reach_walk(int e, int rank, int drank, const unsigned dmask, const unsigned* tab,
                                                       const unsigned* msk, int* fin) {
    constexpr int NW = NT / 32;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int k = 0; k < w; ++k) {
        const unsigned t = tab[k * 8 + e];
        rank += (t >> 3) & 63;
        drank += (t >> 9) & 63;
        e = t & 7;
    }
    const unsigned mine = msk[w * 8 + e];
    if (fin && w == NW - 1 && lane == 0) {
        const unsigned t = tab[w * 8 + e];
        fin[0] = rank + ((t >> 3) & 63);
        fin[1] = drank + ((t >> 9) & 63);
        fin[2] = t & 7;
    }
    const unsigned lt = (1u << lane) - 1u;
    GiantReach r;
    r.reach = (mine >> lane) & 1u;
    r.rank = rank + __popc(mine & lt);
    r.drank = drank + __popc(mine & dmask & lt);
    return r;
}
template <int NT>
__device__ __forceinline__ GiantReach giant_reach(const int c, const bool dense, unsigned* tab, unsigned* msk, int* fin) {
    const unsigned dmask = giant_reach_tables<NT>(c, dense, tab, msk);
    return giant_reach_walk<NT>(0, 0, 0, dmask, tab, msk, fin);
}

constexpr int GIANT_MAXCS = 8;      // CTAs of a cluster that share one chain's windows

// Shared-memory layout (byte offsets), shared by the kernel and the host's launch code.
template <int NT, int MAXE>
struct GiantLayout {
    static constexpr int NW = NT / 32;
    static constexpr int MD = NT / 4 + 8;          // proposals per CTA per window: every one consumes >= 4 offsets
    static constexpr size_t al(size_t b) { return (b + 15) & ~(size_t)15; }
    static constexpr size_t o_T = 0;
    static constexpr size_t o_lp = o_T + al(sizeof(double) * GIANT_TMAX);
    static constexpr size_t o_wtot = o_lp + al(sizeof(double) * GIANT_LPCHUNK);
    static constexpr size_t o_wlast = o_wtot + al(sizeof(long long) * NW * 2 * MAXE);
    static constexpr size_t o_nls = o_wlast + al(sizeof(int) * NW * MAXE);
    static constexpr size_t o_ent = o_nls + al(sizeof(double) * MAXE * MD);
    static constexpr size_t o_ecnt = o_ent + al(sizeof(int) * GIANT_MAXENT * MD);
    static constexpr size_t o_mv = o_ecnt + al(sizeof(int) * GIANT_MAXENT * MD);
    static constexpr size_t o_hkey = o_mv + al(sizeof(int) * GIANT_MAXMV * MD);
    static constexpr size_t o_hlane = o_hkey + al(sizeof(int) * GIANT_HASH);
    static constexpr size_t o_rtab = o_hlane + al(sizeof(int) * GIANT_HASH);
    static constexpr size_t o_rmsk = o_rtab + al(sizeof(unsigned) * NW * 8);
    static constexpr size_t o_ctab = o_rmsk + al(sizeof(unsigned) * NW * 8);
    static constexpr size_t o_ctot = o_ctab + al(sizeof(unsigned) * GIANT_MAXCS * 8);
    static constexpr size_t o_clast = o_ctot + al(sizeof(long long) * GIANT_MAXCS * 2 * MAXE);
    static constexpr size_t o_sc = o_clast + al(sizeof(int) * GIANT_MAXCS * MAXE);
    static constexpr size_t bytes = o_sc + 512;    // the Scalars block (static_assert in the kernel)
};

// logBinomial for the tentative decisions only: same formula as log_binomial_dev's normal branch with a hardware
// reciprocal and a single-precision logarithm (absolute error ~1e-6).  Whatever it decides is verified with the exact
// arithmetic; a wrong guess only shortens a window.
static __device__ __forceinline__ double giant_lb_approx(long long k, long long n, double p, double logp, double log1mp,
                                                         double log2pi, const double* __restrict__ logtab, int ntab) {
    const double dn = (double)n;
    if (dn * p <= 5 || dn * (1 - p) <= 5) return log_binomial_dev(k, n, p, logp, log1mp, log2pi, logtab, ntab);
    const double mu = dn * p, sig = mu * (1 - p), d = (double)k - mu;
    return -(d * d) * __drcp_rn(2 * sig) - 0.5 * (log2pi + (double)__logf((float)sig));
}

// MAXE = compile-time bound of the experiments (2 or 4).  CS = CTAs per chain: a thread-block cluster of CS CTAs evaluates
// windows of CS * NT offsets, CTA r the offsets [r * NT, (r + 1) * NT); what the CTAs of a window owe each other — transfer
// tables of the reachability walk, prefix sums of the accepted totals, the logBinomial values accepted proposals publish,
// the rows they write, the cut — goes through distributed shared memory, one cluster barrier per exchange.
template <int NT, int MAXE, int CS>
__global__ void __launch_bounds__(NT, 1) mbsv_giant_kernel(const __grid_constant__ DevProblem P, const __grid_constant__ DevRun R) {
    using LY = GiantLayout<NT, MAXE>;
    constexpr int MD = LY::MD, NW = LY::NW;
    static_assert(CS >= 1 && CS <= GIANT_MAXCS, "cluster size");
    extern __shared__ __align__(16) unsigned char giant_smem[];
    double* sT = (double*)(giant_smem + LY::o_T);
    double* lpbuf = (double*)(giant_smem + LY::o_lp);
    long long* wtot = (long long*)(giant_smem + LY::o_wtot);   // [warp][2 * MAXE]: accepted (dE, dL) of the warp
    int* wlast = (int*)(giant_smem + LY::o_wlast);             // [warp][MAXE]: last accepted proposal of the warp touching x
    double* nls = (double*)(giant_smem + LY::o_nls);           // [x][p]: logBinomial of experiment x after accepted proposal p
    int* ent = (int*)(giant_smem + LY::o_ent);        // [e][p]: slot << 1 | (1 = increment)
    int* ecnt = (int*)(giant_smem + LY::o_ecnt);      // [e][p]: count of the slot before the entry applies
    int* mv = (int*)(giant_smem + LY::o_mv);          // [q][p]: AlterSV: fragment << 1 | (1 = currently unmapped)
    int* hkey = (int*)(giant_smem + LY::o_hkey);
    int* hlane = (int*)(giant_smem + LY::o_hlane);
    unsigned* rtab = (unsigned*)(giant_smem + LY::o_rtab);
    unsigned* rmsk = (unsigned*)(giant_smem + LY::o_rmsk);
    unsigned* ctab = (unsigned*)(giant_smem + LY::o_ctab);     // [cta][8]: the CTAs' transfer tables (every CTA holds all)
    long long* ctot = (long long*)(giant_smem + LY::o_ctot);   // [cta][2 * MAXE]: accepted (dE, dL) of the CTA
    int* clast = (int*)(giant_smem + LY::o_clast);             // [cta][MAXE]: last accepted proposal of the CTA touching x
    struct Scalars {
        unsigned long long s_base, jumpA, jumpC, bar;
        long long Etot[MAXE], Ltot[MAXE];
        double LB[MAXE];
        unsigned long long cnt[5];
        int fin[4];
        int iter_base, cut, errlane, done, err, nerr, frag_base;
    };
    static_assert(sizeof(Scalars) <= 512, "Scalars block");
    Scalars& sc = *(Scalars*)(giant_smem + LY::o_sc);

    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = CS > 1 ? (int)cluster.block_rank() : 0;
    auto csync = [&]() { if (CS > 1) cluster.sync(); else __syncthreads(); };
    // the same shared-memory object in CTA k of the cluster (distributed shared memory)
    auto remote = [&](auto* ptr, int k) { return CS > 1 ? cluster.map_shared_rank(ptr, k) : ptr; };

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int goff = crank * NT + tid;                             // this thread's offset inside a window
    const int witem = R.warp_begin + blockIdx.x / CS;
    if (witem >= R.warp_end) return;                               // (whole clusters leave together)
    const long long item = R.item_base + (long long)(witem - R.warp_base);
    const int sub = R.sub_order[item / R.n_chains];
    const int chain = (int)(item % R.n_chains);
    GiantStateGlobal st{R.workspace + (size_t)R.warp_row_off[witem - R.global_warp0], R.sub_W[sub]};

    const int f0 = P.sub_frag_ptr[sub], F = P.sub_frag_ptr[sub + 1] - f0;
    const int c0 = P.sub_cluster_ptr[sub], C = P.sub_cluster_ptr[sub + 1] - c0;
    const int E = P.n_exp;
    const int sv0 = P.sub_sv_ptr[sub], nsv = P.sub_sv_ptr[sub + 1] - sv0;
    const int Tstride = P.max_support + 1;
    const long long scx = (long long)sub * R.n_chains + chain;
    const int N = R.num_iters;
    const int W = R.sub_W[sub];
    const double neglogF = -P.sub_logF[sub];
    const double logperr = R.logperr;
    const double dF = (double)F;
    const unsigned long long perr_thr = prob_threshold53(R.perr);
    constexpr unsigned long long NAIVE_THR = 0x1CCCCCCCCCCCCDULL;   // 0.9 * 2^53
    const int win_base = R.win_base;
    const double neglogNsv = P.sub_neglogNsv[sub];
    const bool tsm = E * Tstride <= GIANT_TMAX;                    // logprobcov staged in shared memory
    const double* Tp = tsm ? sT : P.T;

    auto wrapi = [&](long long t) -> long long { return R.java_int_wrap ? (long long)(int)(unsigned int)(unsigned long long)t : t; };
    auto lb_of = [&](int x, long long e, long long l) -> double {
        return log_binomial_dev(wrapi(e), wrapi(l), P.exp_pseq[x], P.exp_logp[x], P.exp_log1mp[x], P.log2pi, P.logint, P.maxn);
    };
    // slot = cluster * E + experiment (one and two experiments without an integer division)
    auto slot_exp = [&](int slot) -> int { return E == 1 ? 0 : E == 2 ? (slot & 1) : slot % E; };
    auto slot_cluster = [&](int slot) -> int { return E == 1 ? slot : E == 2 ? (slot >> 1) : slot / E; };
    // the set of rows written in a window is ONE hash set spread over the CTAs of the cluster: a key lives in the table of
    // CTA owner_of(key), at slot hash_of(key) (linear probing inside that table)
    auto hash_of = [](int key) -> unsigned { return (((unsigned)key * 2654435761u) >> 19) & (GIANT_HASH - 1); };
    auto owner_of = [](int key) -> int { return CS > 1 ? (int)((((unsigned)key * 2654435761u) >> 16) & (CS - 1)) : 0; };

#ifdef MBSV_GIANT_PROF
    long long sp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int spn = 0;
#define SPROF() do { if (tid == 0 && spn < 8) sp[spn++] = clock64(); } while (0)
#else
#define SPROF() do {} while (0)
#endif
    SPROF();
    // ---- setup: TMA-staged logprobcov, zeroed state, scalars ----
    if (tid == 0) {
        giant_bar_init(&sc.bar);
        JavaRandom r0; r0.seed(R.seed + chain);
        sc.s_base = r0.s;
        sc.iter_base = 0; sc.done = 0; sc.err = 0; sc.nerr = 0; sc.frag_base = 0;
        for (int x = 0; x < MAXE; ++x) { sc.Etot[x] = 0; sc.Ltot[x] = 0; sc.LB[x] = 0.0; }
        for (int i = 0; i < 5; ++i) sc.cnt[i] = 0;
    }
    __syncthreads();
    if (tid == 0 && tsm) {
        // (the device copy of logprobcov is padded to a multiple of 16 bytes)
        const unsigned bytes = (unsigned)(((size_t)E * Tstride * sizeof(double) + 15) & ~(size_t)15);
        giant_bulk_load(sT, P.T, bytes, &sc.bar);
    }
    for (int i = goff; i < W; i += CS * NT) st.stw(i, 0);
    if (tsm) giant_bar_wait(&sc.bar, 0);
    csync();

    // thread j's jump-ahead constants: the state j steps after s is jA * s + jC (one step = step_mul, step_add)
    unsigned long long jA = 1, jC = 0;
    auto make_jump = [&](unsigned long long step_mul, unsigned long long step_add, int steps, int total) {
        jA = 1; jC = 0;
        for (int i = 0; i < steps; ++i) { jA = jA * step_mul; jC = jC * step_mul + step_add; }
        if (tid == NT - 1) {                                             // `total` steps: the whole window
            unsigned long long a = jA, c = jC;
            for (int i = steps; i < total; ++i) { a = a * step_mul; c = c * step_mul + step_add; }
            sc.jumpA = a; sc.jumpC = c;
        }
    };

    SPROF();
    // ---------------- initializeMatrix (MultiBreakSV.java:990-1099): the first CTA of the cluster ----------------
    if (crank == 0) {
        if (R.init_assign) {
            for (int f = tid; f < F; f += NT) {
                const int a = R.init_assign[f0 + f];
                st.stw(f, a);
                if (R.initial_assign) R.initial_assign[(size_t)chain * P.n_frag + f0 + f] = a;
            }
        } else if (R.perr < 0.004) {                 // (windows of ~min(NT, 1/perr) fragments; the general scheme below manages ~NT/3)
            // Almost every fragment consumes three LCG steps (nextDouble() >= perr, then nextInt): thread t takes fragment
            // base + t at offset 3 t.  That is right for every fragment up to and including the first one of the window that
            // draws "unmapped" (two steps) or has to redraw its nextInt; the window is cut behind that fragment.
            make_jump(JavaRandom::jump_mul(3), JavaRandom::jump_add(3), tid, NT);
            __syncthreads();
            for (;;) {
                const int fb = sc.frag_base;
                if (fb >= F) break;
                const unsigned long long sb = sc.s_base;
                const int f = fb + tid;
                const unsigned long long s0 = jA * sb + jC;
                const unsigned long long s1 = s0 * JavaRandom::MULT + JavaRandom::ADD;
                const unsigned long long s2 = s1 * JavaRandom::MULT + JavaRandom::ADD;
                const unsigned long long s3 = s2 * JavaRandom::MULT + JavaRandom::ADD;
                const unsigned long long m53 = (((s1 & JavaRandom::MASK) >> 22) << 27) + ((s2 & JavaRandom::MASK) >> 21);
                if (tid == 0) sc.cut = GIANT_NOCUT;
                __syncthreads();
                int a = -1;
                unsigned long long s_after = s2;                          // unmapped: two steps
                if (f < F) {
                    if (m53 < perr_thr) atomicMin(&sc.cut, tid);
                    else {
                        const int bound = P.frag_aln_ptr[f0 + f + 1] - P.frag_aln_ptr[f0 + f];
                        const int r31 = (int)((s3 & JavaRandom::MASK) >> 17);
                        const int m = bound - 1;
                        s_after = s3;
                        if ((bound & m) == 0) a = (int)(((long long)bound * (long long)r31) >> 31);
                        else {
                            a = r31 % bound;
                            if ((int)((unsigned)r31 - (unsigned)a + (unsigned)m) < 0) {      // java.util.Random.nextInt redraws
                                JavaRandom q; q.s = s2;
                                a = q.nextInt(bound);
                                s_after = q.s;
                                atomicMin(&sc.cut, tid);
                            }
                        }
                    }
                }
                __syncthreads();
                const int cut = sc.cut;
                if (f < F && tid <= cut) {
                    st.stw(f, a);
                    if (R.initial_assign) R.initial_assign[(size_t)chain * P.n_frag + f0 + f] = a;
                }
                const int last = min(min(cut, NT - 1), F - 1 - fb);      // last fragment of this window that is final
                if (tid == last) { sc.s_base = s_after; sc.frag_base = fb + last + 1; }
                __syncthreads();
            }
        } else {
            make_jump(JavaRandom::MULT, JavaRandom::ADD, tid, NT);
            __syncthreads();
            for (;;) {
                const int fb = sc.frag_base;
                if (fb >= F) break;
                const unsigned long long sb = sc.s_base;
                // a fragment whose draws start after offset j: nextDouble() = steps j+1, j+2; nextInt = step j+3
                const unsigned long long s0 = jA * sb + jC;
                const unsigned long long s1 = s0 * JavaRandom::MULT + JavaRandom::ADD;
                const unsigned long long s2 = s1 * JavaRandom::MULT + JavaRandom::ADD;
                const unsigned long long s3 = s2 * JavaRandom::MULT + JavaRandom::ADD;
                const unsigned long long m53 = (((s1 & JavaRandom::MASK) >> 22) << 27) + ((s2 & JavaRandom::MASK) >> 21);
                const bool unmapped = m53 < perr_thr;                     // nextDouble() < perr
                if (tid == 0) sc.cut = GIANT_NOCUT;
                const GiantReach rr = giant_reach<NT>(unmapped ? 2 : 3, false, rtab, rmsk, sc.fin);
                const int f = fb + rr.rank;
                int a = -1;
                bool reject = false;
                unsigned long long s_after = s0;
                if (rr.reach && f < F && !unmapped) {
                    const int bound = P.frag_aln_ptr[f0 + f + 1] - P.frag_aln_ptr[f0 + f];
                    const int r31 = (int)((s3 & JavaRandom::MASK) >> 17);
                    const int m = bound - 1;
                    if ((bound & m) == 0) a = (int)(((long long)bound * (long long)r31) >> 31);
                    else {
                        a = r31 % bound;
                        // java.util.Random.nextInt redraws when u - r + m overflows int (about 1e-8 of the draws): serial
                        if ((int)((unsigned)r31 - (unsigned)a + (unsigned)m) < 0) {
                            reject = true;
                            JavaRandom q; q.s = s2;
                            a = q.nextInt(bound);
                            s_after = q.s;
                            atomicMin(&sc.cut, tid);
                        }
                    }
                }
                if (rr.reach && f == F) atomicMin(&sc.cut, tid);            // first offset after the last fragment
                __syncthreads();
                const int cut = sc.cut;
                // fragments in front of the cut are final; a redrawing fragment AT the cut has finished itself serially
                if (rr.reach && f < F && (tid < cut || (tid == cut && reject))) {
                    st.stw(f, a);
                    if (R.initial_assign) R.initial_assign[(size_t)chain * P.n_frag + f0 + f] = a;
                }
                if (cut == GIANT_NOCUT) {
                    if (tid == 0) {
                        unsigned long long s = sc.jumpA * sb + sc.jumpC;
                        for (int i = 0; i < sc.fin[2]; ++i) s = s * JavaRandom::MULT + JavaRandom::ADD;
                        sc.s_base = s;
                        sc.frag_base = fb + sc.fin[0];
                    }
                } else if (tid == cut) {
                    if (reject) { sc.s_base = s_after; sc.frag_base = f + 1; }
                    else { sc.s_base = s0; sc.frag_base = f; }              // f == F: the state after the last fragment's draws
                }
                __syncthreads();
            }
        }
    }
    csync();
    SPROF();
    // counts, per-experiment totals, unmapped fragments (all CTAs; pooled in the first CTA's scalars)
    {
        long long e_loc[MAXE], l_loc[MAXE];
#pragma unroll
        for (int x = 0; x < MAXE; ++x) { e_loc[x] = 0; l_loc[x] = 0; }
        int nerr_loc = 0;
        constexpr int IB = 8;
        for (int fb = goff * IB; fb < F; fb += CS * NT * IB) {
            int a[IB], ap[IB];
            int4 rec[IB];
#pragma unroll
            for (int k = 0; k < IB; ++k) a[k] = (fb + k < F) ? st.ld(fb + k) : -2;
#pragma unroll
            for (int k = 0; k < IB; ++k) ap[k] = a[k] >= 0 ? __ldg(P.frag_aln_ptr + f0 + fb + k) : 0;
#pragma unroll
            for (int k = 0; k < IB; ++k) rec[k] = a[k] >= 0 ? __ldg(P.aln_rec + ap[k] + a[k]) : make_int4(0, 0, 0, 0);
#pragma unroll
            for (int k = 0; k < IB; ++k) {
                if (a[k] == -1) ++nerr_loc;
                if (a[k] < 0) continue;
                const int x = rec[k].w & 0xff, ne = rec[k].w >> 8;
                for (int jj = 0; jj < ne; ++jj) {
                    const int slot = P.aln_esp_slot[rec[k].z + jj];
                    if (slot >= 0) st.inc(F + slot);
                }
#pragma unroll
                for (int y = 0; y < MAXE; ++y) if (y == x) { e_loc[y] += rec[k].x; l_loc[y] += rec[k].y; }
            }
        }
        Scalars* s0p = remote(&sc, 0);
#pragma unroll
        for (int x = 0; x < MAXE; ++x) {
            if (e_loc[x]) atomicAdd((unsigned long long*)&s0p->Etot[x], (unsigned long long)e_loc[x]);
            if (l_loc[x]) atomicAdd((unsigned long long*)&s0p->Ltot[x], (unsigned long long)l_loc[x]);
        }
        if (nerr_loc) atomicAdd(&s0p->nerr, nerr_loc);
    }
    __threadfence();
    csync();

    // MultiBreakSVAssignmentMatrix.java:130-246 in the reference's summation order: the per-cluster terms of a chunk are
    // gathered by the whole CTA, the chunk is added up by one thread (floating-point addition does not reassociate).
    // (first CTA of the cluster only)
    auto full_logprob = [&]() -> double {
        double lp = 0;                                                    // (meaningful in thread 0 only)
        for (int i0 = 0; i0 < C; i0 += GIANT_LPCHUNK) {
            const int nchunk = min(GIANT_LPCHUNK, C - i0);
            for (int i = tid; i < nchunk; i += NT) {
                const int cl = P.supports_order_local[c0 + i0 + i];
                double lpc = 0.0;
                for (int x = 0; x < E; ++x) {
                    const int n = st.ld(F + cl * E + x);
                    if (n > 0) lpc = lpc + Tp[(size_t)x * Tstride + n];
                }
                lpbuf[i] = lpc;
            }
            __syncthreads();
            if (tid == 0) for (int i = 0; i < nchunk; ++i) lp += lpbuf[i];
            __syncthreads();
        }
        if (tid == 0) {
            lp += 0.0;
            lp += (double)sc.nerr * logperr;
            for (int x = 0; x < E; ++x) lp += lb_of(x, sc.Etot[x], sc.Ltot[x]);
        }
        return lp;
    };
    if (R.initial_logprob && crank == 0) {
        const double lp = full_logprob();
        if (tid == 0) R.initial_logprob[scx] = lp;
    }

    SPROF();
    // ---------------- iterations (MultiBreakSV.java:796-837, 1101-1182) ----------------
    make_jump(JavaRandom::MULT2, JavaRandom::ADD2, goff, CS * NT);        // one offset = one nextDouble()
    int n_naive_prop = 0, n_naive_acc = 0, n_sv_prop = 0, n_sv_acc = 0;
    long long n_reassign = 0;
    int nerr_delta = 0;                                                   // unmapped fragments: net change committed by this thread
    // the chain's scalars live in every CTA of the cluster; the first CTA holds them now
    if (crank == 0 && tid == 0) {
        // logBinomial of the current totals (the chain's cached "old" value): computed once, then handed from window to window
        for (int x = 0; x < E; ++x) sc.LB[x] = lb_of(x, sc.Etot[x], sc.Ltot[x]);
        sc.done = N <= 0;
        for (int k = 1; k < CS; ++k) {
            Scalars* o = remote(&sc, k);
            o->s_base = sc.s_base; o->iter_base = 0; o->done = sc.done; o->err = 0;
            for (int x = 0; x < MAXE; ++x) { o->Etot[x] = sc.Etot[x]; o->Ltot[x] = sc.Ltot[x]; o->LB[x] = sc.LB[x]; }
        }
    }
    csync();
#ifdef MBSV_GIANT_PROF
    long long prof[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, pc = 0;
    long long nwindows = 0, ncuts = 0;
#define GPROF(i) do { if (tid == 0) { const long long c_ = clock64(); prof[i] += c_ - pc; pc = c_; } } while (0)
    if (tid == 0) pc = clock64();
#else
#define GPROF(i) do {} while (0)
#endif
    // a cut found by any CTA is posted in every CTA's scalars (cuts are rare: a handful of remote atomics per window)
    auto post_cut = [&](int off) {
#pragma unroll
        for (int k = 0; k < CS; ++k) atomicMin(&remote(&sc, k)->cut, off);
    };
    while (!sc.done) {
        // ================= P1: one iteration per offset =================
        const unsigned long long sb = sc.s_base;
        const int iter_base = sc.iter_base;
        if (tid == 0) { sc.cut = GIANT_NOCUT; sc.errlane = GIANT_NOCUT; }
        JavaRandom rng;
        rng.s = jA * sb + jC;
        int c = 1;
        bool dense = false, naive = false;
        int f = 0;                                                        // Naive: the fragment; AlterSV: the candidate
        unsigned long long m53 = 0, s_mid = 0;                            // s_mid: state after the fragment / candidate pick
        {
            const unsigned long long m0 = rng.bits53();
            if (m0 >= (1ULL << 52)) {                                     // !(nextDouble() < 0.5)
                dense = true;
                naive = true;
                if (nsv == 0) rng.skip<2>(); else naive = rng.bits53() < NAIVE_THR;
                const double u = rng.nextDouble();
                s_mid = rng.s;
                if (naive) { f = (int)(u * dF); m53 = rng.bits53(); } else { f = sv0 + (int)(u * (double)nsv); c = 4; }
            }
        }
        int a0 = 0, fa1 = 0, prev = -1;
        int sv_q0 = 0, sv_q1 = 0, sv_e0 = 0, sv_e1 = 0, sv_nch = 0;       // AlterSV: the candidate's fragment list and entry list
        if (naive) { a0 = __ldg(P.frag_aln_ptr + f0 + f); fa1 = __ldg(P.frag_aln_ptr + f0 + f + 1); prev = st.ld(f); }
        else if (dense) {
            sv_q0 = __ldg(P.sv_frag_ptr + f); sv_q1 = __ldg(P.sv_frag_ptr + f + 1);
            sv_e0 = __ldg(P.sv_ent_ptr + f); sv_e1 = __ldg(P.sv_ent_ptr + f + 1);
            sv_nch = __ldg(P.sv_numchanged + f);
        }
        // (while those loads are in flight)
        for (int i = tid; i < GIANT_HASH; i += NT) { hkey[i] = -1; hlane[i] = GIANT_NOCUT; }
        const int n = fa1 - a0;
        if (naive) {
            const bool from_err = prev == -1;
            const bool to_err = !from_err && (m53 < perr_thr || n == 1);
            c = 5 + ((!from_err && !to_err) ? 1 : 0);
        }
        GPROF(0);
        // ================= P2: which offsets are iterations =================
        const unsigned dmask = giant_reach_tables<NT>(c, dense, rtab, rmsk);
        int entry = 0, rank0 = 0, drank0 = 0;
        int cbase[CS + 1];                                                // proposals in front of every CTA of the window
        cbase[0] = 0;
        if (CS > 1) {
            // every CTA publishes its transfer table (entered at offset e -> exit, items, proposals) in every CTA
            if (tid < 8 * CS) remote(ctab, tid >> 3)[crank * 8 + (tid & 7)] = giant_reach_cta_entry<NT>(tid & 7, rtab);
            cluster.sync();
            int e = 0, rk = 0, dr = 0;
#pragma unroll
            for (int k = 0; k < CS; ++k) {
                if (k == crank) { entry = e; rank0 = rk; drank0 = dr; }
                const unsigned t = ctab[k * 8 + e];
                rk += (t >> 3) & 8191;
                dr += t >> 16;
                e = t & 7;
                cbase[k + 1] = dr;
            }
            if (tid == 0) { sc.fin[0] = rk; sc.fin[1] = dr; sc.fin[2] = e; }     // the whole window: iterations, proposals, exit
        }
        const GiantReach rr = giant_reach_walk<NT>(entry, rank0, drank0, dmask, rtab, rmsk, CS > 1 ? nullptr : sc.fin);
        const bool act = rr.reach && dense;                               // this thread's offset is a proposal of the chain
        const int pg = rr.drank;                                          // its index among the window's proposals
        const int p = pg - drank0;                                        // its slot in this CTA's per-proposal arrays
        const int t_iter = iter_base + rr.rank;
        if (rr.reach && t_iter == N) post_cut(goff);                      // the run ends in front of this offset
        GPROF(1);
        // ================= P3: gather + tentative decision =================
        int now = -1, nmoves = 0, nent = 0, q0 = 0;
        bool bad = false, flagA = false, flagB = false;
        double dcov = 0.0, derr = 0.0, Alogq = 0.0, Anewlogq = 0.0, r_acc = 0.0;
        long long dE[MAXE], dL[MAXE];
#pragma unroll
        for (int x = 0; x < MAXE; ++x) { dE[x] = 0; dL[x] = 0; }
        unsigned touched = 0;
        int dnerr = 0;
        long long reassign = 0;
        if (act) {
            rng.s = s_mid;
            if (naive) {
                naive_pick(rng, n, prev, neglogF, perr_thr, logperr, R.log1mperr, P.logint, P.invint, now, Alogq, Anewlogq);
                reassign = 1;
                if (now == prev) {                                        // Error at MultiBreakSV.java:1122-1123
                    bad = true;
                    post_cut(goff);
#pragma unroll
                    for (int k = 0; k < CS; ++k) atomicMin(&remote(&sc, k)->errlane, goff);
                } else {
                    // one round trip: both alignment records and their counted slots
                    int4 recF = make_int4(0, 0, 0, 0), recT = make_int4(0, 0, 0, 0);
                    int4 slF = make_int4(-1, -1, -1, -1), slT = make_int4(-1, -1, -1, -1);
                    if (prev != -1) { recF = __ldg(P.aln_rec + a0 + prev); slF = __ldg(P.aln_ent4 + a0 + prev); }
                    if (now != -1) { recT = __ldg(P.aln_rec + a0 + now); slT = __ldg(P.aln_ent4 + a0 + now); }
                    if (prev != -1) {
                        const int x = recF.w & 0xff;
                        const int sl[4] = {slF.x, slF.y, slF.z, slF.w};
#pragma unroll
                        for (int k = 0; k < 4; ++k) if (sl[k] >= 0) { ent[nent * MD + p] = sl[k] << 1; ++nent; }
#pragma unroll
                        for (int y = 0; y < MAXE; ++y) if (y == x) { dE[y] -= recF.x; dL[y] -= recF.y; }
                        touched |= 1u << x;
                    } else { derr = derr - logperr; --dnerr; }
                    if (now != -1) {
                        const int x = recT.w & 0xff;
                        const int sl[4] = {slT.x, slT.y, slT.z, slT.w};
#pragma unroll
                        for (int k = 0; k < 4; ++k) if (sl[k] >= 0) { ent[nent * MD + p] = (sl[k] << 1) | 1; ++nent; }
#pragma unroll
                        for (int y = 0; y < MAXE; ++y) if (y == x) { dE[y] += recT.x; dL[y] += recT.y; }
                        touched |= 1u << x;
                    } else { derr = derr + logperr; ++dnerr; }
                }
            } else {
                // ---- svMove (MultiBreakSV.java:1291-1343): toggle the cluster's single-alignment fragments ----
                // (the candidate's fragment list and entry list were loaded with the window's first round trip)
                q0 = sv_q0;
                const int e0 = sv_e0;
                reassign = sv_nch;
                nmoves = sv_q1 - sv_q0;
                nent = sv_e1 - sv_e0;
                Alogq = neglogNsv;
                Anewlogq = neglogNsv;
                // round trip 1: fragments and slots (four loads in flight at a time)
                for (int qb = 0; qb < nmoves; qb += 4) {
                    int g[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) g[i] = qb + i < nmoves ? __ldg(&P.sv_mv[q0 + qb + i].x) : 0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) if (qb + i < nmoves) mv[(qb + i) * MD + p] = g[i];
                }
                for (int kb = 0; kb < nent; kb += 4) {
                    int2 v[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[i] = kb + i < nent ? __ldg(P.sv_ent + e0 + kb + i) : make_int2(0, 0);
#pragma unroll
                    for (int i = 0; i < 4; ++i) if (kb + i < nent) ent[(kb + i) * MD + p] = (v[i].x << 5) | v[i].y;
                }
                // round trip 2: current assignment of every fragment and current count of every slot, together
                for (int qb = 0; qb < nmoves; qb += 4) {
                    int a[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) a[i] = qb + i < nmoves ? st.ld(mv[(qb + i) * MD + p]) : 0;
                    if (qb == 0) {
                        for (int kb = 0; kb < nent; kb += 4) {
                            int v[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) v[i] = kb + i < nent ? st.ld(F + (ent[(kb + i) * MD + p] >> 5)) : 0;
#pragma unroll
                            for (int i = 0; i < 4; ++i) if (kb + i < nent) ecnt[(kb + i) * MD + p] = v[i];
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) if (qb + i < nmoves) mv[(qb + i) * MD + p] = (mv[(qb + i) * MD + p] << 1) | (a[i] == -1 ? 1 : 0);
                }
                for (int k = 0; k < nent; ++k) {                          // direction of every entry = direction of its fragment
                    const int v = ent[k * MD + p];
                    ent[k * MD + p] = ((v >> 5) << 1) | (mv[(v & 31) * MD + p] & 1);
                }
                for (int q = 0; q < nmoves; ++q) {                        // (records again: cached)
                    const int4 m = __ldg(P.sv_mv + q0 + q);
                    const int x = m.w;
                    if (mv[q * MD + p] & 1) {                             // -1 -> 0
                        derr = derr - logperr; --dnerr;
#pragma unroll
                        for (int y = 0; y < MAXE; ++y) if (y == x) { dE[y] += m.y; dL[y] += m.z; }
                    } else {                                              // 0 -> -1
#pragma unroll
                        for (int y = 0; y < MAXE; ++y) if (y == x) { dE[y] -= m.y; dL[y] -= m.z; }
                        derr = derr + logperr; ++dnerr;
                    }
                    touched |= 1u << x;
                }
            }
            r_acc = rng.nextDouble();
            if (!bad) {
                // counts the entries see: the stored count plus the earlier entries of this proposal on the same slot
                for (int kb = 0; naive && kb < nent; kb += 4) {            // (AlterSV: loaded above with the assignments)
                    int v[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[i] = kb + i < nent ? st.ld(F + (ent[(kb + i) * MD + p] >> 1)) : 0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) if (kb + i < nent) ecnt[(kb + i) * MD + p] = v[i];
                }
                for (int k = 0; k < nent; ++k) {
                    const int ek = ent[k * MD + p];
                    int cn = ecnt[k * MD + p];
                    for (int k2 = 0; k2 < k; ++k2) {
                        const int e2 = ent[k2 * MD + p];
                        if ((e2 >> 1) == (ek >> 1)) cn += (e2 & 1) ? 1 : -1;
                    }
                    ecnt[k * MD + p] = cn;
                    const double* Tx = Tp + (size_t)slot_exp(ek >> 1) * Tstride;
                    dcov = dcov + (Tx[cn + ((ek & 1) ? 1 : -1)] - Tx[cn]);
                }
                // tentative decision against the totals at the window start (approximate arithmetic, see giant_lb_approx)
                double dlb = 0.0;
                unsigned mask = touched & ((1u << E) - 1u);
                while (mask) {
                    const int x = __ffs(mask) - 1;
                    mask &= mask - 1;
                    long long de = 0, dl = 0;
#pragma unroll
                    for (int y = 0; y < MAXE; ++y) if (y == x) { de = dE[y]; dl = dL[y]; }
                    dlb += giant_lb_approx(wrapi(sc.Etot[x] + de), wrapi(sc.Ltot[x] + dl), P.exp_pseq[x], P.exp_logp[x], P.exp_log1mp[x],
                                           P.log2pi, P.logint, P.maxn) -
                           giant_lb_approx(wrapi(sc.Etot[x]), wrapi(sc.Ltot[x]), P.exp_pseq[x], P.exp_logp[x], P.exp_log1mp[x],
                                           P.log2pi, P.logint, P.maxn);
                }
                const double arg = ((dcov + derr) + dlb + Alogq) - Anewlogq;
                const double e = (double)__expf((float)arg);
                flagA = fmin(1.0, e) > r_acc;                             // (NaN: not accepted, like the exact test)
            }
        }
        GPROF(2);
        // ================= P4: exact totals per proposal: prefix sums over the tentatively accepted ones =================
        long long pre[2 * MAXE];                                         // inclusive prefix of (dE, dL)
        int last[MAXE];                                                   // last accepted proposal touching x, this one included
#pragma unroll
        for (int x = 0; x < MAXE; ++x) {
            pre[x] = flagA ? dE[x] : 0; pre[MAXE + x] = flagA ? dL[x] : 0;
            last[x] = (flagA && ((touched >> x) & 1u)) ? pg : -1;
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
            for (int v = 0; v < 2 * MAXE; ++v) {
                const long long y = __shfl_up_sync(0xffffffffu, pre[v], o);
                if (lane >= o) pre[v] += y;
            }
#pragma unroll
            for (int x = 0; x < MAXE; ++x) {
                const int y = __shfl_up_sync(0xffffffffu, last[x], o);
                if (lane >= o) last[x] = max(last[x], y);
            }
        }
        if (lane == 31) {
#pragma unroll
            for (int v = 0; v < 2 * MAXE; ++v) wtot[wid * 2 * MAXE + v] = pre[v];
#pragma unroll
            for (int x = 0; x < MAXE; ++x) wlast[wid * MAXE + x] = last[x];
        }
        // exclusive within the warp
        long long Ecur[MAXE], Lcur[MAXE];
        int lastx[MAXE];
#pragma unroll
        for (int x = 0; x < MAXE; ++x) {
            Ecur[x] = pre[x] - (flagA ? dE[x] : 0);
            Lcur[x] = pre[MAXE + x] - (flagA ? dL[x] : 0);
            const int y = __shfl_up_sync(0xffffffffu, last[x], 1);
            lastx[x] = lane > 0 ? y : -1;
        }
        __syncthreads();
        for (int k = 0; k < wid; ++k) {
#pragma unroll
            for (int x = 0; x < MAXE; ++x) {
                Ecur[x] += wtot[k * 2 * MAXE + x];
                Lcur[x] += wtot[k * 2 * MAXE + MAXE + x];
                lastx[x] = max(lastx[x], wlast[k * MAXE + x]);
            }
        }
        if (CS > 1) {
            // the CTA's own totals go to every CTA; the ones of the CTAs in front are added
            if (tid == NT - 1) {
#pragma unroll
                for (int k = 0; k < CS; ++k) {
                    long long* ct = remote(ctot, k) + crank * 2 * MAXE;
                    int* cl = remote(clast, k) + crank * MAXE;
#pragma unroll
                    for (int x = 0; x < MAXE; ++x) {
                        ct[x] = Ecur[x] + (flagA ? dE[x] : 0);
                        ct[MAXE + x] = Lcur[x] + (flagA ? dL[x] : 0);
                        cl[x] = max(lastx[x], (flagA && ((touched >> x) & 1u)) ? pg : -1);
                    }
                }
            }
            cluster.sync();
            for (int k = 0; k < crank; ++k) {
#pragma unroll
                for (int x = 0; x < MAXE; ++x) {
                    Ecur[x] += ctot[k * 2 * MAXE + x];
                    Lcur[x] += ctot[k * 2 * MAXE + MAXE + x];
                    lastx[x] = max(lastx[x], clast[k * MAXE + x]);
                }
            }
        }
#pragma unroll
        for (int x = 0; x < MAXE; ++x) { Ecur[x] += sc.Etot[x]; Lcur[x] += sc.Ltot[x]; }   // the totals this offset would see
        GPROF(3);
        // logBinomial of the totals AFTER this proposal; the accepted ones publish it: it is the "old" value of whoever follows
        double nl[MAXE];
#pragma unroll
        for (int x = 0; x < MAXE; ++x) nl[x] = 0.0;
        if (act && !bad) {
            unsigned mask = touched & ((1u << E) - 1u);
            while (mask) {
                const int x = __ffs(mask) - 1;
                mask &= mask - 1;
                long long ex = 0, lx = 0;
#pragma unroll
                for (int y = 0; y < MAXE; ++y) if (y == x) { ex = Ecur[y] + dE[y]; lx = Lcur[y] + dL[y]; }
                const double v = lb_of(x, ex, lx);
#pragma unroll
                for (int y = 0; y < MAXE; ++y) if (y == x) nl[y] = v;
                if (flagA) nls[x * MD + p] = v;
            }
        }
        csync();
        GPROF(4);
        // the value proposal `q` (window index) published for experiment x: in the shared memory of the CTA that owns q
        auto published = [&](int x, int q) -> double {
            int k = 0;
#pragma unroll
            for (int kk = 1; kk < CS; ++kk) if (q >= cbase[kk]) k = kk;
            return remote(nls, k)[x * MD + (q - cbase[k])];
        };
        if (act && !bad) {
            // ---- the exact decision (MultiBreakSV.java:1146), FAST arithmetic ----
            double dlb = 0.0;
            unsigned mask = touched & ((1u << E) - 1u);
            while (mask) {
                const int x = __ffs(mask) - 1;
                mask &= mask - 1;
                double nlx = 0.0;
                int lx = -1;
#pragma unroll
                for (int y = 0; y < MAXE; ++y) if (y == x) { nlx = nl[y]; lx = lastx[y]; }
                const double old = lx >= 0 ? published(x, lx) : sc.LB[x];
                dlb = dlb + (nlx - old);
            }
            const double delta = (dcov + derr) + dlb;
            const double arg = (delta + Alogq) - Anewlogq;
            const double e = exp_dev(arg);
            const double ratio = (e != e) ? e : (1.0 < e ? 1.0 : e);      // Math.min(1, e)
            flagB = ratio > r_acc;
            if (flagB != flagA) post_cut(goff);                           // the hypothesis fails here: redo from this iteration
        }
        // ================= P5: stale reads.  Accepted proposals publish the rows they write ... =================
        const int nfr = naive ? 1 : nmoves;
        const int nkeys = (act && !bad) ? nfr + nent : 0;
        auto key_at = [&](int i) -> int { return i < nfr ? (naive ? f : (mv[i * MD + p] >> 1)) : F + (ent[(i - nfr) * MD + p] >> 1); };
        if (flagB) {
            bool ok = true;
            for (int kb = 0; kb < nkeys; kb += 4) {
                int key[4], old[4];
                unsigned h[4];
                int* hk[4];
                int* hl[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) if (kb + i < nkeys) {
                    key[i] = key_at(kb + i); h[i] = hash_of(key[i]);
                    const int ow = owner_of(key[i]);
                    hk[i] = remote(hkey, ow); hl[i] = remote(hlane, ow);
                    old[i] = atomicCAS(&hk[i][h[i]], -1, key[i]);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) if (kb + i < nkeys) {
                    bool placed = old[i] == -1 || old[i] == key[i];
                    for (int guard = 0; !placed && guard < GIANT_HASH; ++guard) {
                        h[i] = (h[i] + 1) & (GIANT_HASH - 1);
                        const int o = atomicCAS(&hk[i][h[i]], -1, key[i]);
                        placed = o == -1 || o == key[i];
                    }
                    if (placed) atomicMin(&hl[i][h[i]], goff); else ok = false;
                }
            }
            // set full (cannot happen below ~4000 written rows per window): nothing behind this proposal is committed
            if (!ok && goff + c < CS * NT) post_cut(goff + c);
        }
        csync();
        GPROF(5);
        // ... and the first proposal that read a row written in front of it cuts the window
        if (nkeys) {
            bool stale = false;
            for (int kb = 0; kb < nkeys; kb += 4) {
                int key[4], k0[4], l0[4];
                unsigned h[4];
                const int* hk[4];
                const int* hl[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) if (kb + i < nkeys) {
                    key[i] = key_at(kb + i); h[i] = hash_of(key[i]);
                    const int ow = owner_of(key[i]);
                    hk[i] = remote(hkey, ow); hl[i] = remote(hlane, ow);
                    k0[i] = hk[i][h[i]]; l0[i] = hl[i][h[i]];
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) if (kb + i < nkeys) {
                    int k = k0[i], l = l0[i];
                    for (int guard = 0; k != -1 && k != key[i] && guard < GIANT_HASH; ++guard) {
                        h[i] = (h[i] + 1) & (GIANT_HASH - 1);
                        k = hk[i][h[i]]; l = hl[i][h[i]];
                    }
                    if (k == key[i] && l < goff) stale = true;
                }
            }
            if (stale) post_cut(goff);
        }
        csync();
        GPROF(6);
        // ================= P6: commit the iterations in front of the cut =================
        const int cut = sc.cut;
        const int errlane = sc.errlane;
        if (act && goff < cut) {
            if (naive) ++n_naive_prop; else ++n_sv_prop;
            n_reassign += reassign;
            if (flagB) {
                if (naive) ++n_naive_acc; else ++n_sv_acc;
                nerr_delta += dnerr;
                const unsigned d = t_iter > win_base ? (unsigned)(t_iter - win_base) : 0u;
                for (int k = 0; k < nent; ++k) {
                    const int ek = ent[k * MD + p];
                    const int slot = ek >> 1, cn = ecnt[k * MD + p], cn2 = cn + ((ek & 1) ? 1 : -1);
                    st.stw(F + slot, cn2);
                    if (d) {
                        unsigned int* h = R.cluster_hist + P.cluster_hist_ptr[c0 + slot_cluster(slot)];
                        if (cn > 0) atomicAdd(&h[cn], d);
                        if (cn2 > 0) atomicAdd(&h[cn2], 0u - d);
                    }
                }
                if (naive) {
                    st.stw(f, now);
                    if (d) {
                        if (prev == -1) atomicAdd(&R.frag_err_count[f0 + f], d); else atomicAdd(&R.aln_count[a0 + prev], d);
                        if (now == -1) atomicAdd(&R.frag_err_count[f0 + f], 0u - d); else atomicAdd(&R.aln_count[a0 + now], 0u - d);
                    }
                } else {
                    for (int q = 0; q < nmoves; ++q) {
                        const int g = mv[q * MD + p] >> 1;
                        const bool from_err = mv[q * MD + p] & 1;
                        st.stw(g, from_err ? 0 : -1);
                        if (d) {
                            const int ga0 = P.frag_aln_ptr[f0 + g];
                            if (from_err) { atomicAdd(&R.frag_err_count[f0 + g], d); atomicAdd(&R.aln_count[ga0], 0u - d); }
                            else { atomicAdd(&R.aln_count[ga0], d); atomicAdd(&R.frag_err_count[f0 + g], 0u - d); }
                        }
                    }
                }
            }
        }
        // base of the next window: RNG position, iteration count, totals and cached logBinomial (= what the first uncommitted
        // offset would see), written into every CTA of the cluster by the one thread that knows them
        bool writer = false;
        unsigned long long nb_s = 0;
        int nb_iter = 0, nb_err = 0;
        if (act && bad && goff == cut) {
            // the reference throws here: the chain stops with the proposal counted and its draws consumed
            ++n_naive_prop; n_reassign += 1;
            writer = true; nb_s = rng.s; nb_iter = t_iter; nb_err = -5;   // MBSV_ERR_INVARIANT
        } else if (cut == GIANT_NOCUT) {
            if (crank == CS - 1 && tid == NT - 1) {
                writer = true;
                nb_s = sc.jumpA * sb + sc.jumpC;
                for (int i = 0; i < sc.fin[2]; ++i) nb_s = nb_s * JavaRandom::MULT2 + JavaRandom::ADD2;
                nb_iter = iter_base + sc.fin[0];
#pragma unroll
                for (int x = 0; x < MAXE; ++x) {                          // inclusive: this thread's own proposal counts
                    if (flagA && ((touched >> x) & 1u)) lastx[x] = -2;    // (its logBinomial is in nl[x])
                    Ecur[x] += flagA ? dE[x] : 0; Lcur[x] += flagA ? dL[x] : 0;
                }
            }
        } else if (goff == cut && cut != errlane) {
            writer = true; nb_s = jA * sb + jC; nb_iter = t_iter;
        }
        if (writer) {
            double lbn[MAXE];
#pragma unroll
            for (int x = 0; x < MAXE; ++x) lbn[x] = lastx[x] == -2 ? nl[x] : lastx[x] >= 0 ? published(x, lastx[x]) : sc.LB[x];
#pragma unroll
            for (int k = 0; k < CS; ++k) {
                Scalars* o = remote(&sc, k);
                o->s_base = nb_s;
                if (nb_err) { o->err = nb_err; o->done = 1; }
                else {
                    o->iter_base = nb_iter;
                    if (nb_iter >= N) o->done = 1;
#pragma unroll
                    for (int x = 0; x < MAXE; ++x) { o->Etot[x] = Ecur[x]; o->Ltot[x] = Lcur[x]; o->LB[x] = lbn[x]; }
                }
            }
        }
        csync();
        GPROF(7);
#ifdef MBSV_GIANT_PROF
        ++nwindows; if (cut != GIANT_NOCUT) ++ncuts;
#endif
    }
#ifdef MBSV_GIANT_PROF
    if (tid == 0 && blockIdx.x == 0)
        printf("giant prof (CS %d): windows %lld cuts %lld iters %d | cycles/window: P1 %lld reach %lld gather+A %lld scan %lld nl %lld evalB+ins %lld stale %lld commit %lld\n",
               CS, nwindows, ncuts, N, prof[0] / max(nwindows, 1LL), prof[1] / max(nwindows, 1LL), prof[2] / max(nwindows, 1LL), prof[3] / max(nwindows, 1LL),
               prof[4] / max(nwindows, 1LL), prof[5] / max(nwindows, 1LL), prof[6] / max(nwindows, 1LL), prof[7] / max(nwindows, 1LL));
#endif

    SPROF();
    // ---------------- epilogue: final tallies, state, counters (pooled in the first CTA of the cluster) ----------------
    {
        Scalars* s0p = remote(&sc, 0);
        const long long v[5] = {n_naive_prop, n_naive_acc, n_sv_prop, n_sv_acc, n_reassign};
#pragma unroll
        for (int i = 0; i < 5; ++i) if (v[i]) atomicAdd(&s0p->cnt[i], (unsigned long long)v[i]);
        if (nerr_delta) atomicAdd(&s0p->nerr, nerr_delta);
    }
    csync();
    const int err = sc.err;
    if (err == 0) {
        const long long nwin = (long long)N - win_base;
        if (nwin > 0) {
            const unsigned d = (unsigned)nwin;
            // eight loads in flight per thread: these passes over the whole state are latency bound
            constexpr int EB = 8;
            for (int fb = goff; fb < F; fb += CS * NT * EB) {
                int a[EB], ap[EB];
#pragma unroll
                for (int k = 0; k < EB; ++k) { const int ff = fb + k * CS * NT; a[k] = ff < F ? st.ld(ff) : -2; }
#pragma unroll
                for (int k = 0; k < EB; ++k) { const int ff = fb + k * CS * NT; ap[k] = a[k] >= 0 ? __ldg(P.frag_aln_ptr + f0 + ff) : 0; }
#pragma unroll
                for (int k = 0; k < EB; ++k) {
                    const int ff = fb + k * CS * NT;
                    if (a[k] == -1) atomicAdd(&R.frag_err_count[f0 + ff], d);
                    else if (a[k] >= 0) atomicAdd(&R.aln_count[ap[k] + a[k]], d);
                }
            }
            const int nslot = C * E;
            for (int sb = goff; sb < nslot; sb += CS * NT * EB) {
                int n[EB], hp[EB];
#pragma unroll
                for (int k = 0; k < EB; ++k) { const int s = sb + k * CS * NT; n[k] = s < nslot ? st.ld(F + s) : 0; }
#pragma unroll
                for (int k = 0; k < EB; ++k) { const int s = sb + k * CS * NT; hp[k] = n[k] > 0 ? __ldg(P.cluster_hist_ptr + c0 + slot_cluster(s)) : 0; }
#pragma unroll
                for (int k = 0; k < EB; ++k) if (n[k] > 0) atomicAdd(&R.cluster_hist[hp[k] + n[k]], d);
            }
        }
    }
    if (R.final_assign)
        for (int ff = goff; ff < F; ff += CS * NT) R.final_assign[(size_t)chain * P.n_frag + f0 + ff] = st.ld(ff);
    if (crank == 0) {
        if (R.final_logprob) {
            const double lp = full_logprob();
            if (tid == 0) R.final_logprob[scx] = lp;
        }
        if (tid == 0) {
            if (R.counters) {
                long long* k = R.counters + scx * 5;
                for (int i = 0; i < 5; ++i) k[i] = (long long)sc.cnt[i];
            }
            if (R.rng_state) R.rng_state[scx] = sc.s_base & JavaRandom::MASK;
            if (R.error) R.error[scx] = err;
            if (R.totals) {
                for (int i = 0; i < 5; ++i) atomicAdd(&R.totals[i], sc.cnt[i]);
                if (err) atomicCAS(&R.totals[5], 0ULL, (unsigned long long)(long long)err);
            }
        }
    }
    csync();       // no CTA of the cluster leaves while another may still address its shared memory
#ifdef MBSV_GIANT_PROF
    SPROF();
    if (tid == 0 && blockIdx.x == 0)
        printf("giant prof setup (cycles): zero+stage %lld init %lld counts %lld iterations %lld epilogue %lld\n", sp[1] - sp[0], sp[2] - sp[1],
               sp[3] - sp[2], sp[4] - sp[3], sp[5] - sp[4]);
#endif
}

}  // namespace mbsv
