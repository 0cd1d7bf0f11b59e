// ORACLE — TEST INFRASTRUCTURE ONLY (see jsemantics.hpp header).
//
// CPU restatement of the reference's MCMC hot path on flat index arrays:
//   MultiBreakSV.java:773-877    runMCMC           -> run_chain()
//   MultiBreakSV.java:990-1099   initializeMatrix  -> Chain::initialize()
//   MultiBreakSV.java:1101-1182  iterate           -> Chain::iterate_*()
//   MultiBreakSV.java:1184-1289  naiveMove         -> Chain::propose_naive()
//   MultiBreakSV.java:1291-1353  svMove            -> Chain::propose_sv()
//   MultiBreakSV.java:1601-1641  getChangedClusters
//   MultiBreakSV.java:879-988    incrementAllCounters (only the *long tallies reach the output files)
//   MultiBreakSV.java:1644-1695  logPoisson / logBinomial / logbinomialcoefficient / logNormal
//   MultiBreakSVAssignmentMatrix.java:130-246 computeLogProb, :264-414 setBreakpoints
//
// Two variants:
//   FAITHFUL     — every proposal recomputes the whole log-posterior in the reference's summation order
//                  (supports.keySet() order, experiments.keySet() order) and tallies literally every
//                  iteration.  This is the parity reference for the REPLAY_EXACT kernel.
//   INCREMENTAL  — same RNG stream, same proposals, Delta-log-posterior and O(1)
//                  difference tallies.  Defines, operation for operation, the arithmetic of the FAST
//                  kernel; also the honest CPU baseline.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <vector>

#include "jsemantics.hpp"

namespace orc {

// One independent problem inside a (possibly batched) set of flat arrays.  Fragment / cluster indices in
// the arrays are global; [f0,f0+F) and [c0,c0+C) are this problem's ranges.
struct View {
    int F = 0, C = 0, E = 0, f0 = 0, c0 = 0;
    const int32_t* frag_aln_ptr = nullptr;     // [Ftot+1]
    const int32_t* aln_numerrors = nullptr;    // [A]
    const int32_t* aln_len = nullptr;          // [A]
    const int32_t* aln_exp = nullptr;          // [A]
    const int32_t* aln_esp_ptr = nullptr;      // [A+1]
    const int32_t* aln_esp_cluster = nullptr;  // [K] global cluster index or -1
    const int32_t* aln_esp_leaf = nullptr;     // [K] leaf slot local to the cluster (conn-comps), may be null
    const uint8_t* cluster_kind = nullptr;     // [Ctot]
    const int32_t* supports_order = nullptr;   // [Ctot] global cluster idx, this problem's slice at c0
    const int32_t* cluster_leaf_ptr = nullptr; // [Ctot+1]
    const int32_t* leaf_owner = nullptr;       // per leaf slot, global fragment index or -1
    const int32_t* cluster_root_ptr = nullptr; // [Ctot+1]
    const int32_t* root_leaf_ptr = nullptr;    // [R+1]
    const int32_t* root_leaf = nullptr;        // leaf slots local to the cluster
    const int32_t* cluster_hist_ptr = nullptr; // [Ctot+1] offsets into the histogram output
    int nsv = 0;
    const int32_t* sv_cluster = nullptr;       // [nsv] global cluster index (this problem's slice)
    const int32_t* sv_frag_ptr = nullptr;      // [nsv+1] (slice; absolute offsets into sv_frag)
    const int32_t* sv_frag = nullptr;          // global fragment indices
    const int32_t* sv_numchanged = nullptr;    // [nsv]
    const double* exp_pseq = nullptr;          // [E]
    const double* logprobcov = nullptr;        // [E][max_support+1]
    int max_support = 0;
};

struct RunParams {
    int mode = 0;                 // 0 FAITHFUL, 1 INCREMENTAL
    int32_t num_iters = 0;
    int32_t burnin = 0;           // reference: numiters/10, 0 -> 10 (MultiBreakSV.java:186-193)
    double perr = 0.001;
    int64_t seed = 123456;
    int java_int_wrap = 1;        // totals wrap like Java int (MultiBreakSVAssignmentMatrix.java:38-39,218-219)
    const int32_t* init_assign = nullptr;   // [F] local (matchfile path, :1030-1085) or null
    int32_t memo_states = 0;      // INCREMENTAL: problems without connected components whose joint state space
                                  // prod(n_f + 1) is at most this use the full log-posterior of the proposed state
                                  // (the reference's own memo of state log-probs, MultiBreakSV.java:1133-1137)
};

// Optional per-iteration trace (mcmc.txt columns + accepted-move log).
struct Trace {
    double* logprob = nullptr;    // [N] log-posterior of the state after the iteration
    int32_t* move_frag = nullptr; // [N] -1 none; >=0 local fragment (Naive accepted); -(2+k) AlterSV of sv entry k
    int32_t* move_assign = nullptr; // [N] new assignment of that fragment (Naive)
};

struct ChainResult {
    double final_logprob = 0.0;
    int64_t naive_proposed = 0, naive_accepted = 0, sv_proposed = 0, sv_accepted = 0;
    int64_t reassignments = 0;    // Naive: 1, AlterSV: numchanged (MultiBreakSV.java:1121,1311); lazy: 0
    uint64_t rng_state = 0;
    int64_t rng_steps = 0;
    int error = 0;                // -5: invariant violated (Error thrown by the reference)
};

// MultiBreakSV.java:1661-1695.  k,n as 64-bit so that the non-wrapping mode can use them unchanged.
static inline double log_binomial_coefficient(int64_t n, int64_t r) {
    double t = 0;
    int64_t m = n - r;
    if (r < m) r = m;
    for (int64_t i = n, j = 1; i > r; --i, ++j) t = t + j_log((double)i) - j_log((double)j);
    return t;
}
static inline double log_normal(double x, double mu, double sigmasq, double log2_plus_logpi) {
    double d = x - mu;
    return -(d * d) / (2 * sigmasq) - (log2_plus_logpi + j_log(sigmasq)) / 2;   // Math.pow(d,2) == d*d
}
struct BinomialConsts { double logp, log1mp, log2pi; };
static inline double log_binomial(int64_t k, int64_t n, double p, const BinomialConsts& bc) {
    if ((double)n * p <= 5 || (double)n * (1 - p) <= 5)
        return log_binomial_coefficient(n, k) + (double)k * bc.logp + (double)(n - k) * bc.log1mp;
    return log_normal((double)k, (double)n * p, (double)n * p * (1 - p), bc.log2pi);
}

struct Chain {
    const View& v;
    const RunParams& rp;
    JRandom rand;
    std::vector<int32_t> assign;                 // [F], -1 = error
    std::vector<int32_t> cnt;                    // simple clusters: [C][E] selected-ESP counts
    std::vector<std::vector<int32_t>> ccsup;     // conn-comps: [C*E] sorted set-cover multisets
    std::vector<int64_t> Etot, Ltot;             // [E]
    std::vector<double> LB;                      // [E] cached logBinomial (INCREMENTAL)
    std::vector<BinomialConsts> bc;
    int64_t nerr = 0;
    double logprob = 0.0;
    double logperr, log1mperr, logF;
    std::vector<double> logint;                  // log(n) for n <= max alignments per fragment
    // outputs (difference tallies or literal tallies)
    uint32_t* aln_count;                         // [A_total] global alignment index
    uint32_t* frag_err_count;                    // [F_total] global fragment index
    uint32_t* cluster_hist;                      // compact, cluster_hist_ptr
    ChainResult res;
    int32_t win_base;                            // first iteration inside the long-burn-in window
    bool use_full = false;                       // evaluate proposals with full_logprob() (FAITHFUL, or memo-eligible)
    // memo-eligible problems (INCREMENTAL): log-posterior of every joint state, filled on first use and shared by
    // the chains of the problem — the reference's `logprobs` map (MultiBreakSV.java:58,1133-1137) as an array
    std::vector<double>* lpcache = nullptr;
    std::vector<int64_t> stride;
    int64_t sidx = 0;
    // scratch reused across iterations (no heap traffic in the hot loop)
    std::vector<int> s_dirty, s_dd;
    std::vector<uint8_t> s_touched, s_tt;
    std::vector<int64_t> s_E0, s_L0;
    std::vector<std::vector<int32_t>> s_cc_old;
    std::vector<int32_t> s_sv_prev;
    std::vector<double> s_newLB;

    Chain(const View& v_, const RunParams& rp_, int64_t seed, uint32_t* ac, uint32_t* fe, uint32_t* ch)
        : v(v_), rp(rp_), rand(seed), aln_count(ac), frag_err_count(fe), cluster_hist(ch) {
        assign.assign(v.F, -1);
        cnt.assign((size_t)v.C * v.E, 0);
        ccsup.resize((size_t)v.C * v.E);
        Etot.assign(v.E, 0); Ltot.assign(v.E, 0); LB.assign(v.E, 0.0);
        s_touched.assign(v.E, 0); s_tt.assign(v.E, 0);
        logperr = j_log(rp.perr);
        log1mperr = j_log(1 - rp.perr);
        logF = j_log((double)v.F);
        int maxn = 1;
        for (int f = 0; f < v.F; ++f) maxn = std::max(maxn, nalign(f));
        logint.assign(maxn + 1, 0.0);
        for (int n = 1; n <= maxn; ++n) logint[n] = j_log((double)n);
        double log2pi = j_log(2.0) + j_log(3.141592653589793);   // Math.log(2)+Math.log(Math.PI)
        for (int x = 0; x < v.E; ++x) bc.push_back({j_log(v.exp_pseq[x]), j_log(1 - v.exp_pseq[x]), log2pi});
        // window: iter > numiters - burnin  (MultiBreakSV.java:903,926,962)
        int64_t b = (int64_t)rp.num_iters - rp.burnin + 1;
        win_base = (int32_t)std::max<int64_t>(b, 0);
        use_full = rp.mode == 0;
        if (rp.mode == 1 && rp.memo_states > 0) {
            bool cc = false;
            for (int cl = 0; cl < v.C; ++cl) cc = cc || !simple(cl);
            int64_t nst = 1;
            for (int f = 0; f < v.F && nst <= rp.memo_states; ++f) nst *= (int64_t)nalign(f) + 1;
            use_full = !cc && nst <= rp.memo_states;
            if (use_full) {
                int64_t sp = 1;
                for (int f = 0; f < v.F; ++f) { stride.push_back(sp); sp *= (int64_t)nalign(f) + 1; }
            }
        }
    }

    inline int aln0(int f) const { return v.frag_aln_ptr[v.f0 + f]; }
    inline int nalign(int f) const { return v.frag_aln_ptr[v.f0 + f + 1] - v.frag_aln_ptr[v.f0 + f]; }
    inline const double* T(int x) const { return v.logprobcov + (size_t)x * (v.max_support + 1); }
    inline bool simple(int cl) const { return v.cluster_kind[v.c0 + cl] == 0; }
    inline int64_t wrap(int64_t t) const { return rp.java_int_wrap ? (int64_t)(int32_t)(uint32_t)(uint64_t)t : t; }

    // ---- support bookkeeping ------------------------------------------------------------------
    // Greedy set cover of MultiBreakSVAssignmentMatrix.java:337-409 for conn-comp cluster cl.
    void recompute_conncomp(int cl) {
        int cg = v.c0 + cl;
        int l0 = v.cluster_leaf_ptr[cg], nl = v.cluster_leaf_ptr[cg + 1] - l0;
        std::vector<int8_t> selexp(nl, -1);
        int numselected = 0;
        for (int l = 0; l < nl; ++l) {
            int fg = v.leaf_owner[l0 + l];
            if (fg < 0) continue;
            int a = assign[fg - v.f0];
            if (a == -1) continue;
            int al = v.frag_aln_ptr[fg] + a;
            for (int j = v.aln_esp_ptr[al]; j < v.aln_esp_ptr[al + 1]; ++j)
                if (v.aln_esp_cluster[j] == cg && v.aln_esp_leaf[j] == l) { selexp[l] = (int8_t)v.aln_exp[al]; ++numselected; }
        }
        for (int x = 0; x < v.E; ++x) ccsup[(size_t)cl * v.E + x].clear();
        if (numselected == 0) return;
        for (int x = 0; x < v.E; ++x) {
            std::vector<int32_t>& sup = ccsup[(size_t)cl * v.E + x];
            std::vector<char> found(nl, 0);
            std::vector<int> maximal;
            for (int r = v.cluster_root_ptr[cg]; r < v.cluster_root_ptr[cg + 1]; ++r) maximal.push_back(r);
            for (;;) {
                int maxcount = 0, maxind = -1;
                for (size_t i = 0; i < maximal.size(); ++i) {
                    int cur = 0;
                    for (int j = v.root_leaf_ptr[maximal[i]]; j < v.root_leaf_ptr[maximal[i] + 1]; ++j) {
                        int l = v.root_leaf[j];
                        if (selexp[l] == x && !found[l]) ++cur;
                    }
                    if (cur > maxcount) { maxcount = cur; maxind = (int)i; }
                }
                if (maxind == -1) {
                    for (int l = 0; l < nl; ++l) if (selexp[l] == x && !found[l]) { res.error = -5; }
                    break;
                }
                for (int j = v.root_leaf_ptr[maximal[maxind]]; j < v.root_leaf_ptr[maximal[maxind] + 1]; ++j)
                    if (selexp[v.root_leaf[j]] == x) found[v.root_leaf[j]] = 1;
                sup.push_back(maxcount);
                maximal.erase(maximal.begin() + maxind);
                bool all = true;
                for (int l = 0; l < nl; ++l) if (selexp[l] == x && !found[l]) all = false;
                if (all) break;
            }
            std::sort(sup.begin(), sup.end());
        }
    }

    // log-probability-of-coverage of one cluster in Java order (AM.java:147-169)
    double cluster_lpc(int cl) const {
        double lpc = 0.0;
        for (int x = 0; x < v.E; ++x) {
            if (simple(cl)) {
                int n = cnt[(size_t)cl * v.E + x];
                if (n > 0) lpc = lpc + T(x)[n];
            } else {
                for (int s : ccsup[(size_t)cl * v.E + x]) lpc = lpc + T(x)[s];
            }
        }
        return lpc;
    }

    // MultiBreakSVAssignmentMatrix.java:130-246 with the totals kept incrementally (integers: exact).
    double full_logprob() const {
        double lp = 0;
        for (int i = 0; i < v.C; ++i) {
            int cl = v.supports_order[v.c0 + i] - v.c0;
            lp += cluster_lpc(cl);
        }
        lp += 0.0;                                   // prior[numbreakpoints] == 0 (MultiBreakSV.java:628)
        lp += (double)nerr * logperr;
        for (int x = 0; x < v.E; ++x) lp += log_binomial(wrap(Etot[x]), wrap(Ltot[x]), v.exp_pseq[x], bc[x]);
        return lp;
    }

    void rebuild_from_assign() {
        std::fill(cnt.begin(), cnt.end(), 0);
        std::fill(Etot.begin(), Etot.end(), 0);
        std::fill(Ltot.begin(), Ltot.end(), 0);
        nerr = 0;
        for (int f = 0; f < v.F; ++f) {
            if (assign[f] == -1) { ++nerr; continue; }
            int al = aln0(f) + assign[f];
            int x = v.aln_exp[al];
            Etot[x] += v.aln_numerrors[al];
            Ltot[x] += v.aln_len[al];
            for (int j = v.aln_esp_ptr[al]; j < v.aln_esp_ptr[al + 1]; ++j) {
                int cg = v.aln_esp_cluster[j];
                if (cg >= 0 && simple(cg - v.c0)) ++cnt[(size_t)(cg - v.c0) * v.E + x];
            }
        }
        for (int cl = 0; cl < v.C; ++cl) if (!simple(cl)) recompute_conncomp(cl);
        for (int x = 0; x < v.E; ++x) LB[x] = log_binomial(wrap(Etot[x]), wrap(Ltot[x]), v.exp_pseq[x], bc[x]);
    }

    // MultiBreakSV.java:990-1099
    void initialize() {
        if (rp.init_assign) {
            for (int f = 0; f < v.F; ++f) assign[f] = rp.init_assign[f];
        } else {
            for (int f = 0; f < v.F; ++f) {
                if (rand.nextDouble() < rp.perr) assign[f] = -1;
                else assign[f] = rand.nextInt(nalign(f));
            }
        }
        rebuild_from_assign();
        logprob = full_logprob();
        if (!stride.empty()) { sidx = 0; for (int f = 0; f < v.F; ++f) sidx += (assign[f] + 1) * stride[f]; }
    }

    // ---- proposals ------------------------------------------------------------------------------
    struct Proposal {
        int kind = 0;            // 0 Naive, 1 AlterSV
        int f = -1, prev = -1, now = -1;   // Naive
        int sv = -1;             // AlterSV
        double Alogq = 0, Anewlogq = 0;
    };

    // MultiBreakSV.java:1184-1273 (choice of fragment and new assignment, proposal log-probabilities)
    Proposal propose_naive() {
        Proposal p;
        p.kind = 0;
        p.f = (int)(rand.nextDouble() * v.F);
        int n = nalign(p.f);
        p.prev = assign[p.f];
        double r = rand.nextDouble();
        if (p.prev == -1) {
            int thisassign = 0;
            double sum = (double)1 / n;
            while (thisassign < n - 1 && r > sum) { ++thisassign; sum += (double)1 / n; }
            p.now = thisassign;
            if (n == 1) { p.Alogq = -logF + 0.0; p.Anewlogq = -logF + 0.0; }
            else { p.Alogq = -logF + logperr; p.Anewlogq = -logF - logint[n]; }
        } else if (r < rp.perr || n == 1) {
            p.now = -1;
            if (n == 1) { p.Alogq = -logF + 0.0; p.Anewlogq = -logF + 0.0; }
            else { p.Alogq = -logF - logint[n]; p.Anewlogq = -logF + logperr; }
        } else {
            int thisassign = 0;
            r = rand.nextDouble();
            double sum = (double)1 / (n - 1);
            while (thisassign < n - 1 && (thisassign == p.prev || r > sum)) {
                ++thisassign;
                if (thisassign != p.prev) sum += (double)1 / (n - 1);
            }
            p.now = thisassign;
            p.Anewlogq = -logF + log1mperr - logint[n - 1];
            p.Alogq = -logF + log1mperr - logint[n - 1];
        }
        return p;
    }

    // MultiBreakSV.java:1291-1343
    Proposal propose_sv() {
        Proposal p;
        p.kind = 1;
        p.sv = (int)(rand.nextDouble() * v.nsv);
        double l = j_log((double)v.nsv);
        p.Alogq = -l;
        p.Anewlogq = -l;
        return p;
    }

    // Apply / revert the state change of a proposal.  dir=+1 apply, -1 revert.  Returns (via dcov) the
    // coverage log-probability delta in the INCREMENTAL association; touched conn-comps are recomputed.
    void move_fragment(int f, int from, int to, double& dcov, double& derr, std::vector<int>& dirty_cc,
                       uint8_t* exp_touched) {
        if (from != -1) {
            int al = aln0(f) + from, x = v.aln_exp[al];
            for (int j = v.aln_esp_ptr[al]; j < v.aln_esp_ptr[al + 1]; ++j) {
                int cg = v.aln_esp_cluster[j];
                if (cg < 0) continue;
                int cl = cg - v.c0;
                if (simple(cl)) {
                    int32_t& n = cnt[(size_t)cl * v.E + x];
                    dcov = dcov + (T(x)[n - 1] - T(x)[n]);
                    n = n - 1;
                } else if (std::find(dirty_cc.begin(), dirty_cc.end(), cl) == dirty_cc.end()) dirty_cc.push_back(cl);
            }
            Etot[x] -= v.aln_numerrors[al];
            Ltot[x] -= v.aln_len[al];
            exp_touched[x] = 1;
        } else {
            derr = derr - logperr;
            --nerr;
        }
        if (to != -1) {
            int al = aln0(f) + to, x = v.aln_exp[al];
            for (int j = v.aln_esp_ptr[al]; j < v.aln_esp_ptr[al + 1]; ++j) {
                int cg = v.aln_esp_cluster[j];
                if (cg < 0) continue;
                int cl = cg - v.c0;
                if (simple(cl)) {
                    int32_t& n = cnt[(size_t)cl * v.E + x];
                    dcov = dcov + (T(x)[n + 1] - T(x)[n]);
                    n = n + 1;
                } else if (std::find(dirty_cc.begin(), dirty_cc.end(), cl) == dirty_cc.end()) dirty_cc.push_back(cl);
            }
            Etot[x] += v.aln_numerrors[al];
            Ltot[x] += v.aln_len[al];
            exp_touched[x] = 1;
        } else {
            derr = derr + logperr;
            ++nerr;
        }
        assign[f] = to;
    }

    // Tally bookkeeping of one fragment / cluster transition with difference tallies: an accepted move
    // at iteration t makes the old value "leave" and the new value "enter" at u = max(t, win_base).
    inline void tally_fragment_change(int f, int from, int to, int32_t t) {
        if (t <= win_base) return;
        uint32_t d = (uint32_t)(t - win_base);
        if (from == -1) frag_err_count[v.f0 + f] += d; else aln_count[aln0(f) + from] += d;
        if (to == -1) frag_err_count[v.f0 + f] -= d; else aln_count[aln0(f) + to] -= d;
    }
    inline void tally_support_change(int cl, int32_t from, int32_t to, int32_t t) {
        if (t <= win_base) return;
        uint32_t d = (uint32_t)(t - win_base);
        uint32_t* h = cluster_hist + v.cluster_hist_ptr[v.c0 + cl];
        if (from > 0) h[from] += d;
        if (to > 0) h[to] -= d;
    }
    void tally_finalize() {
        int64_t n = (int64_t)rp.num_iters - win_base;
        if (n <= 0) return;
        uint32_t d = (uint32_t)n;
        for (int f = 0; f < v.F; ++f) {
            if (assign[f] == -1) frag_err_count[v.f0 + f] += d; else aln_count[aln0(f) + assign[f]] += d;
        }
        for (int cl = 0; cl < v.C; ++cl) {
            uint32_t* h = cluster_hist + v.cluster_hist_ptr[v.c0 + cl];
            for (int x = 0; x < v.E; ++x) {
                if (simple(cl)) { int n2 = cnt[(size_t)cl * v.E + x]; if (n2 > 0) h[n2] += d; }
                else for (int s : ccsup[(size_t)cl * v.E + x]) h[s] += d;
            }
        }
    }
    // Literal incrementAllCounters for the *long tallies (MultiBreakSV.java:892-936, 962-969).
    void tally_literal(int32_t iter) {
        if (!(iter > rp.num_iters - rp.burnin)) return;
        for (int cl = 0; cl < v.C; ++cl) {
            uint32_t* h = cluster_hist + v.cluster_hist_ptr[v.c0 + cl];
            for (int x = 0; x < v.E; ++x) {
                if (simple(cl)) { int n = cnt[(size_t)cl * v.E + x]; if (n > 0) ++h[n]; }
                else for (int s : ccsup[(size_t)cl * v.E + x]) ++h[s];
            }
        }
        for (int f = 0; f < v.F; ++f) {
            if (assign[f] == -1) ++frag_err_count[v.f0 + f]; else ++aln_count[aln0(f) + assign[f]];
        }
    }

    // One iteration (MultiBreakSV.java:1101-1182).  Returns true when the state changed.
    bool iterate(int32_t iter, Trace* tr) {
        if (tr && tr->move_frag) { tr->move_frag[iter] = -1; tr->move_assign[iter] = -1; }
        double r = rand.nextDouble();
        if (r < 0.5) return false;                       // lazy chain
        r = rand.nextDouble();
        Proposal p;
        if (r < 0.9 || v.nsv == 0) { p = propose_naive(); ++res.naive_proposed; ++res.reassignments; }
        else { p = propose_sv(); ++res.sv_proposed; res.reassignments += v.sv_numchanged[p.sv]; }

        // ---- apply tentatively ----
        double dcov = 0.0, derr = 0.0;
        std::vector<int>& dirty = s_dirty; dirty.clear();
        std::vector<uint8_t>& touched = s_touched; std::fill(touched.begin(), touched.end(), 0);
        s_E0 = Etot; s_L0 = Ltot;
        std::vector<std::vector<int32_t>>& cc_old = s_cc_old; cc_old.clear();
        std::vector<int32_t>& sv_prev = s_sv_prev; sv_prev.clear();
        if (p.kind == 0) {
            if (p.now == p.prev) { res.error = -5; return false; }   // Error at :1122-1123
            move_fragment(p.f, p.prev, p.now, dcov, derr, dirty, touched.data());
        } else {
            for (int q = v.sv_frag_ptr[p.sv]; q < v.sv_frag_ptr[p.sv + 1]; ++q) {
                int f = v.sv_frag[q] - v.f0;
                int prev = assign[f];
                sv_prev.push_back(prev);
                move_fragment(f, prev, prev == -1 ? 0 : -1, dcov, derr, dirty, touched.data());
            }
        }
        for (int cl : dirty) {
            double before = cluster_lpc(cl);
            for (int x = 0; x < v.E; ++x) cc_old.push_back(ccsup[(size_t)cl * v.E + x]);
            recompute_conncomp(cl);
            dcov = dcov + (cluster_lpc(cl) - before);
        }
        // ---- acceptance ratio ----
        double newlogprob, arg;
        std::vector<double>& newLB = s_newLB; newLB = LB;
        int64_t newidx = sidx;
        if (use_full) {
            if (lpcache && !stride.empty()) {
                if (p.kind == 0) newidx += (int64_t)(p.now - p.prev) * stride[p.f];
                else for (int q = v.sv_frag_ptr[p.sv], i = 0; q < v.sv_frag_ptr[p.sv + 1]; ++q, ++i)
                    newidx += (sv_prev[i] == -1 ? 1 : -1) * stride[v.sv_frag[q] - v.f0];
                double& c = (*lpcache)[newidx];
                if (c != c) c = full_logprob();      // NaN = not computed yet
                newlogprob = c;
            } else newlogprob = full_logprob();
            arg = (newlogprob + p.Alogq) - (logprob + p.Anewlogq);               // :1146
        } else {
            double dlb = 0.0;
            for (int x = 0; x < v.E; ++x) if (touched[x]) {
                newLB[x] = log_binomial(wrap(Etot[x]), wrap(Ltot[x]), v.exp_pseq[x], bc[x]);
                dlb = dlb + (newLB[x] - LB[x]);
            }
            double delta = (dcov + derr) + dlb;
            newlogprob = logprob + delta;
            arg = (delta + p.Alogq) - p.Anewlogq;
        }
        double e = j_exp(arg);
        double ratio = (e != e) ? e : (1.0 < e ? 1.0 : e);                       // Math.min(1, e)
        r = rand.nextDouble();
        if (ratio > r) {
            if (p.kind == 0) ++res.naive_accepted; else ++res.sv_accepted;
            logprob = newlogprob;
            sidx = newidx;
            LB.swap(newLB);
            if (rp.mode == 1) {
                // difference tallies
                if (p.kind == 0) tally_fragment_change(p.f, p.prev, p.now, iter);
                else for (int q = v.sv_frag_ptr[p.sv], i = 0; q < v.sv_frag_ptr[p.sv + 1]; ++q, ++i)
                    tally_fragment_change(v.sv_frag[q] - v.f0, sv_prev[i], sv_prev[i] == -1 ? 0 : -1, iter);
                if (iter > win_base) retally_supports(p, sv_prev, dirty, cc_old, iter);
            }
            if (tr && tr->move_frag) {
                if (p.kind == 0) { tr->move_frag[iter] = p.f; tr->move_assign[iter] = p.now; }
                else tr->move_frag[iter] = -(2 + p.sv);
            }
            return true;
        }
        // ---- reject: revert ----
        double d1 = 0, d2 = 0;
        s_dd.clear();
        if (p.kind == 0) move_fragment(p.f, p.now, p.prev, d1, d2, s_dd, s_tt.data());
        else for (int q = v.sv_frag_ptr[p.sv + 1] - 1, i = (int)sv_prev.size() - 1; q >= v.sv_frag_ptr[p.sv]; --q, --i)
            move_fragment(v.sv_frag[q] - v.f0, assign[v.sv_frag[q] - v.f0], sv_prev[i], d1, d2, s_dd, s_tt.data());
        size_t k = 0;
        for (int cl : dirty) for (int x = 0; x < v.E; ++x) ccsup[(size_t)cl * v.E + x] = cc_old[k++];
        Etot = s_E0; Ltot = s_L0;
        return false;
    }

    // Support-histogram difference tallies for an accepted move inside the window: replay the count
    // transitions of the move one unit step at a time (intermediate values telescope away).
    void retally_supports(const Proposal& p, const std::vector<int32_t>& sv_prev, const std::vector<int>& dirty,
                          const std::vector<std::vector<int32_t>>& cc_old, int32_t iter) {
        auto steps = [&](int f, int from, int to) {
            // counts already hold the post-move values: walk the move backwards
            if (to != -1) {
                int al = aln0(f) + to, x = v.aln_exp[al];
                for (int j = v.aln_esp_ptr[al + 1] - 1; j >= v.aln_esp_ptr[al]; --j) {
                    int cg = v.aln_esp_cluster[j];
                    if (cg < 0 || !simple(cg - v.c0)) continue;
                    int32_t& n = cnt[(size_t)(cg - v.c0) * v.E + x];
                    tally_support_change(cg - v.c0, n - 1, n, iter); n = n - 1;
                }
            }
            if (from != -1) {
                int al = aln0(f) + from, x = v.aln_exp[al];
                for (int j = v.aln_esp_ptr[al + 1] - 1; j >= v.aln_esp_ptr[al]; --j) {
                    int cg = v.aln_esp_cluster[j];
                    if (cg < 0 || !simple(cg - v.c0)) continue;
                    int32_t& n = cnt[(size_t)(cg - v.c0) * v.E + x];
                    tally_support_change(cg - v.c0, n + 1, n, iter); n = n + 1;
                }
            }
        };
        auto redo = [&](int f, int from, int to) {
            if (from != -1) {
                int al = aln0(f) + from, x = v.aln_exp[al];
                for (int j = v.aln_esp_ptr[al]; j < v.aln_esp_ptr[al + 1]; ++j) {
                    int cg = v.aln_esp_cluster[j];
                    if (cg >= 0 && simple(cg - v.c0)) --cnt[(size_t)(cg - v.c0) * v.E + x];
                }
            }
            if (to != -1) {
                int al = aln0(f) + to, x = v.aln_exp[al];
                for (int j = v.aln_esp_ptr[al]; j < v.aln_esp_ptr[al + 1]; ++j) {
                    int cg = v.aln_esp_cluster[j];
                    if (cg >= 0 && simple(cg - v.c0)) ++cnt[(size_t)(cg - v.c0) * v.E + x];
                }
            }
        };
        if (p.kind == 0) { steps(p.f, p.prev, p.now); redo(p.f, p.prev, p.now); }
        else {
            for (int q = v.sv_frag_ptr[p.sv + 1] - 1, i = (int)sv_prev.size() - 1; q >= v.sv_frag_ptr[p.sv]; --q, --i)
                steps(v.sv_frag[q] - v.f0, sv_prev[i], sv_prev[i] == -1 ? 0 : -1);
            for (int q = v.sv_frag_ptr[p.sv], i = 0; q < v.sv_frag_ptr[p.sv + 1]; ++q, ++i)
                redo(v.sv_frag[q] - v.f0, sv_prev[i], sv_prev[i] == -1 ? 0 : -1);
        }
        size_t k = 0;
        for (int cl : dirty) for (int x = 0; x < v.E; ++x) {
            for (int s : cc_old[k]) tally_support_change(cl, s, 0, iter);
            for (int s : ccsup[(size_t)cl * v.E + x]) tally_support_change(cl, 0, s, iter);
            ++k;
        }
    }
};

// MultiBreakSV.java:773-877.  Tallies are ACCUMULATED into aln_count / frag_err_count / cluster_hist
// (callers zero them; several chains of one problem pool into the same arrays).
static inline ChainResult run_chain(const View& v, const RunParams& rp, int64_t seed, uint32_t* aln_count,
                                    uint32_t* frag_err_count, uint32_t* cluster_hist, int32_t* final_assign,
                                    Trace* tr, int32_t* initial_assign = nullptr, double* initial_logprob = nullptr,
                                    std::vector<double>* lpcache = nullptr) {
    Chain ch(v, rp, seed, aln_count, frag_err_count, cluster_hist);
    if (lpcache && ch.use_full && !ch.stride.empty()) {
        if (lpcache->empty()) {
            int64_t nst = 1;
            for (int f = 0; f < v.F; ++f) nst *= (int64_t)ch.nalign(f) + 1;
            lpcache->assign((size_t)nst, std::numeric_limits<double>::quiet_NaN());
        }
        ch.lpcache = lpcache;
    }
    ch.initialize();
    if (initial_assign) for (int f = 0; f < v.F; ++f) initial_assign[f] = ch.assign[f];
    if (initial_logprob) *initial_logprob = ch.logprob;
    for (int32_t iter = 0; iter < rp.num_iters && ch.res.error == 0; ++iter) {
        ch.iterate(iter, tr);
        if (rp.mode == 0) ch.tally_literal(iter);
        if (tr && tr->logprob) tr->logprob[iter] = ch.logprob;
    }
    if (rp.mode == 1) ch.tally_finalize();
    ch.res.final_logprob = (rp.mode == 0) ? ch.logprob : ch.full_logprob();
    ch.res.rng_state = ch.rand.s;
    ch.res.rng_steps = ch.rand.ncalls;
    if (final_assign) for (int f = 0; f < v.F; ++f) final_assign[f] = ch.assign[f];
    return ch.res;
}

}  // namespace orc
