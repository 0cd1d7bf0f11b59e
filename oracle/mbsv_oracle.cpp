// ORACLE — TEST INFRASTRUCTURE ONLY.
// C entry points (for ctypes) of the CPU restatement of raphael-group/multibreak-sv's sampler path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
// library; the product (multibreak_sv_b200/) never does.  Parity status: the reference ships no golden
// outputs and no JVM exists in this image, so this oracle is pinned by (i) public java.util.Random
// constants, (ii) glibc log/exp within 1 ulp, (iii) exact enumeration of tiny posteriors, (iv) the live Perl
// partition script and (v) the derived known-answers of SURVEY.md §8c (an independent line-by-line
// emulation).  Versus a real JVM it is "parity unpinned" (see DESIGN.md).
#include <atomic>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <mutex>
#include <thread>
#include <vector>

#include "jsemantics.hpp"
#include "refmodel.hpp"
#include "sampler.hpp"

using namespace orc;

extern "C" {

// Field-for-field the same layout as mbsv_problem_desc (include/mbsv_b200.h); declared separately so that
// the oracle does not depend on product headers.
struct orc_desc {
    int32_t abi_version;
    int32_t n_sub, n_frag, n_aln, n_aln_esp, n_cluster, n_exp, max_support;
    int32_t n_sv, n_sv_frag, n_leaf, n_root, n_root_leaf, n_hist;
    const int32_t* sub_frag_ptr;
    const int32_t* sub_cluster_ptr;
    const int32_t* sub_sv_ptr;
    const int32_t* frag_aln_ptr;
    const int32_t* aln_numerrors;
    const int32_t* aln_len;
    const int32_t* aln_exp;
    const int32_t* aln_esp_ptr;
    const int32_t* aln_esp_cluster;
    const int32_t* aln_esp_leaf;
    const uint8_t* cluster_kind;
    const int32_t* supports_order;
    const int32_t* cluster_hist_ptr;
    const int32_t* cluster_leaf_ptr;
    const int32_t* leaf_owner;
    const int32_t* cluster_root_ptr;
    const int32_t* root_leaf_ptr;
    const int32_t* root_leaf;
    const int32_t* sv_cluster;
    const int32_t* sv_frag_ptr;
    const int32_t* sv_frag;
    const int32_t* sv_numchanged;
    const double* exp_pseq;
    const double* exp_lambda;
    const double* logprobcov;
};

struct orc_params {
    int32_t mode;          // 0 faithful, 1 incremental
    int32_t n_chains;
    int32_t num_iters;
    int32_t burnin;
    double perr;
    int64_t seed;          // chain c of every subproblem uses seed + c
    int32_t device;        // unused
    int32_t java_int_wrap;
    const int32_t* init_assign;   // [n_frag] or NULL
    int32_t trace;
    int32_t memo_states;
};

struct orc_results {
    uint32_t* aln_count;       // [n_aln]   pooled over chains
    uint32_t* frag_err_count;  // [n_frag]
    uint32_t* cluster_hist;    // [n_hist]
    int32_t* final_assign;     // [n_chains][n_frag]
    double* final_logprob;     // [n_sub][n_chains]
    int64_t* counters;         // [n_sub][n_chains][5]
    uint64_t* rng_state;       // [n_sub][n_chains]
    int32_t* error;            // [n_sub][n_chains]
    double* trace_logprob;     // [n_sub][n_chains][num_iters] or NULL
    int32_t* trace_move_frag;
    int32_t* trace_move_assign;
    int32_t* initial_assign;   // [n_chains][n_frag] state after initializeMatrix, or NULL
    double* initial_logprob;   // [n_sub][n_chains] or NULL
    int64_t* totals;           // [6] or NULL
    double kernel_ms;
};

static View make_view(const orc_desc* d, int s) {
    View v;
    v.f0 = d->sub_frag_ptr[s]; v.F = d->sub_frag_ptr[s + 1] - v.f0;
    v.c0 = d->sub_cluster_ptr[s]; v.C = d->sub_cluster_ptr[s + 1] - v.c0;
    v.E = d->n_exp;
    v.frag_aln_ptr = d->frag_aln_ptr; v.aln_numerrors = d->aln_numerrors; v.aln_len = d->aln_len;
    v.aln_exp = d->aln_exp; v.aln_esp_ptr = d->aln_esp_ptr; v.aln_esp_cluster = d->aln_esp_cluster;
    v.aln_esp_leaf = d->aln_esp_leaf; v.cluster_kind = d->cluster_kind; v.supports_order = d->supports_order;
    v.cluster_leaf_ptr = d->cluster_leaf_ptr; v.leaf_owner = d->leaf_owner;
    v.cluster_root_ptr = d->cluster_root_ptr; v.root_leaf_ptr = d->root_leaf_ptr; v.root_leaf = d->root_leaf;
    v.cluster_hist_ptr = d->cluster_hist_ptr;
    int sv0 = d->sub_sv_ptr ? d->sub_sv_ptr[s] : 0;
    v.nsv = d->sub_sv_ptr ? d->sub_sv_ptr[s + 1] - sv0 : 0;
    v.sv_cluster = d->sv_cluster ? d->sv_cluster + sv0 : nullptr;
    v.sv_frag_ptr = d->sv_frag_ptr ? d->sv_frag_ptr + sv0 : nullptr;
    v.sv_frag = d->sv_frag;
    v.sv_numchanged = d->sv_numchanged ? d->sv_numchanged + sv0 : nullptr;
    v.exp_pseq = d->exp_pseq; v.logprobcov = d->logprobcov; v.max_support = d->max_support;
    return v;
}

// Runs every (subproblem, chain).  Subproblems are spread over nthreads host threads; a batch with fewer subproblems than
// twice the threads (a giant component) spreads its (subproblem, chain) pairs instead, every thread pooling its tallies in
// private arrays that are added up at the end (the chains of one subproblem share their tally bins).
int orc_run(const orc_desc* d, const orc_params* p, orc_results* r, int nthreads) {
    if (!d || !p || !r || p->n_chains < 1 || p->num_iters < 0) return -1;
    std::memset(r->aln_count, 0, sizeof(uint32_t) * (size_t)d->n_aln);
    std::memset(r->frag_err_count, 0, sizeof(uint32_t) * (size_t)d->n_frag);
    std::memset(r->cluster_hist, 0, sizeof(uint32_t) * (size_t)d->n_hist);
    std::atomic<int> next{0};
    std::atomic<int> worst{0};
    std::atomic<long long> tot[5];
    for (auto& t : tot) t.store(0);
    const bool by_chain = nthreads > 1 && d->n_sub < 2 * nthreads && p->n_chains > 1;
    std::mutex merge;
    auto chain_worker = [&]() {
        std::vector<uint32_t> ac((size_t)d->n_aln, 0), fe((size_t)d->n_frag, 0), ch((size_t)d->n_hist, 0);
        const long long n_items = (long long)d->n_sub * p->n_chains;
        for (;;) {
            const long long it = next.fetch_add(1);
            if (it >= n_items) break;
            const int s = (int)(it / p->n_chains), c = (int)(it % p->n_chains);
            View v = make_view(d, s);
            RunParams rp;
            rp.mode = p->mode; rp.num_iters = p->num_iters; rp.burnin = p->burnin; rp.perr = p->perr;
            rp.java_int_wrap = p->java_int_wrap;
            rp.memo_states = p->memo_states;
            rp.init_assign = p->init_assign ? p->init_assign + v.f0 : nullptr;
            std::vector<double> lpcache;     // (a pure cache: not sharing it between the chains only costs time)
            size_t sc = (size_t)s * p->n_chains + c;
            Trace tr, *trp = nullptr;
            if (p->trace && r->trace_logprob) {
                tr.logprob = r->trace_logprob + sc * p->num_iters;
                tr.move_frag = r->trace_move_frag + sc * p->num_iters;
                tr.move_assign = r->trace_move_assign + sc * p->num_iters;
                trp = &tr;
            }
            int32_t* fa = r->final_assign ? r->final_assign + (size_t)c * d->n_frag + v.f0 : nullptr;
            int32_t* ia = r->initial_assign ? r->initial_assign + (size_t)c * d->n_frag + v.f0 : nullptr;
            double* il = r->initial_logprob ? r->initial_logprob + sc : nullptr;
            ChainResult cr = run_chain(v, rp, p->seed + c, ac.data(), fe.data(), ch.data(), fa, trp, ia, il, &lpcache);
            if (trp) for (int i = 0; i < p->num_iters; ++i) if (tr.move_frag[i] >= 0) tr.move_frag[i] += v.f0;
            if (r->final_logprob) r->final_logprob[sc] = cr.final_logprob;
            if (r->counters) {
                int64_t* k = r->counters + sc * 5;
                k[0] = cr.naive_proposed; k[1] = cr.naive_accepted; k[2] = cr.sv_proposed; k[3] = cr.sv_accepted;
                k[4] = cr.reassignments;
            }
            tot[0] += cr.naive_proposed; tot[1] += cr.naive_accepted; tot[2] += cr.sv_proposed; tot[3] += cr.sv_accepted;
            tot[4] += cr.reassignments;
            if (r->rng_state) r->rng_state[sc] = cr.rng_state;
            if (r->error) r->error[sc] = cr.error;
            if (cr.error) worst.store(cr.error);
        }
        std::lock_guard<std::mutex> lk(merge);
        for (size_t i = 0; i < ac.size(); ++i) r->aln_count[i] += ac[i];
        for (size_t i = 0; i < fe.size(); ++i) r->frag_err_count[i] += fe[i];
        for (size_t i = 0; i < ch.size(); ++i) r->cluster_hist[i] += ch[i];
    };
    auto worker = [&]() {
        if (by_chain) { chain_worker(); return; }
        for (;;) {
            int s = next.fetch_add(1);
            if (s >= d->n_sub) break;
            View v = make_view(d, s);
            RunParams rp;
            rp.mode = p->mode; rp.num_iters = p->num_iters; rp.burnin = p->burnin; rp.perr = p->perr;
            rp.java_int_wrap = p->java_int_wrap;
            rp.memo_states = p->memo_states;
            rp.init_assign = p->init_assign ? p->init_assign + v.f0 : nullptr;
            std::vector<double> lpcache;     // shared by the chains of this subproblem
            for (int c = 0; c < p->n_chains; ++c) {
                size_t sc = (size_t)s * p->n_chains + c;
                Trace tr, *trp = nullptr;
                if (p->trace && r->trace_logprob) {
                    tr.logprob = r->trace_logprob + sc * p->num_iters;
                    tr.move_frag = r->trace_move_frag + sc * p->num_iters;
                    tr.move_assign = r->trace_move_assign + sc * p->num_iters;
                    trp = &tr;
                }
                int32_t* fa = r->final_assign ? r->final_assign + (size_t)c * d->n_frag + v.f0 : nullptr;
                int32_t* ia = r->initial_assign ? r->initial_assign + (size_t)c * d->n_frag + v.f0 : nullptr;
                double* il = r->initial_logprob ? r->initial_logprob + sc : nullptr;
                ChainResult cr = run_chain(v, rp, p->seed + c, r->aln_count, r->frag_err_count, r->cluster_hist, fa, trp, ia, il, &lpcache);
                if (trp) for (int i = 0; i < p->num_iters; ++i) if (tr.move_frag[i] >= 0) tr.move_frag[i] += v.f0;
                if (r->final_logprob) r->final_logprob[sc] = cr.final_logprob;
                if (r->counters) {
                    int64_t* k = r->counters + sc * 5;
                    k[0] = cr.naive_proposed; k[1] = cr.naive_accepted; k[2] = cr.sv_proposed; k[3] = cr.sv_accepted;
                    k[4] = cr.reassignments;
                }
                tot[0] += cr.naive_proposed; tot[1] += cr.naive_accepted; tot[2] += cr.sv_proposed; tot[3] += cr.sv_accepted;
                tot[4] += cr.reassignments;
                if (r->rng_state) r->rng_state[sc] = cr.rng_state;
                if (r->error) r->error[sc] = cr.error;
                if (cr.error) worst.store(cr.error);
            }
        }
    };
    if (nthreads <= 1) worker();
    else {
        std::vector<std::thread> th;
        for (int i = 0; i < nthreads; ++i) th.emplace_back(worker);
        for (auto& t : th) t.join();
    }
    if (r->totals) { for (int i = 0; i < 5; ++i) r->totals[i] = tot[i].load(); r->totals[5] = worst.load(); }
    return worst.load();
}

// Exact enumeration of one tiny subproblem: logprobs[s] for every joint state s, fragment 0 the least
// significant digit, digit d in [0, n_f] meaning assignment d-1 (the idea of MultiBreakSV.java:1411-1548,
// without its BASE=10 limitation).  Returns the number of states or -1.
int64_t orc_enumerate(const orc_desc* d, int sub, double perr, int java_int_wrap, double* logprobs, int64_t cap) {
    View v = make_view(d, sub);
    RunParams rp;
    rp.perr = perr; rp.java_int_wrap = java_int_wrap; rp.num_iters = 0; rp.burnin = 10;
    int64_t nstates = 1;
    for (int f = 0; f < v.F; ++f) {
        nstates *= (v.frag_aln_ptr[v.f0 + f + 1] - v.frag_aln_ptr[v.f0 + f]) + 1;
        if (nstates > cap) return -1;
    }
    std::vector<uint32_t> a(d->n_aln, 0), fe(d->n_frag, 0), h(d->n_hist, 0);
    Chain ch(v, rp, 0, a.data(), fe.data(), h.data());
    for (int64_t s = 0; s < nstates; ++s) {
        int64_t t = s;
        for (int f = 0; f < v.F; ++f) {
            int n = ch.nalign(f) + 1;
            ch.assign[f] = (int)(t % n) - 1;
            t /= n;
        }
        ch.rebuild_from_assign();
        logprobs[s] = ch.full_logprob();
    }
    return nstates;
}

// ---- Java-semantics test hooks -----------------------------------------------------------------
double orc_log(double x) { return j_log(x); }
double orc_exp(double x) { return j_exp(x); }
void orc_random_doubles(int64_t seed, int n, double* out) { JRandom r(seed); for (int i = 0; i < n; ++i) out[i] = r.nextDouble(); }
int32_t orc_random_nextint(int64_t seed) { JRandom r(seed); return r.nextInt(); }
void orc_random_bounded(int64_t seed, int32_t bound, int n, int32_t* out) { JRandom r(seed); for (int i = 0; i < n; ++i) out[i] = r.nextInt(bound); }
int32_t orc_string_hash(const char* s) { return java_string_hash(s); }
// keys: n NUL-terminated strings back to back; remove_mask[i]=1 removes key i after all insertions.
// out_order receives the indices of the surviving keys in keySet() order; returns count, capacity via *cap.
int orc_hashmap_order(const char* keys, int n, const uint8_t* remove_mask, int32_t* out_order, int32_t* cap, int32_t* tree_bin) {
    JHashOrder m;
    std::vector<std::string> ks;
    const char* p = keys;
    std::map<std::string, int> first;
    for (int i = 0; i < n; ++i) { ks.emplace_back(p); p += ks.back().size() + 1; if (!first.count(ks.back())) first[ks.back()] = i; m.put(ks.back()); }
    if (remove_mask) for (int i = 0; i < n; ++i) if (remove_mask[i]) m.remove(ks[i]);
    int k = 0;
    for (int id : m.order()) out_order[k++] = first[m.key(id)];
    if (cap) *cap = m.capacity();
    if (tree_bin) *tree_bin = m.tree_bin;
    return k;
}

// ---- reference input side ------------------------------------------------------------------------
struct Loaded { RefModel m; Packed p; std::string err; };

void* orc_load(const char* clusterfile, const char* assignmentfile, const char* experimentfile, int clusters_first,
               char* err, int errlen) {
    Loaded* L = new Loaded();
    try {
        L->m.clusterfile = clusterfile; L->m.assignmentfile = assignmentfile; L->m.experimentfile = experimentfile;
        L->m.clusters_first = clusters_first != 0;
        L->m.load();
        L->p = L->m.pack();
    } catch (const std::exception& e) {
        if (err && errlen > 0) { std::strncpy(err, e.what(), errlen - 1); err[errlen - 1] = 0; }
        delete L;
        return nullptr;
    }
    if (err && errlen > 0) err[0] = 0;
    return L;
}
void orc_free(void* h) { delete (Loaded*)h; }

static const std::vector<int>* i32_field(const Packed& p, const std::string& n) {
    static const std::map<std::string, std::vector<int> Packed::*> tab = {
        {"frag_aln_ptr", &Packed::frag_aln_ptr}, {"sorted_frag", &Packed::sorted_frag},
        {"aln_numerrors", &Packed::aln_numerrors}, {"aln_len", &Packed::aln_len}, {"aln_exp", &Packed::aln_exp},
        {"aln_esp_ptr", &Packed::aln_esp_ptr}, {"aln_esp_id", &Packed::aln_esp_id},
        {"aln_esp_rawcluster", &Packed::aln_esp_rawcluster}, {"aln_esp_cluster", &Packed::aln_esp_cluster},
        {"aln_esp_leaf", &Packed::aln_esp_leaf}, {"supports_order", &Packed::supports_order},
        {"cluster_leaf_ptr", &Packed::cluster_leaf_ptr}, {"leaf_esp_id", &Packed::leaf_esp_id},
        {"leaf_owner", &Packed::leaf_owner}, {"cluster_root_ptr", &Packed::cluster_root_ptr},
        {"root_leaf_ptr", &Packed::root_leaf_ptr}, {"root_leaf", &Packed::root_leaf},
        {"sv_cluster", &Packed::sv_cluster}, {"sv_frag_ptr", &Packed::sv_frag_ptr}, {"sv_frag", &Packed::sv_frag},
        {"sv_numchanged", &Packed::sv_numchanged}, {"hashmap_caps", &Packed::hashmap_caps}};
    auto it = tab.find(n);
    return it == tab.end() ? nullptr : &(p.*(it->second));
}
int orc_i32_len(void* h, const char* name) {
    const auto* v = i32_field(((Loaded*)h)->p, name);
    return v ? (int)v->size() : -1;
}
int orc_i32_get(void* h, const char* name, int32_t* out) {
    const auto* v = i32_field(((Loaded*)h)->p, name);
    if (!v) return -1;
    for (size_t i = 0; i < v->size(); ++i) out[i] = (*v)[i];
    return (int)v->size();
}
// scalars: F A K C E P max_support tree_bin
void orc_scalars(void* h, int32_t* out) {
    const Packed& p = ((Loaded*)h)->p;
    out[0] = p.F; out[1] = p.A; out[2] = p.K; out[3] = p.C; out[4] = p.E; out[5] = p.P; out[6] = p.max_support;
    out[7] = p.tree_bin ? 1 : 0;
}
int orc_u8_cluster_kind(void* h, uint8_t* out) {
    const Packed& p = ((Loaded*)h)->p;
    for (size_t i = 0; i < p.cluster_kind.size(); ++i) out[i] = p.cluster_kind[i];
    return (int)p.cluster_kind.size();
}
int orc_f64_get(void* h, const char* name, double* out) {
    const Packed& p = ((Loaded*)h)->p;
    const std::vector<double>* v = nullptr;
    std::string n = name;
    if (n == "exp_pseq") v = &p.exp_pseq; else if (n == "exp_lambda") v = &p.exp_lambda; else if (n == "logprobcov") v = &p.logprobcov;
    if (!v) return -1;
    for (size_t i = 0; i < v->size(); ++i) out[i] = (*v)[i];
    return (int)v->size();
}
const char* orc_str_get(void* h, const char* name, int i) {
    const Packed& p = ((Loaded*)h)->p;
    std::string n = name;
    const std::vector<std::string>* v = nullptr;
    if (n == "frag_id") v = &p.frag_id; else if (n == "cluster_id") v = &p.cluster_id; else if (n == "exp_label") v = &p.exp_label;
    else if (n == "esp_names") v = &p.esp_names; else if (n == "aln_esp_name") v = &p.aln_esp_name;
    else if (n == "unsupported") return p.unsupported.c_str();
    else if (n == "warning") return ((Loaded*)h)->m.warning.c_str();
    if (!v || i < 0 || i >= (int)v->size()) return nullptr;
    return (*v)[i].c_str();
}
// Literal setBreakpoints from ESP identities for cluster c under `assign`; writes supports as
// out[x*stride + 0] = length, out[x*stride + 1..] = sorted values.  Returns 0, or -5 on the reference's Error.
int orc_literal_supports(void* h, int c, const int32_t* assign, int32_t* out, int stride) {
    const Packed& p = ((Loaded*)h)->p;
    try {
        auto sup = literal_set_breakpoints(p, c, assign);
        for (int x = 0; x < p.E; ++x) {
            out[x * stride] = (int)sup[x].size();
            for (size_t i = 0; i < sup[x].size() && (int)i + 1 < stride; ++i) out[x * stride + 1 + i] = sup[x][i];
        }
    } catch (...) { return -5; }
    return 0;
}

}  // extern "C"
