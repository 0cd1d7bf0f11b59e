// libmbsv_b200.so — core C ABI (include/mbsv_b200.h): descriptor validation, device packing, launch
// planning and the sampler launches.  There is no CPU fallback in this file or anywhere in the product.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <numeric>
#include <atomic>
#include <chrono>
#include <cmath>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/mbsv_b200.h"
#include "jmath.cuh"
#include "sampler_types.cuh"

namespace {

thread_local std::string g_err;
thread_local int64_t g_launches = 0;
// One run launches up to NSTREAM - 1 kernels (one per state-size class) that must overlap.  With the default of 8
// hardware work queues, the 9th stream shares a queue with the 1st and its kernel silently waits for that one to finish
// (measured: a 200 ms tail on a latency-bound batch).  CUDA reads CUDA_DEVICE_MAX_CONNECTIONS when the context is created:
// the HOST sets it to 32 before its first CUDA call (include/mbsv_b200.h "Host setup"; bench.py and the Python package do).
// The library never touches the process environment: setenv in a library constructor races with getenv in a JVM host.

struct LaunchInfo { int rows; int warps; int grid; int smem_bytes; float ms; int memo; };
thread_local std::vector<LaunchInfo> g_launch_info;
thread_local std::vector<long long> g_launch_totals;    // [launch][6]: the counters of results.totals per launch

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                               \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            return fail(_e == cudaErrorMemoryAllocation ? MBSV_ERR_OOM : MBSV_ERR_CUDA,              \
                        std::string(#expr) + ": " + cudaGetErrorString(_e));                         \
    } while (0)

// Device allocations of at least POOL_MIN bytes are kept when a problem releases them and handed to the next problem that
// asks for a block of about that size on the same device: a host that runs one problem after another (the reference's
// workflow is one run per subset of clusters) pays the multi-gigabyte cudaMalloc / cudaFree of the chain-state workspace
// and the memo tables once.  Measured on the 5M-fragment batch: create + run + destroy 5.16 -> see profiles/ (fresh blocks
// also made the first kernel touching them ~250 ms slower than reused ones).  mbsv_trim() returns the kept blocks.
struct DevPool {
    static constexpr size_t POOL_MIN = 1 << 20;
    struct Block { void* p; size_t bytes; int device; };
    std::mutex mu;
    std::vector<Block> free_blocks;
    cudaError_t get(void** out, size_t bytes) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (bytes >= POOL_MIN) {
            std::lock_guard<std::mutex> lk(mu);
            int best = -1;
            for (int i = 0; i < (int)free_blocks.size(); ++i) {
                const Block& b = free_blocks[i];
                if (b.device != dev || b.bytes < bytes || b.bytes > bytes + bytes / 4 + POOL_MIN) continue;
                if (best < 0 || b.bytes < free_blocks[best].bytes) best = i;
            }
            if (best >= 0) {
                *out = free_blocks[best].p;
                live[*out] = free_blocks[best].bytes;
                free_blocks.erase(free_blocks.begin() + best);
                return cudaSuccess;
            }
        }
        cudaError_t e = cudaMalloc(out, bytes);
        if (e == cudaErrorMemoryAllocation) {          // give the kept blocks back and try once more
            cudaGetLastError();
            trim(dev);
            e = cudaMalloc(out, bytes);
        }
        if (e == cudaSuccess && bytes >= POOL_MIN) { std::lock_guard<std::mutex> lk(mu); live[*out] = bytes; }
        return e;
    }
    void put(void* ptr) {
        int dev = 0;
        cudaGetDevice(&dev);
        {
            std::lock_guard<std::mutex> lk(mu);
            auto it = live.find(ptr);
            if (it != live.end()) {
                cudaPointerAttributes a{};
                if (cudaPointerGetAttributes(&a, ptr) == cudaSuccess) dev = a.device;
                free_blocks.push_back({ptr, it->second, dev});
                live.erase(it);
                return;
            }
        }
        cudaFree(ptr);
    }
    void trim(int device) {                             // device < 0: every device
        std::vector<Block> drop;
        {
            std::lock_guard<std::mutex> lk(mu);
            for (size_t i = 0; i < free_blocks.size();)
                if (device < 0 || free_blocks[i].device == device) { drop.push_back(free_blocks[i]); free_blocks.erase(free_blocks.begin() + i); }
                else ++i;
        }
        int cur = 0;
        cudaGetDevice(&cur);
        for (const Block& b : drop) { cudaSetDevice(b.device); cudaFree(b.p); }
        cudaSetDevice(cur);
    }
    size_t kept(int device) {                           // bytes waiting for reuse on `device`
        std::lock_guard<std::mutex> lk(mu);
        size_t b = 0;
        for (const Block& k : free_blocks) if (k.device == device) b += k.bytes;
        return b;
    }
    std::unordered_map<void*, size_t> live;             // pooled-size blocks currently handed out
};
DevPool g_pool;

// fn(begin, end) over [0, n) on up to 16 host threads (the scans of mbsv_problem_create over millions of entries)
template <typename Fn>
void parallel_for(int64_t n, Fn fn) {
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)std::thread::hardware_concurrency(), 16, n / 65536}));
    if (nt <= 1) { fn((int64_t)0, n); return; }
    std::vector<std::thread> th;
    const int64_t chunk = (n + nt - 1) / nt;
    for (int t = 1; t < nt; ++t) th.emplace_back([&, t] { fn(std::min(n, t * chunk), std::min(n, (t + 1) * chunk)); });
    fn((int64_t)0, std::min(n, chunk));
    for (auto& x : th) x.join();
}

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    ~DevBuf() { release(); }
    void release() { if (p) g_pool.put(p); p = nullptr; n = 0; }
    cudaError_t alloc(size_t count) {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        return g_pool.get((void**)&p, count * sizeof(T));
    }
    cudaError_t upload_raw(const T* src, size_t count) {
        cudaError_t e = alloc(count);
        if (e != cudaSuccess || count == 0) return e;
        return cudaMemcpy(p, src, count * sizeof(T), cudaMemcpyHostToDevice);
    }
    cudaError_t upload(const std::vector<T>& v) {
        cudaError_t e = alloc(v.size());
        if (e != cudaSuccess || v.empty()) return e;
        return cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    }
};

constexpr int MAX_EXP = 4;
constexpr int WARPS_PER_CTA = 4;
constexpr int NSTREAM = 20;

}  // namespace

struct mbsv_problem {
    int device = 0;
    mbsv::DevProblem dp{};
    // host-side planning data
    int n_sub = 0, n_frag = 0, n_aln = 0, n_cluster = 0, n_exp = 0, n_hist = 0;
    std::vector<int> sub_W;       // state rows of a chain of subproblem s (F_s + C_s*E + conn-comp blocks)
    std::vector<int> sub_F;       // fragments of subproblem s
    std::vector<int> sub_nsv;     // AlterSV candidates of subproblem s
    std::vector<int> sub_nstates; // joint state space prod(n_f + 1), saturated at 2^30; 0 when it has a conn-comp
    std::vector<int> frag_nal;    // alignments per fragment (validation of a host-supplied initial state)
    std::vector<int> sub_maxent;  // most count entries / toggled fragments of one AlterSV proposal (giant_kernel.cuh)
    std::vector<int> sub_maxpos;  // most counted slots of one alignment
    bool narrow_ok = false;       // every assignment (-1 .. alignments - 1) and every count (<= max_support) fits a signed char
    DevBuf<int4> d_aln_ent4, d_sv_mv;      // flattened views for the speculative-window kernel
    DevBuf<int> d_sv_ent_ptr;
    DevBuf<int2> d_sv_ent;
    // launch plan, depends on mbsv_run_params.memo_states (cached for the last value used)
    int plan_memo = -1, plan_tab = -1;
    std::vector<int> sub_rows;    // tile rows of a warp of subproblem s (state + memo table)
    std::vector<int> sub_memo;    // joint states when the subproblem uses the memoised arithmetic, else 0
    std::vector<char> sub_tab;    // 1: runs in the memoised kernel
    std::vector<long long> sub_tab_off;   // its table: offset (doubles) in the global table buffer, -1 = shared memory
    long long memo_global_doubles = 0;
    std::vector<long long> tab_state_ptr;  // device-memory tables: first global state of each such subproblem (+ total)
    std::vector<int> tab_sub;              // ... and the subproblem
    int tab_scratch_rows = 0;              // max clusters x experiments among them
    DevBuf<long long> d_sub_tab_off, d_tab_state_ptr;
    DevBuf<int> d_tab_sub;
    cudaEvent_t ev_built = nullptr;
    DevBuf<double> d_memo_global;
    // chain-state workspace of the global class and its per-warp row offsets: kept between runs (grow-only), so that a
    // repeated mbsv_run pays no multi-gigabyte cudaMalloc / cudaFree; released with the problem
    DevBuf<int> ws_cache;
    DevBuf<long long> row_off_cache;
    std::vector<long long> row_off_host;
    std::vector<int> sub_order;   // memoised-kernel subproblems first, each group by ascending sub_rows (stable)
    int n_tab_subs = 0;
    int64_t csr_bytes = 0;
    // device CSR
    DevBuf<int> d_sub_frag_ptr, d_sub_cluster_ptr, d_sub_sv_ptr, d_frag_aln_ptr, d_aln_esp_slot, d_supports_local,
        d_cluster_hist_ptr, d_slot_hist_ptr, d_sv_frag_ptr, d_sv_frag, d_sv_numchanged, d_sub_order, d_sub_W, d_sub_memo;
    DevBuf<int4> d_aln_rec;
    DevBuf<double> d_sub_logF, d_sub_neglogNsv, d_T, d_logint, d_invint, d_pseq, d_logp, d_log1mp;
    DevBuf<unsigned long long> d_launch_totals;     // [NSTREAM][6]: every launch pools its chains' counters in its own row
    cudaStream_t streams[NSTREAM] = {};
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr, ev_done[NSTREAM] = {}, ev_cls0[NSTREAM] = {}, ev_cls1[NSTREAM] = {};
    bool streams_ok = false;
    bool has_cc = false;
    DevBuf<int> d_cluster_cc_off, d_sub_cc_scratch, d_sub_cc_maxleaf, d_cluster_leaf_ptr, d_leaf_owner, d_aln_esp_leaf,
        d_cluster_root_ptr, d_root_leaf_ptr, d_root_leaf;
};

extern "C" {

int mbsv_abi_version(void) { return MBSV_ABI_VERSION; }

const char* mbsv_last_error(void) { return g_err.c_str(); }

// used by the host mirror (mbsv_host.cpp) so that both halves of the library report through mbsv_last_error()
void mbsv_internal_set_error(const char* msg) { g_err = msg ? msg : ""; }

const char* mbsv_status_string(int s) {
    switch (s) {
        case MBSV_OK: return "ok";
        case MBSV_ERR_ARG: return "bad argument";
        case MBSV_ERR_CUDA: return "CUDA error";
        case MBSV_ERR_NCCL: return "NCCL error";
        case MBSV_ERR_OOM: return "out of memory";
        case MBSV_ERR_INVARIANT: return "reference invariant violated";
        case MBSV_ERR_UNSUPPORTED: return "unsupported input";
        case MBSV_ERR_IO: return "I/O error";
        default: return "unknown";
    }
}

// Returns the device memory kept for reuse (see DevPool) to the driver; device < 0 = every device.
int mbsv_trim(int32_t device) {
    g_pool.trim(device);
    return MBSV_OK;
}

int mbsv_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail(MBSV_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    return n;
}

double mbsv_log(double x) { return mbsv::jm_log(x); }
double mbsv_exp(double x) { return mbsv::jm_exp(x); }

int mbsv_fill_logprobcov(double lambda, int32_t max_support, double* out) {
    if (!out || max_support < 0) return fail(MBSV_ERR_ARG, "mbsv_fill_logprobcov: bad argument");
    out[0] = 0.0;
    const double ll = mbsv::jm_log(lambda);
    for (int k = 1; k <= max_support; ++k) {
        double val = -lambda + k * ll;                       // MultiBreakSV.java:1645
        for (int i = 1; i <= k; ++i) val -= mbsv::jm_log((double)i);
        out[k] = val;
    }
    return MBSV_OK;
}

int64_t mbsv_last_launch_count(void) { return g_launches; }

// Per-launch details of the last mbsv_run on this thread: out[i*5..] = {state rows per warp (0 = global
// workspace class; negative = memoised kernel), warps, grid, dynamic smem bytes, milliseconds}.  Returns the count.
int mbsv_last_launch_info(double* out, int32_t max_launches) {
    int n = 0;
    for (const auto& li : g_launch_info) {
        if (n >= max_launches) break;
        if (out) { out[n * 5] = li.memo ? -li.rows : li.rows; out[n * 5 + 1] = li.warps; out[n * 5 + 2] = li.grid; out[n * 5 + 3] = li.smem_bytes; out[n * 5 + 4] = li.ms; }
        ++n;
    }
    return n;
}

// out[i*6..] = {Naive proposals, Naive moves, AlterSV proposals, AlterSV moves, fragment-reassignments, first non-zero chain
// status} of the chains of launch i of the last mbsv_run on this thread (same order as mbsv_last_launch_info).
int mbsv_last_launch_totals(int64_t* out, int32_t max_launches) {
    const int n = (int)(g_launch_totals.size() / 6);
    for (int i = 0; i < n && i < max_launches && out; ++i)
        for (int k = 0; k < 6; ++k) out[i * 6 + k] = g_launch_totals[6 * (size_t)i + k];
    return n < max_launches ? n : max_launches;
}

int mbsv_lpt_shard(const int64_t* weights, int64_t n, int32_t n_parts, int32_t* out_part) {
    if (!weights || !out_part || n < 0 || n_parts < 1) return fail(MBSV_ERR_ARG, "mbsv_lpt_shard: bad argument");
    std::vector<int64_t> idx(n);
    std::iota(idx.begin(), idx.end(), 0);
    std::stable_sort(idx.begin(), idx.end(), [&](int64_t a, int64_t b) { return weights[a] > weights[b]; });
    // min-heap on (load, part)
    std::vector<std::pair<int64_t, int32_t>> heap;
    for (int32_t p = 0; p < n_parts; ++p) heap.push_back({0, p});
    auto cmp = [](const std::pair<int64_t, int32_t>& a, const std::pair<int64_t, int32_t>& b) { return a > b; };
    std::make_heap(heap.begin(), heap.end(), cmp);
    for (int64_t i : idx) {
        std::pop_heap(heap.begin(), heap.end(), cmp);
        auto& top = heap.back();
        out_part[i] = top.second;
        top.first += weights[i];
        std::push_heap(heap.begin(), heap.end(), cmp);
    }
    return MBSV_OK;
}

static int validate(const mbsv_problem_desc* d) {
    if (!d) return fail(MBSV_ERR_ARG, "descriptor is NULL");
    if (d->abi_version != MBSV_ABI_VERSION) return fail(MBSV_ERR_ARG, "abi_version mismatch");
    if (d->n_sub < 1 || d->n_frag < 1 || d->n_aln < 1 || d->n_cluster < 0 || d->n_exp < 1 || d->max_support < 0)
        return fail(MBSV_ERR_ARG, "descriptor sizes out of range");
    if (d->n_exp > MAX_EXP) return fail(MBSV_ERR_UNSUPPORTED, "more than 4 experiments");
    if (!d->sub_frag_ptr || !d->sub_cluster_ptr || !d->frag_aln_ptr || !d->aln_numerrors || !d->aln_len || !d->aln_exp ||
        !d->aln_esp_ptr || (d->n_aln_esp > 0 && !d->aln_esp_cluster) || (d->n_cluster > 0 && (!d->cluster_kind ||
        !d->supports_order || !d->cluster_hist_ptr)) || !d->exp_pseq || !d->logprobcov)
        return fail(MBSV_ERR_ARG, "descriptor has a NULL required array");
    if (d->n_sv > 0 && (!d->sub_sv_ptr || !d->sv_frag_ptr || !d->sv_frag || !d->sv_numchanged))
        return fail(MBSV_ERR_ARG, "descriptor has n_sv > 0 but NULL AlterSV arrays");
    auto mono = [&](const int32_t* p, int n, int last, const char* name) -> int {
        if (p[0] != 0 || p[n] != last) return fail(MBSV_ERR_ARG, std::string(name) + ": bad first/last offset");
        std::atomic<int> bad{0};
        parallel_for(n, [&](int64_t b, int64_t e) { for (int64_t i = b; i < e; ++i) if (p[i + 1] < p[i]) { bad = 1; return; } });
        if (bad) return fail(MBSV_ERR_ARG, std::string(name) + ": not monotone");
        return 0;
    };
    int rc;
    if ((rc = mono(d->sub_frag_ptr, d->n_sub, d->n_frag, "sub_frag_ptr"))) return rc;
    if ((rc = mono(d->sub_cluster_ptr, d->n_sub, d->n_cluster, "sub_cluster_ptr"))) return rc;
    if ((rc = mono(d->frag_aln_ptr, d->n_frag, d->n_aln, "frag_aln_ptr"))) return rc;
    if ((rc = mono(d->aln_esp_ptr, d->n_aln, d->n_aln_esp, "aln_esp_ptr"))) return rc;
    if (d->n_cluster > 0 && (rc = mono(d->cluster_hist_ptr, d->n_cluster, d->n_hist, "cluster_hist_ptr"))) return rc;
    if (d->n_sv > 0) {
        if ((rc = mono(d->sub_sv_ptr, d->n_sub, d->n_sv, "sub_sv_ptr"))) return rc;
        if ((rc = mono(d->sv_frag_ptr, d->n_sv, d->n_sv_frag, "sv_frag_ptr"))) return rc;
    }
    for (int s = 0; s < d->n_sub; ++s)
        if (d->sub_frag_ptr[s + 1] == d->sub_frag_ptr[s])
            return fail(MBSV_ERR_ARG, "a subproblem has no fragments (the reference would index an empty keySet, "
                                      "MultiBreakSV.java:1401)");
    {
        std::atomic<int> bad{0};
        parallel_for(d->n_frag, [&](int64_t b, int64_t e) {
            for (int64_t f = b; f < e; ++f) if (d->frag_aln_ptr[f + 1] == d->frag_aln_ptr[f]) { bad = 1; return; }
        });
        if (bad) return fail(MBSV_ERR_ARG, "a fragment has no alignment");
    }
    bool any_cc = false;
    for (int c = 0; c < d->n_cluster; ++c) {
        if (d->cluster_kind[c] > 1) return fail(MBSV_ERR_ARG, "cluster_kind must be 0 (cluster) or 1 (connected component)");
        any_cc = any_cc || d->cluster_kind[c] == 1;
    }
    if (any_cc) {
        if (!d->cluster_leaf_ptr || !d->leaf_owner || !d->aln_esp_leaf || !d->cluster_root_ptr || !d->root_leaf_ptr || !d->root_leaf)
            return fail(MBSV_ERR_ARG, "connected components need the leaf/root arrays and aln_esp_leaf");
        if ((rc = mono(d->cluster_leaf_ptr, d->n_cluster, d->n_leaf, "cluster_leaf_ptr"))) return rc;
        if ((rc = mono(d->cluster_root_ptr, d->n_cluster, d->n_root, "cluster_root_ptr"))) return rc;
        if ((rc = mono(d->root_leaf_ptr, d->n_root, d->n_root_leaf, "root_leaf_ptr"))) return rc;
        for (int c = 0; c < d->n_cluster; ++c) {
            if (d->cluster_kind[c] != 1) continue;
            const int nl = d->cluster_leaf_ptr[c + 1] - d->cluster_leaf_ptr[c];
            if (nl > d->max_support)
                return fail(MBSV_ERR_UNSUPPORTED, "a connected component has more leaves than max_support");
            if (nl + 1 > d->cluster_hist_ptr[c + 1] - d->cluster_hist_ptr[c])
                return fail(MBSV_ERR_ARG, "cluster_hist_ptr leaves too few bins for a connected component");
            for (int r = d->cluster_root_ptr[c]; r < d->cluster_root_ptr[c + 1]; ++r)
                for (int j = d->root_leaf_ptr[r]; j < d->root_leaf_ptr[r + 1]; ++j)
                    if (d->root_leaf[j] < 0 || d->root_leaf[j] >= nl) return fail(MBSV_ERR_ARG, "root_leaf out of range");
        }
    }
    // Experiment.logprobcov[0] is never written by the reference (Java zero-initialises it, MultiBreakSV.java:612-614) and
    // the Delta arithmetic of MBSV_MODE_FAST relies on it
    for (int x = 0; x < d->n_exp; ++x)
        if (d->logprobcov[(size_t)x * (d->max_support + 1)] != 0.0)
            return fail(MBSV_ERR_ARG, "logprobcov[x][0] must be 0 (MultiBreakSV.java:612-614 fills k = 1..maxSupport)");
    {
        std::atomic<int> bad{0};
        parallel_for(d->n_aln, [&](int64_t b, int64_t e) {
            for (int64_t a = b; a < e; ++a) {
                if (d->aln_exp[a] < 0 || d->aln_exp[a] >= d->n_exp) { bad = 1; return; }
                if (d->aln_esp_ptr[a + 1] - d->aln_esp_ptr[a] >= (1 << 23)) { bad = 2; return; }
            }
        });
        if (bad == 1) return fail(MBSV_ERR_ARG, "aln_exp out of range");
        if (bad == 2) return fail(MBSV_ERR_ARG, "alignment with too many ESPs");
    }
    return 0;
}

int mbsv_problem_create(const mbsv_problem_desc* d, int32_t device, mbsv_problem** out) {
    if (!out) return fail(MBSV_ERR_ARG, "out is NULL");
    *out = nullptr;
    static const bool dbg = std::getenv("MBSV_DEBUG_TIMING") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!dbg) return;
        auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[mbsv create] %-10s %.3f s\n", what, std::chrono::duration<double>(t1 - t0).count());
        t0 = t1;
    };
    int rc = validate(d);
    if (rc) return rc;
    lap("validate");
    const int S = d->n_sub, F = d->n_frag, A = d->n_aln, K = d->n_aln_esp, C = d->n_cluster, E = d->n_exp;
    // Packing + per-subproblem checks, parallel over subproblems (they are independent and contiguous).
    // uninitialised buffers: the worker threads touch (and page in) their own parts
    // (page-locked staging buffers were measured: the upload drops from ~75 to ~40 ms, but the worker threads then write into
    // pages one thread touched, and packing rises from ~100 to ~170 ms: ordinary memory, first touched by its worker, stays)
    std::unique_ptr<int4[]> rec(new (std::nothrow) int4[(size_t)A]);
    std::unique_ptr<int[]> slot(new (std::nothrow) int[(size_t)K + 1]);
    std::unique_ptr<int[]> sup_local(new (std::nothrow) int[(size_t)C + 1]);
    if (!rec || !slot || !sup_local) return fail(MBSV_ERR_OOM, "host allocation failed");
    int maxn = 1;
    std::vector<int> sub_maxent(S, 0), sub_maxpos(S, 0);
    {
        const int nthreads = (int)std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
        std::vector<int> t_maxn(nthreads, 1);
        std::vector<int> t_err(nthreads, 0);
        std::vector<std::string> t_msg(nthreads);
        std::atomic<int> next{0};
        const int CHUNK = 1024;
        const int BIG_SUB = 32768;       // fragments: a larger subproblem (a giant component) is packed by all threads together
        auto worker = [&](int tid) {
            std::vector<int> in_aln, stamp_aln, per_frag, stamp_frag, total;   // indexed by local cluster
            std::vector<char> seen;
            std::vector<int> touched;
            auto bad = [&](int code, const char* msg) { if (!t_err[tid]) { t_err[tid] = code; t_msg[tid] = msg; } };
            for (;;) {
                const int s_begin = next.fetch_add(CHUNK);
                if (s_begin >= S) break;
                const int s_end = std::min(S, s_begin + CHUNK);
                for (int s = s_begin; s < s_end; ++s) {
                    const int c0 = d->sub_cluster_ptr[s], c1 = d->sub_cluster_ptr[s + 1], Cs = c1 - c0;
                    if (d->sub_frag_ptr[s + 1] - d->sub_frag_ptr[s] > BIG_SUB) continue;      // packed by all threads below
                    if ((int)in_aln.size() < Cs) {
                        in_aln.resize(Cs); stamp_aln.resize(Cs); per_frag.resize(Cs); stamp_frag.resize(Cs); total.resize(Cs); seen.resize(Cs);
                    }
                    for (int i = 0; i < Cs; ++i) { stamp_aln[i] = -1; stamp_frag[i] = -1; total[i] = 0; seen[i] = 0; }
                    int max_pos = 0;             // most counted ESP entries of one alignment
                    for (int f = d->sub_frag_ptr[s]; f < d->sub_frag_ptr[s + 1]; ++f) {
                        t_maxn[tid] = std::max(t_maxn[tid], d->frag_aln_ptr[f + 1] - d->frag_aln_ptr[f]);
                        touched.clear();
                        for (int a = d->frag_aln_ptr[f]; a < d->frag_aln_ptr[f + 1]; ++a) {
                            const int x = d->aln_exp[a];
                            const int j0 = d->aln_esp_ptr[a], j1 = d->aln_esp_ptr[a + 1];
                            rec[a] = make_int4(d->aln_numerrors[a], d->aln_len[a], j0, x | ((j1 - j0) << 8));
                            int pos = 0;
                            for (int j = j0; j < j1; ++j) {
                                const int c = d->aln_esp_cluster[j];
                                slot[j] = -1;
                                if (c != -1) max_pos = std::max(max_pos, ++pos);
                                if (c == -1) continue;
                                if (c < c0 || c >= c1) { bad(MBSV_ERR_ARG, "aln_esp_cluster points outside the fragment's subproblem"); continue; }
                                const int cl = c - c0;
                                if (d->cluster_kind[c] == 1) {
                                    const int nl = d->cluster_leaf_ptr[c + 1] - d->cluster_leaf_ptr[c];
                                    if (d->aln_esp_leaf[j] < 0 || d->aln_esp_leaf[j] >= nl) bad(MBSV_ERR_ARG, "aln_esp_leaf out of range");
                                    slot[j] = -(2 + cl);               // connected component: recomputed, not counted
                                    continue;
                                }
                                slot[j] = cl * E + x;
                                // upper bound of the cluster's support: per fragment the max over its alignments of the
                                // entries into the cluster, summed over fragments
                                if (stamp_aln[cl] != a) { stamp_aln[cl] = a; in_aln[cl] = 0; }
                                ++in_aln[cl];
                                if (stamp_frag[cl] != f) { stamp_frag[cl] = f; per_frag[cl] = 0; touched.push_back(cl); }
                                per_frag[cl] = std::max(per_frag[cl], in_aln[cl]);
                            }
                        }
                        for (int cl : touched) total[cl] += per_frag[cl];
                    }
                    sub_maxpos[s] = max_pos;
                    for (int cl = 0; cl < Cs; ++cl) {
                        if (total[cl] > d->max_support)
                            bad(MBSV_ERR_UNSUPPORTED, "a cluster can collect more selected ESPs than max_support "
                                                      "(the reference would index logprobcov out of bounds)");
                        if (total[cl] + 1 > d->cluster_hist_ptr[c0 + cl + 1] - d->cluster_hist_ptr[c0 + cl])
                            bad(MBSV_ERR_ARG, "cluster_hist_ptr leaves too few bins for a cluster");
                    }
                    for (int i = c0; i < c1; ++i) {
                        const int c = d->supports_order[i];
                        if (c < c0 || c >= c1 || seen[c - c0]) { bad(MBSV_ERR_ARG, "supports_order is not a permutation of the subproblem's clusters"); continue; }
                        seen[c - c0] = 1;
                        sup_local[i] = c - c0;
                    }
                }
            }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nthreads; ++t) th.emplace_back(worker, t);
        worker(0);
        for (auto& t : th) t.join();
        // Subproblems too large for one thread: the same packing and checks, the fragments spread over the threads; a
        // fragment's per-cluster maxima come from its (few) entries sorted by cluster instead of per-cluster stamp arrays.
        for (int s = 0; s < S; ++s) {
            const int fs0 = d->sub_frag_ptr[s], fs1 = d->sub_frag_ptr[s + 1];
            if (fs1 - fs0 <= BIG_SUB) continue;
            const int c0 = d->sub_cluster_ptr[s], c1 = d->sub_cluster_ptr[s + 1], Cs = c1 - c0;
            std::unique_ptr<std::atomic<int>[]> total(new (std::nothrow) std::atomic<int>[(size_t)Cs + 1]);
            if (!total) return fail(MBSV_ERR_OOM, "host allocation failed");
            for (int i = 0; i < Cs; ++i) total[i].store(0, std::memory_order_relaxed);
            std::vector<int> t_pos(nthreads, 0);
            std::atomic<int> nextf{fs0};
            auto big_worker = [&](int tid) {
                auto bad = [&](int code, const char* msg) { if (!t_err[tid]) { t_err[tid] = code; t_msg[tid] = msg; } };
                std::vector<std::pair<int, int>> ent;        // (cluster, alignment) of the fragment's counted entries
                for (;;) {
                    const int fb = nextf.fetch_add(4096);
                    if (fb >= fs1) break;
                    for (int f = fb; f < std::min(fs1, fb + 4096); ++f) {
                        t_maxn[tid] = std::max(t_maxn[tid], d->frag_aln_ptr[f + 1] - d->frag_aln_ptr[f]);
                        ent.clear();
                        for (int a = d->frag_aln_ptr[f]; a < d->frag_aln_ptr[f + 1]; ++a) {
                            const int x = d->aln_exp[a];
                            const int j0 = d->aln_esp_ptr[a], j1 = d->aln_esp_ptr[a + 1];
                            rec[a] = make_int4(d->aln_numerrors[a], d->aln_len[a], j0, x | ((j1 - j0) << 8));
                            int pos = 0;
                            for (int j = j0; j < j1; ++j) {
                                const int c = d->aln_esp_cluster[j];
                                slot[j] = -1;
                                if (c != -1) t_pos[tid] = std::max(t_pos[tid], ++pos);
                                if (c == -1) continue;
                                if (c < c0 || c >= c1) { bad(MBSV_ERR_ARG, "aln_esp_cluster points outside the fragment's subproblem"); continue; }
                                const int cl = c - c0;
                                if (d->cluster_kind[c] == 1) {
                                    const int nl = d->cluster_leaf_ptr[c + 1] - d->cluster_leaf_ptr[c];
                                    if (d->aln_esp_leaf[j] < 0 || d->aln_esp_leaf[j] >= nl) bad(MBSV_ERR_ARG, "aln_esp_leaf out of range");
                                    slot[j] = -(2 + cl);
                                    continue;
                                }
                                slot[j] = cl * E + x;
                                ent.emplace_back(cl, a);
                            }
                        }
                        std::sort(ent.begin(), ent.end());
                        for (size_t i = 0; i < ent.size();) {                 // per cluster: the most entries of one alignment
                            const int cl = ent[i].first;
                            int best = 0;
                            while (i < ent.size() && ent[i].first == cl) {
                                const int a = ent[i].second;
                                int n = 0;
                                while (i < ent.size() && ent[i].first == cl && ent[i].second == a) { ++n; ++i; }
                                best = std::max(best, n);
                            }
                            total[cl].fetch_add(best, std::memory_order_relaxed);
                        }
                    }
                }
            };
            std::vector<std::thread> tb;
            for (int t = 1; t < nthreads; ++t) tb.emplace_back(big_worker, t);
            big_worker(0);
            for (auto& t : tb) t.join();
            sub_maxpos[s] = *std::max_element(t_pos.begin(), t_pos.end());
            std::vector<char> seen(Cs, 0);
            for (int cl = 0; cl < Cs; ++cl) {
                const int tot = total[cl].load(std::memory_order_relaxed);
                if (tot > d->max_support)
                    return fail(MBSV_ERR_UNSUPPORTED, "a cluster can collect more selected ESPs than max_support "
                                                      "(the reference would index logprobcov out of bounds)");
                if (tot + 1 > d->cluster_hist_ptr[c0 + cl + 1] - d->cluster_hist_ptr[c0 + cl])
                    return fail(MBSV_ERR_ARG, "cluster_hist_ptr leaves too few bins for a cluster");
            }
            for (int i = c0; i < c1; ++i) {
                const int c = d->supports_order[i];
                if (c < c0 || c >= c1 || seen[c - c0]) return fail(MBSV_ERR_ARG, "supports_order is not a permutation of the subproblem's clusters");
                seen[c - c0] = 1;
                sup_local[i] = c - c0;
            }
        }
        for (int t = 0; t < nthreads; ++t) {
            if (t_err[t]) return fail(t_err[t], t_msg[t]);
            maxn = std::max(maxn, t_maxn[t]);
        }
    }
    const int max_nal = maxn;            // most alignments of one fragment (maxn grows below for the log table)
    lap("pack");
    std::vector<int> sub_sv_ptr(S + 1, 0), sv_frag_ptr(1, 0), sv_frag, sv_numchanged, sv_ent_ptr(1, 0);
    std::vector<int4> sv_mv;
    std::vector<int2> sv_ent;
    if (d->n_sv > 0) {
        sub_sv_ptr.assign(d->sub_sv_ptr, d->sub_sv_ptr + S + 1);
        sv_frag_ptr.assign(d->sv_frag_ptr, d->sv_frag_ptr + d->n_sv + 1);
        sv_frag.assign(d->sv_frag, d->sv_frag + d->n_sv_frag);
        sv_numchanged.assign(d->sv_numchanged, d->sv_numchanged + d->n_sv);
        for (int s = 0; s < S; ++s)
            for (int k = sub_sv_ptr[s]; k < sub_sv_ptr[s + 1]; ++k) {
                int entries = 0;
                for (int q = sv_frag_ptr[k]; q < sv_frag_ptr[k + 1]; ++q) {
                    const int f = sv_frag[q];
                    if (f < d->sub_frag_ptr[s] || f >= d->sub_frag_ptr[s + 1] ||
                        d->frag_aln_ptr[f + 1] - d->frag_aln_ptr[f] != 1)
                        return fail(MBSV_ERR_ARG, "sv_frag must list single-alignment fragments of the same subproblem");
                    for (int q2 = sv_frag_ptr[k]; q2 < q; ++q2)
                        if (sv_frag[q2] == f) return fail(MBSV_ERR_ARG, "sv_frag lists a fragment twice for one cluster");
                    const int a = d->frag_aln_ptr[f];
                    sv_mv.push_back(make_int4(f - d->sub_frag_ptr[s], d->aln_numerrors[a], d->aln_len[a], d->aln_exp[a]));
                    for (int j = d->aln_esp_ptr[a]; j < d->aln_esp_ptr[a + 1]; ++j)
                        if (slot[j] >= 0) { sv_ent.push_back(make_int2(slot[j], q - sv_frag_ptr[k])); ++entries; }
                }
                sv_ent_ptr.push_back((int)sv_ent.size());
                sub_maxent[s] = std::max(sub_maxent[s], std::max(entries, sv_frag_ptr[k + 1] - sv_frag_ptr[k]));
            }
        if ((int)sv_ent_ptr.size() != d->n_sv + 1) return fail(MBSV_ERR_ARG, "sub_sv_ptr does not cover every AlterSV candidate");
    }
    // the exact-binomial branch (n*p <= 5 or n*(1-p) <= 5, MultiBreakSV.java:1664) only sees n <= 5/min(p, 1-p)
    for (int x = 0; x < E; ++x) {
        const double pmin = std::min(d->exp_pseq[x], 1 - d->exp_pseq[x]);
        if (pmin > 0) maxn = std::max(maxn, (int)std::min(65536.0, 5.0 / pmin + 2));
    }
    std::vector<double> logF(S), neglogNsv(S, 0.0), logint(maxn + 1, 0.0), logp(E), log1mp(E);
    parallel_for(S, [&](int64_t b, int64_t e) {
        for (int64_t s = b; s < e; ++s) {
            logF[s] = mbsv::jm_log((double)(d->sub_frag_ptr[s + 1] - d->sub_frag_ptr[s]));
            const int nsv = sub_sv_ptr[s + 1] - sub_sv_ptr[s];
            if (nsv > 0) neglogNsv[s] = -mbsv::jm_log((double)nsv);
        }
    });
    for (int n = 1; n <= maxn; ++n) logint[n] = mbsv::jm_log((double)n);
    std::vector<double> invint(maxn + 1);
    for (int n = 0; n <= maxn; ++n) invint[n] = (double)1 / n;
    for (int x = 0; x < E; ++x) { logp[x] = mbsv::jm_log(d->exp_pseq[x]); log1mp[x] = mbsv::jm_log(1 - d->exp_pseq[x]); }

    lap("sv+logs");
    // every host-side check is done; from here on a CUDA device is required (no CPU fallback)
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(MBSV_ERR_CUDA, "no such CUDA device (this library has no CPU fallback)");
    CUDA_TRY(cudaSetDevice(device));

    mbsv_problem* p = new (std::nothrow) mbsv_problem();
    if (!p) return fail(MBSV_ERR_OOM, "host allocation failed");
    p->device = device;
    p->n_sub = S; p->n_frag = F; p->n_aln = A; p->n_cluster = C; p->n_exp = E; p->n_hist = d->n_hist;
    p->sub_W.resize(S);
    // chain state rows: assignments + counts (+ per conn-comp {stamp, current and saved multisets} + set-cover scratch)
    std::vector<int> cluster_cc_off(C, -1), sub_cc_scratch(S, -1), sub_cc_maxleaf(S, 0);
    p->sub_nstates.resize(S); p->sub_F.resize(S); p->sub_nsv.resize(S);
    std::atomic<int> any_cc{0};
    parallel_for(S, [&](int64_t sb, int64_t se) {
        for (int64_t s = sb; s < se; ++s) {
            int W = (d->sub_frag_ptr[s + 1] - d->sub_frag_ptr[s]) + (d->sub_cluster_ptr[s + 1] - d->sub_cluster_ptr[s]) * E;
            int off = 0, maxleaf = 0, maxroot = 0;
            for (int c = d->sub_cluster_ptr[s]; c < d->sub_cluster_ptr[s + 1]; ++c) {
                if (d->cluster_kind[c] != 1) continue;
                const int nr = d->cluster_root_ptr[c + 1] - d->cluster_root_ptr[c];
                cluster_cc_off[c] = off;
                off += 1 + 2 * E * (1 + nr);
                maxleaf = std::max(maxleaf, d->cluster_leaf_ptr[c + 1] - d->cluster_leaf_ptr[c]);
                maxroot = std::max(maxroot, nr);
            }
            if (off > 0) { sub_cc_scratch[s] = off; sub_cc_maxleaf[s] = maxleaf; W += off + maxleaf + maxroot; any_cc = 1; }
            p->sub_W[s] = W;
            long long nst = off > 0 ? 0 : 1;
            for (int f = d->sub_frag_ptr[s]; f < d->sub_frag_ptr[s + 1] && nst > 0 && nst < (1LL << 30); ++f)
                nst *= (long long)(d->frag_aln_ptr[f + 1] - d->frag_aln_ptr[f]) + 1;
            p->sub_nstates[s] = (int)std::min<long long>(nst, 1LL << 30);
            p->sub_F[s] = d->sub_frag_ptr[s + 1] - d->sub_frag_ptr[s];
            p->sub_nsv[s] = sub_sv_ptr[s + 1] - sub_sv_ptr[s];
        }
    });
    if (any_cc) p->has_cc = true;
    p->sub_maxent = sub_maxent;
    p->frag_nal.resize(F);
    parallel_for(F, [&](int64_t b, int64_t e2) { for (int64_t f = b; f < e2; ++f) p->frag_nal[f] = d->frag_aln_ptr[f + 1] - d->frag_aln_ptr[f]; });
    lap("rows");
    p->sub_maxpos = sub_maxpos;
    p->narrow_ok = max_nal <= 127 && d->max_support <= 127;

    auto up = [&](auto& buf, const auto& vec) -> cudaError_t { return buf.upload(vec); };
    cudaError_t e = cudaSuccess;
    auto chk = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    chk(p->d_sub_frag_ptr.upload_raw(d->sub_frag_ptr, S + 1));
    chk(p->d_sub_cluster_ptr.upload_raw(d->sub_cluster_ptr, S + 1));
    chk(up(p->d_sub_sv_ptr, sub_sv_ptr));
    chk(p->d_frag_aln_ptr.upload_raw(d->frag_aln_ptr, F + 1));
    chk(p->d_aln_rec.upload_raw(rec.get(), A));
    chk(p->d_aln_esp_slot.upload_raw(slot.get(), K));
    chk(p->d_supports_local.upload_raw(sup_local.get(), C));
    if (C > 0) chk(p->d_cluster_hist_ptr.upload_raw(d->cluster_hist_ptr, C + 1)); else chk(up(p->d_cluster_hist_ptr, std::vector<int>(1, 0)));
    {
        std::vector<int> slot_hist((size_t)std::max(C, 0) * E + 1, 0);
        for (int c = 0; c < C; ++c) for (int x = 0; x < E; ++x) slot_hist[(size_t)c * E + x] = d->cluster_hist_ptr[c];
        chk(up(p->d_slot_hist_ptr, slot_hist));
    }
    chk(up(p->d_sv_frag_ptr, sv_frag_ptr));
    chk(up(p->d_sv_frag, sv_frag));
    chk(up(p->d_sv_numchanged, sv_numchanged));
    chk(up(p->d_sv_mv, sv_mv));
    chk(up(p->d_sv_ent_ptr, sv_ent_ptr));
    chk(up(p->d_sv_ent, sv_ent));
    chk(up(p->d_sub_logF, logF));
    chk(up(p->d_sub_neglogNsv, neglogNsv));
    {   // padded to a multiple of 16 bytes: the speculative-window kernel stages it with one bulk copy (cp.async.bulk)
        std::vector<double> T(d->logprobcov, d->logprobcov + (size_t)E * (d->max_support + 1));
        T.resize((T.size() + 1) & ~(size_t)1, 0.0);
        chk(up(p->d_T, T));
    }
    chk(up(p->d_logint, logint));
    chk(up(p->d_invint, invint));
    chk(up(p->d_pseq, std::vector<double>(d->exp_pseq, d->exp_pseq + E)));
    chk(up(p->d_logp, logp));
    chk(up(p->d_log1mp, log1mp));
    if (p->has_cc) {
        chk(up(p->d_cluster_cc_off, cluster_cc_off));
        chk(up(p->d_sub_cc_scratch, sub_cc_scratch));
        chk(up(p->d_sub_cc_maxleaf, sub_cc_maxleaf));
        chk(up(p->d_cluster_leaf_ptr, std::vector<int>(d->cluster_leaf_ptr, d->cluster_leaf_ptr + C + 1)));
        chk(up(p->d_leaf_owner, std::vector<int>(d->leaf_owner, d->leaf_owner + d->n_leaf)));
        chk(up(p->d_aln_esp_leaf, std::vector<int>(d->aln_esp_leaf, d->aln_esp_leaf + K)));
        chk(up(p->d_cluster_root_ptr, std::vector<int>(d->cluster_root_ptr, d->cluster_root_ptr + C + 1)));
        chk(up(p->d_root_leaf_ptr, std::vector<int>(d->root_leaf_ptr, d->root_leaf_ptr + d->n_root + 1)));
        chk(up(p->d_root_leaf, std::vector<int>(d->root_leaf, d->root_leaf + d->n_root_leaf)));
    }
    chk(up(p->d_sub_W, p->sub_W));
    lap("upload");
    if (e != cudaSuccess) {
        delete p;
        return fail(e == cudaErrorMemoryAllocation ? MBSV_ERR_OOM : MBSV_ERR_CUDA, std::string("upload: ") + cudaGetErrorString(e));
    }
    p->csr_bytes = (int64_t)(S + 1) * 12 + (int64_t)(F + 1) * 4 + (int64_t)A * 16 + (int64_t)K * 4 + (int64_t)C * 8;
    mbsv::DevProblem& dp = p->dp;
    dp.n_sub = S; dp.n_frag = F; dp.n_aln = A; dp.n_cluster = C; dp.n_exp = E; dp.max_support = d->max_support; dp.maxn = maxn;
    dp.sub_frag_ptr = p->d_sub_frag_ptr.p; dp.sub_cluster_ptr = p->d_sub_cluster_ptr.p; dp.sub_sv_ptr = p->d_sub_sv_ptr.p;
    dp.sub_logF = p->d_sub_logF.p; dp.sub_neglogNsv = p->d_sub_neglogNsv.p; dp.frag_aln_ptr = p->d_frag_aln_ptr.p; dp.aln_rec = p->d_aln_rec.p;
    dp.aln_esp_slot = p->d_aln_esp_slot.p; dp.supports_order_local = p->d_supports_local.p;
    dp.cluster_cc_off = p->d_cluster_cc_off.p; dp.sub_cc_scratch = p->d_sub_cc_scratch.p; dp.sub_cc_maxleaf = p->d_sub_cc_maxleaf.p;
    dp.cluster_leaf_ptr = p->d_cluster_leaf_ptr.p; dp.leaf_owner = p->d_leaf_owner.p; dp.aln_esp_leaf = p->d_aln_esp_leaf.p;
    dp.cluster_root_ptr = p->d_cluster_root_ptr.p; dp.root_leaf_ptr = p->d_root_leaf_ptr.p; dp.root_leaf = p->d_root_leaf.p;
    dp.cluster_hist_ptr = p->d_cluster_hist_ptr.p; dp.slot_hist_ptr = p->d_slot_hist_ptr.p; dp.sv_frag_ptr = p->d_sv_frag_ptr.p; dp.sv_frag = p->d_sv_frag.p;
    dp.sv_mv = p->d_sv_mv.p; dp.sv_ent_ptr = p->d_sv_ent_ptr.p; dp.sv_ent = p->d_sv_ent.p; dp.aln_ent4 = nullptr;
    dp.sv_numchanged = p->d_sv_numchanged.p; dp.T = p->d_T.p; dp.logint = p->d_logint.p; dp.invint = p->d_invint.p; dp.exp_pseq = p->d_pseq.p;
    dp.exp_logp = p->d_logp.p; dp.exp_log1mp = p->d_log1mp.p;
    dp.log2pi = mbsv::jm_log(2.0) + mbsv::jm_log(3.141592653589793);
    int prio_lo = 0, prio_hi = 0;
    chk(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    for (int i = 0; i < NSTREAM; ++i) {
        // streams[1] runs the longest-running class (large subproblems) and gets the highest priority
        chk(cudaStreamCreateWithPriority(&p->streams[i], cudaStreamNonBlocking, i == 1 ? prio_hi : prio_lo));
        chk(cudaEventCreateWithFlags(&p->ev_done[i], cudaEventDisableTiming));
        chk(cudaEventCreate(&p->ev_cls0[i]));
        chk(cudaEventCreate(&p->ev_cls1[i]));
    }
    chk(cudaEventCreateWithFlags(&p->ev_built, cudaEventDisableTiming));
    chk(cudaEventCreate(&p->ev_start));
    chk(cudaEventCreate(&p->ev_stop));
    // The uploads above are cudaMemcpy calls from pageable host memory on the legacy stream: they return when the data is
    // staged, not when the last DMA has landed, and the kernels run on non-blocking streams that do not wait for it.
    chk(cudaStreamSynchronize(0));
    if (e != cudaSuccess) { delete p; return fail(MBSV_ERR_CUDA, std::string("stream setup: ") + cudaGetErrorString(e)); }
    p->streams_ok = true;
    *out = p;
    return MBSV_OK;
}

int mbsv_problem_destroy(mbsv_problem* p) {
    if (!p) return MBSV_OK;
    cudaSetDevice(p->device);
    if (p->streams_ok) {
        for (int i = 0; i < NSTREAM; ++i) {
            cudaStreamDestroy(p->streams[i]); cudaEventDestroy(p->ev_done[i]);
            cudaEventDestroy(p->ev_cls0[i]); cudaEventDestroy(p->ev_cls1[i]);
        }
        cudaEventDestroy(p->ev_built);
        cudaEventDestroy(p->ev_start);
        cudaEventDestroy(p->ev_stop);
    }
    delete p;
    return MBSV_OK;
}

int mbsv_problem_stats(const mbsv_problem* p, int64_t* out, int32_t n) {
    if (!p || !out) return fail(MBSV_ERR_ARG, "mbsv_problem_stats: NULL");
    int64_t memo = 0;
    for (int v2 : p->sub_memo) memo += v2 > 0;
    int64_t v[9] = {p->n_sub, p->n_frag, p->n_aln, p->n_cluster, p->csr_bytes,
                    p->sub_W.empty() ? 0 : *std::max_element(p->sub_W.begin(), p->sub_W.end()), memo, p->n_hist, p->device};
    for (int i = 0; i < n && i < 9; ++i) out[i] = v[i];
    return MBSV_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// launch
// ------------------------------------------------------------------------------------------------
namespace {

using mbsv::KernelFn;
KernelFn pick(int mode, bool trace, bool smem, int E, int rows, bool narrow_ok, size_t* extra_smem, int* state_bytes) {
    *extra_smem = 0;
    *state_bytes = 4;
    if (trace) return mbsv::get_kernel_trace(mode, smem, E);
    return mode == 0 ? mbsv::get_kernel_replay(smem, E) : mbsv::get_kernel_fast(smem, E, rows, narrow_ok, extra_smem, state_bytes);
}

// (Re)build the launch plan: which subproblems use the memoised arithmetic (memo_states), which of those run in the
// memoised kernel (only when every warp holds chains of one subproblem, i.e. n_chains % 32 == 0, and the table fits
// the largest shared-memory class), tile rows per warp, launch order.
constexpr int MAX_SMEM_ROWS = 128;     // largest shared-memory size class (rows per chain)
constexpr int SMEM_TABLE_ROWS = mbsv::MEMO_SMEM_TABLE_ROWS;   // tiles up to this many rows keep the table in shared memory
// tab_mode: 0 = no memoised kernel (memo-eligible subproblems re-sum the posterior per proposal in the general kernel),
// 1 = warps of one subproblem (strides and, when small, the table in the warp's shared memory), 2 = mixed warps (every
// lane its own subproblem: strides in the lane's column, every table in device memory).  The arithmetic is the same in
// all three, so the choice may depend on the batch shape and on free memory.
int ensure_plan(mbsv_problem* p, int memo_states, int tab_mode) {
    if (p->plan_memo == memo_states && p->plan_tab == tab_mode) return 0;
    const int S = p->n_sub;
    p->sub_rows.resize(S); p->sub_memo.assign(S, 0); p->sub_tab.assign(S, 0); p->sub_order.resize(S);
    p->sub_tab_off.assign(S, -1);
    p->n_tab_subs = 0;
    p->memo_global_doubles = 0;
    p->tab_state_ptr.clear(); p->tab_sub.clear(); p->tab_scratch_rows = 1;
    // device-memory tables may take a quarter of what is free now (the chain-state workspace needs the rest)
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    free_b += g_pool.kept(p->device);                  // blocks kept for reuse are available to this problem
    const long long table_budget = (long long)(free_b / 4 / sizeof(double));
    for (int s = 0; s < S; ++s) {
        int rows = p->sub_W[s];
        const int nst = p->sub_nstates[s];
        if (memo_states > 0 && nst > 0 && nst <= memo_states) {
            p->sub_memo[s] = nst;
            // fragment strides always sit after the state rows; the log-posterior table follows them when the
            // whole tile still fits the small size classes, else it lives in the global table buffer
            const int fr = (p->sub_F[s] + 1) & ~1;
            // one-fragment subproblems without AlterSV candidates also tabulate the acceptance ratios (nst^2)
            // (64-bit: nst may be 65536; a ratio table that cannot fit simply makes with_table exceed SMEM_TABLE_ROWS)
            // so do small subproblems of several fragments without AlterSV candidates (memo_kernel.cuh `rt`); if their ratio
            // table does not fit, the log-posterior table alone may
            const bool multi_rt = p->sub_F[s] > 1 && p->sub_nsv[s] == 0 && nst <= mbsv::MEMO_RT_MAX_STATES;
            long long ratio_words = ((p->sub_F[s] == 1 && p->sub_nsv[s] == 0) || multi_rt) ? 2LL * nst * nst : 0;
            long long with_table = rows + (fr + 2LL * nst + ratio_words + 31) / 32;
            if (multi_rt && with_table > SMEM_TABLE_ROWS) { ratio_words = 0; with_table = rows + (fr + 2LL * nst + 31) / 32; }
            int trows = 0;
            bool global_table = true;
            if (tab_mode == 1 && with_table <= SMEM_TABLE_ROWS) { trows = (int)with_table; global_table = false; }
            else if (tab_mode == 1) trows = rows + (fr + 31) / 32;
            else if (tab_mode == 2) trows = rows + p->sub_F[s];
            // the memoised kernel keeps its state in shared memory: larger tiles (many clusters) stay in the general kernel
            if (tab_mode != 0 && trows <= MAX_SMEM_ROWS && (!global_table || p->memo_global_doubles + nst <= table_budget)) {
                rows = trows;
                p->sub_tab[s] = 1;
                ++p->n_tab_subs;
                if (global_table) {
                    p->sub_tab_off[s] = p->memo_global_doubles;
                    p->tab_state_ptr.push_back(p->memo_global_doubles);
                    p->tab_sub.push_back(s);
                    p->tab_scratch_rows = std::max(p->tab_scratch_rows, p->sub_W[s] - p->sub_F[s]);
                    p->memo_global_doubles += nst;
                }
            }
        }
        p->sub_rows[s] = rows;
    }
    std::iota(p->sub_order.begin(), p->sub_order.end(), 0);
    std::stable_sort(p->sub_order.begin(), p->sub_order.end(), [&](int a, int b) {
        if (p->sub_tab[a] != p->sub_tab[b]) return p->sub_tab[a] > p->sub_tab[b];
        return p->sub_rows[a] < p->sub_rows[b];
    });
    CUDA_TRY(p->d_sub_order.upload(p->sub_order));
    CUDA_TRY(p->d_sub_memo.upload(p->sub_memo));
    CUDA_TRY(p->d_sub_tab_off.upload(p->sub_tab_off));
    CUDA_TRY(p->d_memo_global.alloc((size_t)p->memo_global_doubles));
    p->tab_state_ptr.push_back(p->memo_global_doubles);
    CUDA_TRY(p->d_tab_state_ptr.upload(p->tab_state_ptr));
    CUDA_TRY(p->d_tab_sub.upload(p->tab_sub));
    p->plan_memo = memo_states;
    p->plan_tab = tab_mode;
    return 0;
}

struct OutBufs {
    DevBuf<unsigned int> aln_count, frag_err, hist;
    DevBuf<int> final_assign, error, tr_frag, tr_assign, init_assign_out, init_assign_in;
    DevBuf<double> final_logprob, tr_logprob, init_logprob;
    DevBuf<long long> counters;
    DevBuf<unsigned long long> rng_state, totals;
};

int run_impl(mbsv_problem* p, const mbsv_run_params* rp, mbsv_results* res, bool results_on_device) {
    g_launches = 0;
    g_launch_info.clear();
    if (!p || !rp || !res) return fail(MBSV_ERR_ARG, "mbsv_run: NULL argument");
    static const bool dbg_timing = std::getenv("MBSV_DEBUG_TIMING") != nullptr;
    auto lap_t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!dbg_timing) return;
        auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[mbsv run] %-12s %.3f s\n", what, std::chrono::duration<double>(t1 - lap_t0).count());
        lap_t0 = t1;
    };
    if (rp->mode != MBSV_MODE_REPLAY_EXACT && rp->mode != MBSV_MODE_FAST && rp->mode != MBSV_MODE_GIBBS) return fail(MBSV_ERR_ARG, "unknown mode");
    const bool gibbs = rp->mode == MBSV_MODE_GIBBS;
    if (gibbs && rp->trace) return fail(MBSV_ERR_ARG, "MBSV_MODE_GIBBS has no per-iteration trace");
    if (rp->n_chains < 1 || rp->num_iters < 0 || rp->burnin < 0) return fail(MBSV_ERR_ARG, "bad run parameters");
    if (!(rp->perr >= 0 && rp->perr <= 1)) return fail(MBSV_ERR_ARG, "--perr must be a value between 0 and 1.");
    if (rp->device != p->device) return fail(MBSV_ERR_ARG, "run device differs from the problem's device");
    const int64_t win_base = std::max<int64_t>((int64_t)rp->num_iters - rp->burnin + 1, 0);
    const int64_t nwin = std::max<int64_t>((int64_t)rp->num_iters - win_base, 0);
    if (nwin * rp->n_chains >= (1LL << 32)) return fail(MBSV_ERR_ARG, "window length x chains overflows the 32-bit tallies");
    if (rp->trace && (!res->trace_logprob || !res->trace_move_frag || !res->trace_move_assign))
        return fail(MBSV_ERR_ARG, "trace requested without trace buffers");
    if (rp->memo_states < 0) return fail(MBSV_ERR_ARG, "memo_states must be >= 0");
    CUDA_TRY(cudaSetDevice(p->device));
    const int S = p->n_sub, X = rp->n_chains;
    const int64_t n_items = (int64_t)S * X;
    // Small batches run fewer chains per warp: one warp per scheduler issues an instruction only every ~8 cycles
    // (each chain is one long dependency chain), so until the GPU holds its ~20 resident warps per SM more warps
    // with idle lanes are faster than full warps.  L = chains per warp, a power of two.
    int L = 32;
    int64_t resident_warps = 148 * 20;
    {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device);
        resident_warps = (int64_t)sms * 20;
        while (L > 1 && (n_items + L / 2 - 1) / (L / 2) <= resident_warps) L /= 2;
        if (const char* env = std::getenv("MBSV_LANES_PER_WARP")) L = std::max(1, std::min(32, std::atoi(env)));
    }
    // The memoised kernel runs either full warps of ONE subproblem (n_chains a multiple of 32: strides and small tables in
    // the warp's shared memory) or mixed warps where every lane holds its own (subproblem, chain) with the tables in device
    // memory — any other chain count, and small batches.  (Half / quarter warps of one subproblem for 16 / 8 chains were
    // measured 35 % slower than mixed full warps on the 16-chain long-read config.)
    const int Lu = X % 32 == 0 ? 32 : 0;
    const bool mixed = !(Lu > 0 && L == 32) || std::getenv("MBSV_FORCE_MIXED") != nullptr;   // (env: tuning aid)
    const int Lm = mixed ? L : Lu;
    if (int prc = gibbs ? ensure_plan(p, 0, 0) : ensure_plan(p, rp->memo_states, mixed ? 2 : 1)) return prc;
    const int64_t tab_items = (int64_t)p->n_tab_subs * X;
    const int64_t sc_total = n_items;

    lap("plan");
    OutBufs ob;
    mbsv::DevRun dr{};
    dr.n_chains = X; dr.num_iters = rp->num_iters; dr.burnin = rp->burnin; dr.win_base = (int)win_base;
    dr.java_int_wrap = rp->java_int_wrap; dr.perr = rp->perr; dr.logperr = mbsv::jm_log(rp->perr);
    dr.log1mperr = mbsv::jm_log(1 - rp->perr); dr.seed = rp->seed;
    {   // ceil(perr * 2^53); the product is exact
        const double x = rp->perr * 9007199254740992.0;
        dr.perr_thr53 = !(x > 0.0) ? 0ULL : x >= 9007199254740992.0 ? (1ULL << 53) : (unsigned long long)std::ceil(x);
    }
    dr.sub_order = p->d_sub_order.p; dr.sub_W = p->d_sub_W.p; dr.sub_memo = p->d_sub_memo.p; dr.n_items = n_items;
    dr.sub_tab_off = p->d_sub_tab_off.p; dr.memo_global = p->d_memo_global.p;
    dr.lanes_per_warp = L;

    if (rp->init_assign) {
        if (results_on_device) dr.init_assign = rp->init_assign;
        else {
            // the kernels index the alignment records with it: every entry must be -1 or an alignment of its fragment
            for (int f = 0; f < p->n_frag; ++f)
                if (rp->init_assign[f] < -1 || rp->init_assign[f] >= p->frag_nal[f])
                    return fail(MBSV_ERR_ARG, "init_assign: entry outside [-1, alignments of the fragment)");
            CUDA_TRY(ob.init_assign_in.alloc(p->n_frag));
            CUDA_TRY(cudaMemcpy(ob.init_assign_in.p, rp->init_assign, sizeof(int) * p->n_frag, cudaMemcpyHostToDevice));
            dr.init_assign = ob.init_assign_in.p;
        }
    }
    // outputs
    const size_t trN = rp->trace ? (size_t)sc_total * rp->num_iters : 0;
    if (results_on_device) {
        dr.aln_count = res->aln_count; dr.frag_err_count = res->frag_err_count; dr.cluster_hist = res->cluster_hist;
        dr.final_assign = res->final_assign; dr.final_logprob = res->final_logprob; dr.counters = (long long*)res->counters;
        dr.rng_state = (unsigned long long*)res->rng_state; dr.error = res->error;
        dr.trace_logprob = res->trace_logprob; dr.trace_move_frag = res->trace_move_frag; dr.trace_move_assign = res->trace_move_assign;
        dr.initial_assign = res->initial_assign; dr.initial_logprob = res->initial_logprob;
        dr.totals = (unsigned long long*)res->totals;
        if (!dr.totals) return fail(MBSV_ERR_ARG, "mbsv_run_device needs results.totals (the chain status is read from it)");
        if (!dr.aln_count || !dr.frag_err_count || !dr.cluster_hist) return fail(MBSV_ERR_ARG, "tally buffers are required");
    } else {
        CUDA_TRY(ob.aln_count.alloc(p->n_aln)); CUDA_TRY(ob.frag_err.alloc(p->n_frag)); CUDA_TRY(ob.hist.alloc(std::max(1, p->n_hist)));
        dr.aln_count = ob.aln_count.p; dr.frag_err_count = ob.frag_err.p; dr.cluster_hist = ob.hist.p;
        if (res->final_assign) { CUDA_TRY(ob.final_assign.alloc((size_t)X * p->n_frag)); dr.final_assign = ob.final_assign.p; }
        if (res->final_logprob) { CUDA_TRY(ob.final_logprob.alloc(sc_total)); dr.final_logprob = ob.final_logprob.p; }
        if (res->counters) { CUDA_TRY(ob.counters.alloc(sc_total * 5)); dr.counters = ob.counters.p; }
        if (res->rng_state) { CUDA_TRY(ob.rng_state.alloc(sc_total)); dr.rng_state = ob.rng_state.p; }
        if (res->error) { CUDA_TRY(ob.error.alloc(sc_total)); dr.error = ob.error.p; }
        CUDA_TRY(ob.totals.alloc(6)); dr.totals = ob.totals.p;
        if (rp->trace) {
            CUDA_TRY(ob.tr_logprob.alloc(trN)); CUDA_TRY(ob.tr_frag.alloc(trN)); CUDA_TRY(ob.tr_assign.alloc(trN));
            dr.trace_logprob = ob.tr_logprob.p; dr.trace_move_frag = ob.tr_frag.p; dr.trace_move_assign = ob.tr_assign.p;
        }
        if (res->initial_assign) { CUDA_TRY(ob.init_assign_out.alloc((size_t)X * p->n_frag)); dr.initial_assign = ob.init_assign_out.p; }
        if (res->initial_logprob) { CUDA_TRY(ob.init_logprob.alloc(sc_total)); dr.initial_logprob = ob.init_logprob.p; }
    }
    lap("out alloc");
    cudaStream_t s0 = p->streams[0];
    CUDA_TRY(cudaMemsetAsync(dr.aln_count, 0, sizeof(unsigned) * p->n_aln, s0));
    CUDA_TRY(cudaMemsetAsync(dr.frag_err_count, 0, sizeof(unsigned) * p->n_frag, s0));
    if (p->n_hist > 0) CUDA_TRY(cudaMemsetAsync(dr.cluster_hist, 0, sizeof(unsigned) * p->n_hist, s0));
    if (dr.error) CUDA_TRY(cudaMemsetAsync(dr.error, 0, sizeof(int) * sc_total, s0));
    if (!p->d_launch_totals.p) CUDA_TRY(p->d_launch_totals.alloc(6 * NSTREAM));
    CUDA_TRY(cudaMemsetAsync(p->d_launch_totals.p, 0, sizeof(unsigned long long) * 6 * NSTREAM, s0));

    // ---- plan: warps of 32 consecutive (sorted subproblem, chain) items; size classes by state rows ----
    // 8-bit state words for some shared-memory classes of the MODE_FAST kernels (kern_fast.cu) when every value fits;
    // MBSV_STATE_BITS=32 forces int (tests, A/B)
    const char* bits_env = std::getenv("MBSV_STATE_BITS");
    const bool narrow = p->narrow_ok && !(bits_env && std::atoi(bits_env) == 32);
    std::vector<int> cls = {16, 24, 32, 48, 64, 96, 128};
    if (const char* env = std::getenv("MBSV_SMEM_CLASSES")) {
        cls.clear();
        for (const char* q = env; *q;) { cls.push_back(std::atoi(q)); while (*q && *q != ',') ++q; if (*q == ',') ++q; }
    }
    // Classes are cut in (sorted subproblem) space, separately for the memoised group [0, n_tab_subs) and the rest; every
    // class then picks its own chains-per-warp.  Starting from the batch-wide L, classes are revisited from the largest
    // state down and their L halved while the GPU still has unfilled warp slots: the few chains of the big classes are the
    // critical path of a small batch, and a warp of k diverged chains runs each of them ~k times slower.
    struct Launch { int wb, we, rows; bool smem, memo; int64_t i0, i1; int lanes; bool giant = false; };
    std::vector<Launch> launches;
    auto add_group = [&](int k0, int k1, bool memo) {
        int lo = k0;
        auto rows_of = [&](int k) { return std::max(1, p->sub_rows[p->sub_order[k]]); };
        for (size_t c = 0; c < cls.size() && lo < k1; ++c) {
            int a = lo, b = k1;
            while (a < b) { int m = a + (b - a) / 2; if (rows_of(m) <= cls[c]) a = m + 1; else b = m; }
            if (a > lo) launches.push_back({0, 0, cls[c], true, memo, (int64_t)lo * X, (int64_t)a * X, memo ? Lm : L});
            lo = a;
        }
        if (lo < k1) {
            launches.push_back({0, 0, 0, false, memo, (int64_t)lo * X, (int64_t)k1 * X, memo ? Lm : L});
            // Few chains on large states (a giant component): one CTA per chain evaluates windows of consecutive iterations
            // speculatively (giant_kernel.cuh) instead of one lane walking them one memory round trip at a time.  Up to
            // 8 CTAs per SM's worth of chains; beyond that the lane-per-chain kernel has the better throughput.
            const char* gmax_env = std::getenv("MBSV_GIANT_MAX_CHAINS");
            const int64_t gmax = gmax_env ? std::atoll(gmax_env) : resident_warps / 20 * 8;
            bool ok = !memo && rp->mode == MBSV_MODE_FAST && !rp->trace && !p->has_cc && (int64_t)(k1 - lo) * X <= gmax;
            for (int k = lo; ok && k < k1; ++k) {
                const int s = p->sub_order[k];
                ok = p->sub_memo[s] == 0 && p->sub_maxent[s] <= mbsv::giant_max_entries() &&
                     p->sub_maxpos[s] <= mbsv::giant_max_aln_entries();
            }
            ok = ok && (int64_t)p->n_cluster * p->n_exp < (1 << 25);          // slot and fragment index share a word
            if (ok) { launches.back().giant = true; launches.back().lanes = 1; }
        }
    };
    if (gibbs) {
        // a group of G lanes per (subproblem, chain): 4 for the smallest states (two or three candidates per fragment),
        // 8 for small ones, the whole warp otherwise; a class = the ints of shared memory a group needs (state + one
        // double per option).  `lanes` holds the chains per warp = 32 / G.
        const int extra = 2 * (p->dp.maxn + 2);
        static const int gcls[] = {64, 128, 256, 512, 1024, 2048, 4096, 8192, 12288};
        int lo = 0;
        const bool all_global = std::getenv("MBSV_GIBBS_GLOBAL") != nullptr;       // (tests: every state in the workspace)
        for (int c : gcls) {
            if (all_global) break;
            int a = lo, b = S;
            while (a < b) { const int m = a + (b - a) / 2; if (p->sub_W[p->sub_order[m]] + extra <= c) a = m + 1; else b = m; }
            int group = c <= 128 ? 4 : c <= 1024 ? 8 : 32;
            if (const char* env = std::getenv("MBSV_GIBBS_GROUP")) group = std::atoi(env) == 4 ? 4 : std::atoi(env) == 8 ? 8 : 32;   // (tests)
            if ((size_t)WARPS_PER_CTA * (32 / group) * c * sizeof(int) > 200 * 1024) group = 32;
            if (a > lo) launches.push_back({0, 0, c, true, false, (int64_t)lo * X, (int64_t)a * X, 32 / group});
            lo = a;
        }
        // larger states: one chain per warp, the state a contiguous column of the global workspace (rows = the ints of shared
        // memory a warp keeps for its option weights)
        if (lo < S) launches.push_back({0, 0, extra, false, false, (int64_t)lo * X, (int64_t)S * X, 1});
    } else {
        add_group(0, p->n_tab_subs, true);
        add_group(p->n_tab_subs, S, false);
    }
    // largest state first: the global-workspace class (few, slow chains) starts before the bulk of small chains
    std::stable_sort(launches.begin(), launches.end(), [](const Launch& a, const Launch& b) {
        const int ra = a.smem ? a.rows : 1 << 30, rb = b.smem ? b.rows : 1 << 30;
        return ra > rb;
    });
    {
        auto warps_of = [](const Launch& l, int lanes) { return (l.i1 - l.i0 + lanes - 1) / lanes; };
        int64_t total = 0;
        for (const Launch& l : launches) total += warps_of(l, l.lanes);
        int64_t spare = resident_warps - total;
        // Small batches only (spare > 0): the step ends when the slowest class ends, and a class's time grows with its
        // chains per warp (measured on the 100k-fragment single-chain config: general kernel 190 / 211 / 305 ms at 1-2 / 4 /
        // 8 chains per warp, memoised kernel about half of that).  So: narrow the class with the largest estimate while
        // warp slots are left; when they run out, widen the class that stays cheapest after widening to free some.
        auto est = [&](const Launch& l, int lanes) { return (l.memo ? 0.5 : 1.0) * (1.0 + 0.1 * (lanes - 1)); };
        auto movable = [&](const Launch& l) { return !(l.memo && !mixed) && !l.giant; };     // warps of one subproblem keep their width
        for (int guard = 0; !gibbs && spare >= 0 && guard < 256; ++guard) {
            Launch* worst = nullptr;
            for (Launch& l : launches)
                if (movable(l) && l.lanes > 1 && (!worst || est(l, l.lanes) > est(*worst, worst->lanes))) worst = &l;
            if (!worst) break;
            double top = 0;
            for (const Launch& l : launches) top = std::max(top, est(l, l.lanes));
            if (est(*worst, worst->lanes) < top) break;                          // the slowest class cannot be narrowed further
            const int64_t extra = warps_of(*worst, worst->lanes / 2) - warps_of(*worst, worst->lanes);
            if (extra <= spare) { spare -= extra; worst->lanes /= 2; continue; }
            Launch* give = nullptr;                                              // free slots: widen the cheapest class
            for (Launch& l : launches)
                if (movable(l) && &l != worst && l.lanes < 32 && est(l, l.lanes * 2) < est(*worst, worst->lanes) &&
                    warps_of(l, l.lanes) > warps_of(l, l.lanes * 2) &&
                    (!give || est(l, l.lanes * 2) < est(*give, give->lanes * 2))) give = &l;
            if (!give) break;
            spare += warps_of(*give, give->lanes) - warps_of(*give, give->lanes * 2);
            give->lanes *= 2;
        }
        if (!gibbs && std::getenv("MBSV_LANES_PER_WARP")) for (Launch& l : launches) if (!l.giant) l.lanes = l.memo ? Lm : L;
        int64_t w = 0;
        for (Launch& l : launches) {
            if (l.memo && !l.smem) return fail(MBSV_ERR_ARG, "internal: memoised subproblem does not fit the shared-memory classes");
            l.wb = (int)w;
            w += warps_of(l, l.lanes);
            if (w > 0x7fffffff / 32) return fail(MBSV_ERR_ARG, "too many chains for one launch");
            l.we = (int)w;
        }
    }
    // the global-workspace class: one row range per warp (rows of its largest = last subproblem)
    for (const Launch& l : launches) {
        if (l.smem) continue;
        std::vector<long long> off(l.we - l.wb);
        long long acc = 0;
        for (int w = l.wb; w < l.we; ++w) {
            const int64_t last = std::min<int64_t>(l.i0 + (int64_t)(w - l.wb) * l.lanes + l.lanes - 1, l.i1 - 1);
            off[w - l.wb] = acc;
            acc += std::max(1, p->sub_rows[p->sub_order[last / X]]);
        }
        const bool packed = l.lanes == 1 && !rp->trace;     // one chain per warp: contiguous columns (kern_packed.cu)
        const size_t need = (size_t)acc * (packed ? 1 : 32);
        if (need > p->ws_cache.n) {
            p->ws_cache.release();
            size_t free_b = 0, total_b = 0;
            CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
            if (need * sizeof(int) > free_b + g_pool.kept(p->device)) return fail(MBSV_ERR_OOM, "chain state workspace does not fit in device memory");
            if (p->ws_cache.alloc(need) != cudaSuccess) { cudaGetLastError(); p->ws_cache.release(); return fail(MBSV_ERR_OOM, "chain state workspace does not fit in device memory"); }
        }
        if (off != p->row_off_host) {
            CUDA_TRY(p->row_off_cache.upload(off));
            p->row_off_host = off;
        }
    }
    if ((int)launches.size() > NSTREAM - 1) return fail(MBSV_ERR_ARG, "too many size classes");
    lap("classes+ws");

    for (const Launch& l : launches)
        if (!l.memo && !gibbs && !p->d_aln_ent4.p) {   // first launch of the general / speculative-window kernel: flatten the alignments' slots
            CUDA_TRY(p->d_aln_ent4.alloc(p->n_aln));
            CUDA_TRY(mbsv::launch_giant_prep(p->dp, p->d_aln_ent4.p, s0));
            p->dp.aln_ent4 = p->d_aln_ent4.p;
            ++g_launches;
        }
    CUDA_TRY(cudaEventRecord(p->ev_start, s0));
    // device-memory memo tables are filled once per run by a fully parallel kernel; only the launches that read them
    // wait for it
    const bool have_global_tables = !p->tab_sub.empty();
    if (have_global_tables) {
        mbsv::DevRun rb = dr;
        rb.warp_base = 0; rb.item_base = 0;
        CUDA_TRY(mbsv::launch_memo_build(p->dp, rb, (int)p->tab_sub.size(), p->d_tab_state_ptr.p, p->d_tab_sub.p,
                                         p->memo_global_doubles, p->tab_scratch_rows, s0));
        ++g_launches;
        CUDA_TRY(cudaEventRecord(p->ev_built, s0));
        CUDA_TRY(cudaEventRecord(p->ev_cls0[0], s0));      // (timing of the build, MBSV_DEBUG_TIMING)
    }
    int nstream = 0;
    for (const Launch& l : launches) {
        mbsv::DevRun r = dr;
        r.warp_begin = l.wb; r.warp_end = l.we;
        r.rows_per_warp = (l.smem || gibbs) ? l.rows : 0;
        r.warp_row_off = l.smem ? nullptr : p->row_off_cache.p;
        r.workspace = l.smem ? nullptr : p->ws_cache.p;
        r.global_warp0 = l.wb;
        r.warp_base = l.wb;
        r.item_base = l.i0;
        r.n_items = l.i1;                       // the last warp of a class may be partly filled
        r.lanes_per_warp = l.lanes;
        r.memo_mixed = mixed ? 1 : 0;
        r.totals = p->d_launch_totals.p + 6 * (size_t)nstream;
        int giant_threads = 0, giant_cs = 1;
        size_t giant_smem = 0;
        if (l.giant) {
            // CTAs per chain: a cluster of CS CTAs evaluates windows CS times as long.  As many as the GPU has SMs for
            // (every CTA takes a whole SM), capped where stale reads would cut most windows short: the expected number of
            // proposals reading a row written earlier in the same window grows like (window length)^2 / fragments.
            int sms = 148;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device);
            const int64_t chains = l.i1 - l.i0;
            int minF = 1 << 30;
            for (int64_t it = l.i0; it < l.i1; it += X) minF = std::min(minF, p->sub_F[p->sub_order[it / X]]);
            const double by_conflicts = std::sqrt(32.0 * minF) / mbsv::giant_window_offsets();
            while (giant_cs < 8 && (int64_t)giant_cs * 2 * chains <= sms && giant_cs * 2 <= by_conflicts) giant_cs *= 2;
            if (const char* env = std::getenv("MBSV_GIANT_CLUSTER")) {
                const int v = std::atoi(env);
                if (v == 1 || v == 2 || v == 4 || v == 8) giant_cs = v;
            }
        }
        size_t extra_smem = 0;
        int sb = 4;                   // bytes of a state word of the kernel chosen (1: kern_fast8.cu)
        KernelFn fn = gibbs ? mbsv::get_kernel_gibbs(32 / l.lanes, !l.smem)
                     : l.giant ? mbsv::get_kernel_giant(p->n_exp, giant_cs, &giant_threads, &giant_smem)
                     : l.memo ? mbsv::get_kernel_memo(rp->trace != 0, mixed)
                     : (!l.smem && l.lanes == 1 && !rp->trace) ? mbsv::get_kernel_packed(rp->mode, p->has_cc)
                     : p->has_cc ? mbsv::get_kernel_cc(rp->mode, rp->trace != 0, l.smem)
                                 : pick(rp->mode, rp->trace != 0, l.smem, p->n_exp, l.rows, narrow, &extra_smem, &sb);
        const size_t shm = gibbs ? (size_t)WARPS_PER_CTA * (l.smem ? l.lanes : 1) * l.rows * sizeof(int) : l.giant ? giant_smem
                           : (l.smem ? (size_t)WARPS_PER_CTA * l.rows * 32 * sb : 0) + extra_smem;
        if (shm > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
        // MBSV_SERIAL_LAUNCHES=1 (tuning aid): run the size classes back to back on one stream
        static const bool serial = std::getenv("MBSV_SERIAL_LAUNCHES") != nullptr;
        cudaStream_t st = serial ? p->streams[1] : p->streams[1 + nstream];
        CUDA_TRY(cudaStreamWaitEvent(st, (l.memo && have_global_tables) ? p->ev_built : p->ev_start, 0));
        // giant: one CTA (or one cluster of giant_cs CTAs) per chain
        const int grid = l.giant ? (l.we - l.wb) * giant_cs : (l.we - l.wb + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
        CUDA_TRY(cudaEventRecord(p->ev_cls0[1 + nstream], st));
        if (l.giant) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(grid); cfg.blockDim = dim3(giant_threads); cfg.dynamicSmemBytes = shm; cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = giant_cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            CUDA_TRY(cudaLaunchKernelEx(&cfg, fn, p->dp, r));
        } else {
            fn<<<grid, WARPS_PER_CTA * 32, shm, st>>>(p->dp, r);
        }
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(p->ev_cls1[1 + nstream], st));
        g_launch_info.push_back({l.smem ? l.rows : 0, l.we - l.wb, grid, (int)shm, 0.f, l.memo ? 1 : 0});
        ++g_launches;
        CUDA_TRY(cudaEventRecord(p->ev_done[1 + nstream], st));
        CUDA_TRY(cudaStreamWaitEvent(s0, p->ev_done[1 + nstream], 0));
        ++nstream;
    }
    CUDA_TRY(cudaEventRecord(p->ev_stop, s0));
    lap("launch");
    CUDA_TRY(cudaStreamSynchronize(s0));
    lap("kernels");
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, p->ev_start, p->ev_stop));
    res->kernel_ms = ms;
    for (size_t i = 0; i < g_launch_info.size(); ++i)
        cudaEventElapsedTime(&g_launch_info[i].ms, p->ev_cls0[1 + i], p->ev_cls1[1 + i]);
    if (std::getenv("MBSV_DEBUG_TIMING"))
        for (size_t i = 0; i < g_launch_info.size(); ++i) {
            float t0 = 0, t1 = 0;
            cudaEventElapsedTime(&t0, p->ev_start, p->ev_cls0[1 + i]);
            cudaEventElapsedTime(&t1, p->ev_start, p->ev_cls1[1 + i]);
            std::fprintf(stderr, "[mbsv run] class rows %d%s warps %d: start %.1f end %.1f ms\n", g_launch_info[i].rows,
                         g_launch_info[i].memo ? " (memo)" : "", g_launch_info[i].warps, t0, t1);
        }
    if (have_global_tables && std::getenv("MBSV_DEBUG_TIMING")) {
        float bms = 0;
        cudaEventElapsedTime(&bms, p->ev_start, p->ev_cls0[0]);
        std::fprintf(stderr, "[mbsv run] table build: %zu subproblems, %lld states, %.3f ms; lanes/warp %d (memo %d%s)\n",
                     p->tab_sub.size(), p->memo_global_doubles, bms, L, Lm, mixed ? ", mixed" : "");
    }

    // counters per launch -> totals; chain status
    long long htot[6] = {0, 0, 0, 0, 0, 0};
    g_launch_totals.assign(6 * (size_t)nstream, 0);
    if (nstream > 0) CUDA_TRY(cudaMemcpy(g_launch_totals.data(), p->d_launch_totals.p, sizeof(long long) * 6 * nstream, cudaMemcpyDeviceToHost));
    for (int i = 0; i < nstream; ++i) {
        for (int k = 0; k < 5; ++k) htot[k] += g_launch_totals[6 * i + k];
        if (!htot[5]) htot[5] = g_launch_totals[6 * i + 5];
    }
    if (results_on_device) CUDA_TRY(cudaMemcpy(dr.totals, htot, sizeof(htot), cudaMemcpyHostToDevice));
    const int worst = (int)htot[5];
    if (!results_on_device) {
        auto down = [&](void* dst, const void* src, size_t bytes) -> cudaError_t {
            if (!dst || !src || !bytes) return cudaSuccess;
            return cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost);
        };
        CUDA_TRY(down(res->aln_count, dr.aln_count, sizeof(unsigned) * p->n_aln));
        CUDA_TRY(down(res->frag_err_count, dr.frag_err_count, sizeof(unsigned) * p->n_frag));
        CUDA_TRY(down(res->cluster_hist, dr.cluster_hist, sizeof(unsigned) * p->n_hist));
        CUDA_TRY(down(res->final_assign, dr.final_assign, sizeof(int) * (size_t)X * p->n_frag));
        CUDA_TRY(down(res->final_logprob, dr.final_logprob, sizeof(double) * sc_total));
        CUDA_TRY(down(res->counters, dr.counters, sizeof(long long) * sc_total * 5));
        CUDA_TRY(down(res->rng_state, dr.rng_state, sizeof(unsigned long long) * sc_total));
        CUDA_TRY(down(res->error, dr.error, sizeof(int) * sc_total));
        if (res->totals) std::memcpy(res->totals, htot, sizeof(htot));
        CUDA_TRY(down(res->trace_logprob, dr.trace_logprob, sizeof(double) * trN));
        CUDA_TRY(down(res->trace_move_frag, dr.trace_move_frag, sizeof(int) * trN));
        CUDA_TRY(down(res->trace_move_assign, dr.trace_move_assign, sizeof(int) * trN));
        CUDA_TRY(down(res->initial_assign, dr.initial_assign, sizeof(int) * (size_t)X * p->n_frag));
        CUDA_TRY(down(res->initial_logprob, dr.initial_logprob, sizeof(double) * sc_total));
    }
    lap("d2h");
    if (worst) return fail(worst, "a chain hit an invariant the reference turns into an Error (see results.error)");
    return MBSV_OK;
}

}  // namespace

extern "C" {
int mbsv_run(mbsv_problem* p, const mbsv_run_params* params, mbsv_results* results) {
    return run_impl(p, params, results, false);
}
int mbsv_run_device(mbsv_problem* p, const mbsv_run_params* params, mbsv_results* results) {
    return run_impl(p, params, results, true);
}
}
