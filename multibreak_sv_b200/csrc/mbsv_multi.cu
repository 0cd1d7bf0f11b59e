// libmbsv_b200.so — the giant-component multi-GPU path: a single connected component cannot be
// split spatially, so its chains are split over the GPUs and the integer tallies are summed onto one GPU with ONE
// ncclReduce over NVLink.  The reference has no counterpart (its only advice is to wait, doc/MultiBreakSV-Manual.txt:420-423).
//
// Two host shapes:
//   * one process per GPU (torchrun, MPI, ...): mbsv_comm_unique_id / mbsv_comm_create / mbsv_reduce_tallies;
//   * one process, several GPUs (a JVM calling through Panama FFM): mbsv_run_multi.
// NCCL is loaded with dlopen on first use, so the library itself has no link-time dependency on it (single-GPU hosts never
// touch it); a missing or failing NCCL is MBSV_ERR_NCCL.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mbsv_b200.h"

extern "C" void mbsv_internal_set_error(const char* msg);

namespace {

int fail(int code, const std::string& msg) {
    mbsv_internal_set_error(msg.c_str());
    return code;
}

struct Nccl {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string why;
};

Nccl* nccl() {
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        // (1) an NCCL the host process already loaded (e.g. PyTorch's bundled one): the loader would hand it to every later
        //     user of that SONAME anyway; (2) MBSV_NCCL_LIB = path of the library to use; (3) the system's libnccl.so.2.
        //     A host that links its own NCCL later than this call (import torch AFTER the first mbsv_comm_* call) would get
        //     ours under the same SONAME: such hosts load theirs first.
        n.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!n.handle)
            if (const char* path = std::getenv("MBSV_NCCL_LIB")) n.handle = dlopen(path, RTLD_NOW | RTLD_LOCAL);
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            if (n.handle) break;
            n.handle = dlopen(name, RTLD_NOW | RTLD_LOCAL);
        }
        if (!n.handle) { n.why = std::string("dlopen(libnccl.so.2): ") + (dlerror() ? dlerror() : "not found"); return; }
        auto sym = [&](const char* s) { void* p = dlsym(n.handle, s); if (!p && n.why.empty()) n.why = std::string("missing NCCL symbol ") + s; return p; };
        n.GetUniqueId = (decltype(n.GetUniqueId))sym("ncclGetUniqueId");
        n.CommInitRank = (decltype(n.CommInitRank))sym("ncclCommInitRank");
        n.CommInitAll = (decltype(n.CommInitAll))sym("ncclCommInitAll");
        n.CommDestroy = (decltype(n.CommDestroy))sym("ncclCommDestroy");
        n.Reduce = (decltype(n.Reduce))sym("ncclReduce");
        n.GroupStart = (decltype(n.GroupStart))sym("ncclGroupStart");
        n.GroupEnd = (decltype(n.GroupEnd))sym("ncclGroupEnd");
        n.GetErrorString = (decltype(n.GetErrorString))sym("ncclGetErrorString");
    });
    return n.why.empty() ? &n : nullptr;
}

#define NCCL_TRY(expr)                                                                                        \
    do {                                                                                                      \
        ncclResult_t _r = (expr);                                                                             \
        if (_r != ncclSuccess) return fail(MBSV_ERR_NCCL, std::string(#expr) + ": " + N->GetErrorString(_r)); \
    } while (0)
#define CUDA_TRY(expr)                                                                                        \
    do {                                                                                                      \
        cudaError_t _e = (expr);                                                                              \
        if (_e != cudaSuccess) return fail(_e == cudaErrorMemoryAllocation ? MBSV_ERR_OOM : MBSV_ERR_CUDA,    \
                                           std::string(#expr) + ": " + cudaGetErrorString(_e));               \
    } while (0)

int need_nccl(Nccl** out) {
    *out = nccl();
    if (!*out) {
        static Nccl* probe = nullptr; (void)probe;
        return fail(MBSV_ERR_NCCL, "NCCL is not available");
    }
    return 0;
}

}  // namespace

struct mbsv_comm {
    ncclComm_t comm = nullptr;
    int n_ranks = 0, rank = 0, device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

extern "C" {

int mbsv_comm_unique_id(uint8_t* out128) {
    if (!out128) return fail(MBSV_ERR_ARG, "mbsv_comm_unique_id: NULL");
    Nccl* N;
    if (int rc = need_nccl(&N)) return rc;
    ncclUniqueId id;
    NCCL_TRY(N->GetUniqueId(&id));
    static_assert(sizeof(id) == MBSV_COMM_ID_BYTES, "ncclUniqueId size");
    std::memcpy(out128, &id, sizeof(id));
    return MBSV_OK;
}

int mbsv_comm_create(const uint8_t* id128, int32_t n_ranks, int32_t rank, int32_t device, mbsv_comm** out) {
    if (!id128 || !out || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(MBSV_ERR_ARG, "mbsv_comm_create: bad argument");
    *out = nullptr;
    Nccl* N;
    if (int rc = need_nccl(&N)) return rc;
    CUDA_TRY(cudaSetDevice(device));
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    mbsv_comm* c = new (std::nothrow) mbsv_comm();
    if (!c) return fail(MBSV_ERR_OOM, "host allocation failed");
    c->n_ranks = n_ranks; c->rank = rank; c->device = device;
    ncclResult_t r = N->CommInitRank(&c->comm, n_ranks, id, rank);
    if (r != ncclSuccess) { delete c; return fail(MBSV_ERR_NCCL, std::string("ncclCommInitRank: ") + N->GetErrorString(r)); }
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
    if (e != cudaSuccess) { N->CommDestroy(c->comm); delete c; return fail(MBSV_ERR_CUDA, cudaGetErrorString(e)); }
    *out = c;
    return MBSV_OK;
}

int mbsv_comm_destroy(mbsv_comm* c) {
    if (!c) return MBSV_OK;
    cudaSetDevice(c->device);
    if (Nccl* N = nccl()) if (c->comm) N->CommDestroy(c->comm);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    delete c;
    return MBSV_OK;
}

int mbsv_reduce_tallies(mbsv_comm* c, uint32_t* tallies_dev, int64_t n_words, int32_t root, double* reduce_ms) {
    if (!c || !tallies_dev || n_words < 0 || root < 0 || root >= c->n_ranks) return fail(MBSV_ERR_ARG, "mbsv_reduce_tallies: bad argument");
    Nccl* N;
    if (int rc = need_nccl(&N)) return rc;
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
    // u32 sums wrap exactly like the difference tallies do inside one GPU, so pooling over GPUs is the same arithmetic
    NCCL_TRY(N->Reduce(tallies_dev, tallies_dev, (size_t)n_words, ncclUint32, ncclSum, root, c->comm, c->stream));
    CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    if (reduce_ms) *reduce_ms = ms;
    return MBSV_OK;
}

int mbsv_split_chains(int32_t n_chains, int32_t n_parts, int32_t part, int32_t* first, int32_t* count) {
    if (n_chains < 0 || n_parts < 1 || part < 0 || part >= n_parts || !first || !count) return fail(MBSV_ERR_ARG, "mbsv_split_chains: bad argument");
    const int base = n_chains / n_parts, extra = n_chains % n_parts;
    *first = part * base + (part < extra ? part : extra);
    *count = base + (part < extra ? 1 : 0);
    return MBSV_OK;
}

int mbsv_run_multi(mbsv_problem* const* problems, int32_t n_dev, const mbsv_run_params* params, mbsv_results* results,
                   double* reduce_ms) {
    if (!problems || n_dev < 1 || !params || !results) return fail(MBSV_ERR_ARG, "mbsv_run_multi: bad argument");
    if (results->final_assign || results->final_logprob || results->counters || results->rng_state || results->error ||
        results->trace_logprob || results->trace_move_frag || results->trace_move_assign || results->initial_assign ||
        results->initial_logprob || params->trace)
        return fail(MBSV_ERR_ARG, "mbsv_run_multi pools the tallies and totals only: per-chain result arrays must be NULL");
    if (params->n_chains < n_dev) return fail(MBSV_ERR_ARG, "mbsv_run_multi: fewer chains than devices");
    Nccl* N;
    if (int rc = need_nccl(&N)) return rc;
    std::vector<int64_t> st(9 * n_dev);
    std::vector<int> devs(n_dev);
    for (int d = 0; d < n_dev; ++d) {
        if (!problems[d]) return fail(MBSV_ERR_ARG, "mbsv_run_multi: NULL problem");
        if (int rc = mbsv_problem_stats(problems[d], &st[9 * d], 9)) return rc;
        devs[d] = (int)st[9 * d + 8];
        if (st[9 * d + 1] != st[1] || st[9 * d + 2] != st[2] || st[9 * d + 7] != st[7])
            return fail(MBSV_ERR_ARG, "mbsv_run_multi: every device must hold the same problem");
        for (int e = 0; e < d; ++e) if (devs[e] == devs[d]) return fail(MBSV_ERR_ARG, "mbsv_run_multi: one problem per device");
    }
    const int64_t n_frag = st[1], n_aln = st[2], n_hist = st[7];
    const int64_t words = n_aln + n_frag + n_hist;
    // communicators of this device set (created once per process)
    static std::mutex mu;
    static std::map<std::vector<int>, std::vector<ncclComm_t>> cache;
    std::vector<ncclComm_t> comms;
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = cache.find(devs);
        if (it == cache.end()) {
            std::vector<ncclComm_t> cs(n_dev);
            NCCL_TRY(N->CommInitAll(cs.data(), n_dev, devs.data()));
            it = cache.emplace(devs, cs).first;
        }
        comms = it->second;
    }
    std::vector<uint32_t*> buf(n_dev, nullptr);
    std::vector<int64_t*> tot(n_dev, nullptr);
    std::vector<cudaStream_t> stream(n_dev, nullptr);
    std::vector<int> rc(n_dev, 0);
    std::vector<std::string> msg(n_dev);
    std::vector<double> kms(n_dev, 0.0);
    auto cleanup = [&]() {
        for (int d = 0; d < n_dev; ++d) {
            cudaSetDevice(devs[d]);
            if (buf[d]) cudaFree(buf[d]);
            if (tot[d]) cudaFree(tot[d]);
            if (stream[d]) cudaStreamDestroy(stream[d]);
        }
    };
    // every device runs its share of the chains (chain c keeps the seed params->seed + c wherever it runs)
    std::vector<std::thread> th;
    for (int d = 0; d < n_dev; ++d)
        th.emplace_back([&, d] {
            auto bad = [&](int code, const std::string& m) { rc[d] = code; msg[d] = m; };
            if (cudaSetDevice(devs[d]) != cudaSuccess) return bad(MBSV_ERR_CUDA, "cudaSetDevice");
            if (cudaMalloc((void**)&buf[d], sizeof(uint32_t) * (size_t)(words > 0 ? words : 1)) != cudaSuccess ||
                cudaMalloc((void**)&tot[d], sizeof(int64_t) * 6) != cudaSuccess ||
                cudaStreamCreateWithFlags(&stream[d], cudaStreamNonBlocking) != cudaSuccess)
                return bad(MBSV_ERR_OOM, "mbsv_run_multi: device buffers");
            int32_t first = 0, count = 0;
            mbsv_split_chains(params->n_chains, n_dev, d, &first, &count);
            mbsv_run_params rp = *params;
            rp.n_chains = count; rp.seed = params->seed + first; rp.device = devs[d];
            mbsv_results r;
            std::memset(&r, 0, sizeof(r));
            r.aln_count = buf[d]; r.frag_err_count = buf[d] + n_aln; r.cluster_hist = buf[d] + n_aln + n_frag;
            r.totals = tot[d];
            const int s = mbsv_run_device(problems[d], &rp, &r);
            kms[d] = r.kernel_ms;
            if (s != MBSV_OK) bad(s, mbsv_last_error());
        });
    for (auto& t : th) t.join();
    for (int d = 0; d < n_dev; ++d) if (rc[d] && rc[d] != MBSV_ERR_INVARIANT) { cleanup(); return fail(rc[d], msg[d]); }
    // ONE reduce of the concatenated integer tallies onto device 0
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaSetDevice(devs[0]);
    cudaEventCreate(&ev0); cudaEventCreate(&ev1);
    cudaEventRecord(ev0, stream[0]);
    ncclResult_t nr = N->GroupStart();
    for (int d = 0; d < n_dev && nr == ncclSuccess; ++d)
        nr = N->Reduce(buf[d], buf[d], (size_t)words, ncclUint32, ncclSum, 0, comms[d], stream[d]);
    if (nr == ncclSuccess) nr = N->GroupEnd(); else N->GroupEnd();
    cudaSetDevice(devs[0]);
    cudaEventRecord(ev1, stream[0]);
    cudaError_t ce = cudaSuccess;
    for (int d = 0; d < n_dev; ++d) { cudaSetDevice(devs[d]); cudaError_t e = cudaStreamSynchronize(stream[d]); if (ce == cudaSuccess) ce = e; }
    float ms = 0;
    cudaSetDevice(devs[0]);
    cudaEventElapsedTime(&ms, ev0, ev1);
    cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    if (nr != ncclSuccess) { cleanup(); return fail(MBSV_ERR_NCCL, std::string("ncclReduce: ") + N->GetErrorString(nr)); }
    if (ce != cudaSuccess) { cleanup(); return fail(MBSV_ERR_CUDA, cudaGetErrorString(ce)); }
    if (reduce_ms) *reduce_ms = ms;
    // pooled tallies and totals to the host
    auto down = [&](void* dst, const void* src, size_t bytes) { if (dst && bytes) ce = cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost); };
    down(results->aln_count, buf[0], sizeof(uint32_t) * n_aln);
    down(results->frag_err_count, buf[0] + n_aln, sizeof(uint32_t) * n_frag);
    down(results->cluster_hist, buf[0] + n_aln + n_frag, sizeof(uint32_t) * n_hist);
    int64_t total[6] = {0, 0, 0, 0, 0, 0};
    double kmax = 0;
    for (int d = 0; d < n_dev; ++d) {
        int64_t t[6];
        cudaSetDevice(devs[d]);
        cudaMemcpy(t, tot[d], sizeof(t), cudaMemcpyDeviceToHost);
        for (int i = 0; i < 5; ++i) total[i] += t[i];
        if (!total[5]) total[5] = t[5];
        kmax = kms[d] > kmax ? kms[d] : kmax;
    }
    if (results->totals) std::memcpy(results->totals, total, sizeof(total));
    results->kernel_ms = kmax;
    cleanup();
    if (ce != cudaSuccess) return fail(MBSV_ERR_CUDA, cudaGetErrorString(ce));
    if (total[5]) return fail((int)total[5], "a chain hit an invariant the reference turns into an Error");
    return MBSV_OK;
}

}  // extern "C"
