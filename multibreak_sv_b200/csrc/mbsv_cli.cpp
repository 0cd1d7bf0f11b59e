// mbsv_cli -- command-line driver with the reference's options, file -> file, over the C ABI of libmbsv_b200.so.
//
// Plays the part of `java -jar MultiBreakSV.jar ...` (MultiBreakSV.main, src/MultiBreakSV.java:114-153, options
// :211-267) where no JVM exists: parse + intersect + diagrams + priors (host mirror), runMCMC on the GPU, then the
// reference's writers.  Extra options of this driver: --device N, --mode replay|fast (default replay = the
// reference's arithmetic).  --enumerate is outside the sampler path and is refused.
//
// --partition [--chains N]: the whole multi-subproblem workflow of the manual (IV: generateIndependentSubproblems.pl, then
// one run per subset-k.clusters with --readclustersfirst) in one process: partition + load in memory, one GPU batch, and
// the concatenated .final*.longburnin files.  Implies --savespace.
//
// Exit codes: 0 ok; 1 bad usage (the reference prints its usage and exits); 2 the library reported an error.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mbsv_host.h"

namespace {

void usage(const char* why) {
    std::fprintf(stderr,
                 "ERROR: %s\n\n"
                 "Usage: mbsv_cli --clusterfile F --assignmentfile F --experimentfile F --numiterations N --perr P\n"
                 "                [--prefix OUT] [--matchfile F] [--savespace] [--readclustersfirst]\n"
                 "                [--device N] [--mode replay|fast] [--partition [--chains N]]\n",
                 why);
}

int fail_lib(const char* what) {
    std::fprintf(stderr, "%s: %s\n", what, mbsv_last_error());
    return 2;
}

std::string jd(double v) {
    char buf[64];
    mbsv_java_double_to_string(v, buf, sizeof buf);
    return buf;
}

}  // namespace

int main(int argc, char** argv) {
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);   // host setup (include/mbsv_b200.h): before the first CUDA call
    const auto t0 = std::chrono::steady_clock::now();
    std::string clusterfile, assignmentfile, experimentfile, matchfile, prefix = "out";
    long numiters = -1;
    double perr = -1;
    bool save_space = false, clusters_first = false, partition = false;
    int device = 0, mode = MBSV_MODE_REPLAY_EXACT, chains = 1;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto value = [&]() -> const char* { return i + 1 < argc ? argv[++i] : nullptr; };
        const char* v = nullptr;
        if (a == "--savespace") { save_space = true; continue; }
        if (a == "--readclustersfirst") { clusters_first = true; continue; }
        if (a == "--partition") { partition = true; continue; }
        if (a == "--enumerate") { usage("--enumerate is outside the sampler path"); return 1; }
        if (a != "--clusterfile" && a != "--assignmentfile" && a != "--experimentfile" && a != "--matchfile" && a != "--prefix" &&
            a != "--numiterations" && a != "--perr" && a != "--device" && a != "--mode" && a != "--chains") {
            usage((a + " is not a valid option.").c_str());
            return 1;
        }
        if (!(v = value())) { usage((a + " needs a value").c_str()); return 1; }
        char* end = nullptr;
        if (a == "--clusterfile") clusterfile = v;
        else if (a == "--assignmentfile") assignmentfile = v;
        else if (a == "--experimentfile") experimentfile = v;
        else if (a == "--matchfile") matchfile = v;
        else if (a == "--prefix") prefix = v;
        else if (a == "--numiterations") { numiters = std::strtol(v, &end, 10); if (*end || numiters < 0 || numiters > 2147483647L) { usage("--numiterations must be a positive integer."); return 1; } }
        else if (a == "--perr") { perr = std::strtod(v, &end); if (*end) { usage("--perr must be a float between 0 and 1."); return 1; } }
        else if (a == "--device") device = std::atoi(v);
        else if (a == "--chains") { chains = std::atoi(v); if (chains < 1) { usage("--chains must be a positive integer."); return 1; } }
        else if (a == "--mode") {
            if (!std::strcmp(v, "replay")) mode = MBSV_MODE_REPLAY_EXACT;
            else if (!std::strcmp(v, "fast")) mode = MBSV_MODE_FAST;
            else { usage("--mode is replay or fast"); return 1; }
        }
    }
    if (clusterfile.empty()) { usage("Cluster file is required."); return 1; }
    if (assignmentfile.empty()) { usage("Assignment file is required."); return 1; }
    if (experimentfile.empty()) { usage("Experiment file is required."); return 1; }
    if (numiters == -1) { usage("Number of iterations is required to sample."); return 1; }
    if (perr == -1) { usage("Mapping error probability (perr) is required."); return 1; }
    if (perr < 0 || perr > 1) { usage("--perr must be a value between 0 and 1."); return 1; }

    if (partition) {
        if (!matchfile.empty()) { usage("--matchfile cannot be combined with --partition"); return 1; }
        mbsv_host_batch* b = nullptr;
        int32_t info[2] = {0, 0};
        const int S = mbsv_host_load_partitioned(clusterfile.c_str(), assignmentfile.c_str(), experimentfile.c_str(), 2, &b, info);
        if (S < 0) return fail_lib("partition + load");
        mbsv_problem_desc d;
        if (mbsv_host_batch_desc(b, &d)) return fail_lib("batch");
        std::printf("%d independent subproblems (%d empty skipped): %d fragments, %d alignments, %d clusters\n", S, info[0], d.n_frag,
                    d.n_aln, d.n_cluster);
        if (info[1]) std::printf("%d subproblems have a treeified java.util.HashMap bin (red-black order emulated)\n", info[1]);
        mbsv_problem* prob = nullptr;
        if (mbsv_problem_create(&d, device, &prob)) return fail_lib("create");
        mbsv_run_params rp;
        std::memset(&rp, 0, sizeof rp);
        const int burnin = numiters / 10 ? (int)(numiters / 10) : 10;          // :186-193
        rp.mode = mode; rp.n_chains = chains; rp.num_iters = (int32_t)numiters; rp.burnin = burnin; rp.perr = perr; rp.seed = 123456;
        rp.device = device; rp.java_int_wrap = 1; rp.memo_states = 65536;
        std::vector<uint32_t> aln(d.n_aln), fe(d.n_frag), hist(d.n_hist > 0 ? d.n_hist : 1);
        std::vector<int64_t> totals(6, 0);
        mbsv_results r;
        std::memset(&r, 0, sizeof r);
        r.aln_count = aln.data(); r.frag_err_count = fe.data(); r.cluster_hist = hist.data(); r.totals = totals.data();
        if (mbsv_run(prob, &rp, &r)) return fail_lib("runMCMC");
        mbsv_problem_destroy(prob);
        std::printf("%lld of %lld(%s) Naive Moves\n", (long long)totals[1], (long long)totals[0], jd((double)totals[1] / (double)totals[0]).c_str());
        std::printf("%lld of %lld(%s) AlterSV Moves\n", (long long)totals[3], (long long)totals[2], jd((double)totals[3] / (double)totals[2]).c_str());
        // tallies are pooled over the chains: the window length is chains x burnin samples
        if (mbsv_host_batch_write_final_files(b, burnin * chains, aln.data(), fe.data(), hist.data(), prefix.c_str())) return fail_lib("writeFinalFiles");
        mbsv_host_batch_destroy(b);
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::printf("SAMPLING TIME %s\n", jd(secs).c_str());
        return 0;
    }
    mbsv_host* h = nullptr;
    if (mbsv_host_create(&h)) return fail_lib("create");
    if (mbsv_host_set_files(h, clusterfile.c_str(), assignmentfile.c_str(), experimentfile.c_str(), clusters_first)) return fail_lib("files");
    // main(): getFragmentAlignments, constructClusterDiagrams, precomputePriors, fillNonSingletonArray (:119-128)
    if (mbsv_host_get_fragment_alignments(h)) return fail_lib("getFragmentAlignments");
    if (mbsv_host_construct_cluster_diagrams(h)) return fail_lib("constructClusterDiagrams");
    if (mbsv_host_precompute_priors(h)) return fail_lib("precomputePriors");
    if (mbsv_host_fill_non_singleton_array(h)) return fail_lib("fillNonSingletonArray");
    const char* warn = mbsv_host_name(h, 4, 0);
    if (warn && *warn) std::printf("%s\n", warn);
    mbsv_problem_desc d;
    if (mbsv_host_desc(h, &d)) return fail_lib("pack");
    std::printf("%d fragments, %d alignments, %d clusters, %d experiments\n", d.n_frag, d.n_aln, d.n_cluster, d.n_exp);

    std::vector<int32_t> init;
    if (!matchfile.empty()) {
        init.resize(d.n_frag > 0 ? d.n_frag : 1);
        int32_t c[4];
        if (mbsv_host_matchfile_initial(h, matchfile.c_str(), init.data(), c)) return fail_lib("matchfile");
        // the lines initializeMatrix prints (:1032, :1078-1081)
        std::printf("  Set file %s as initial A.\n%d esps recorded.\n%d total fragments\n", matchfile.c_str(), c[3], d.n_frag);
        std::printf("%d fragments set in A, and %d fragments set to error.\n", c[0], c[1]);
        std::printf("%d fragments had an ESP in the matchfile but no assignment was set.\n", c[2]);
    }
    if (mbsv_host_run_mcmc(h, (int32_t)numiters, perr, mode, device, save_space, init.empty() ? nullptr : init.data(), nullptr))
        return fail_lib("runMCMC");
    int64_t cnt[5];
    double lp0 = 0, lp1 = 0;
    if (mbsv_host_run_summary(h, cnt, &lp0, &lp1)) return fail_lib("summary");
    // :839-843
    std::printf("  After %ld iterations, probability is %s\n\n", numiters, jd(lp1).c_str());
    std::printf("%lld of %lld(%s) Naive Moves\n", (long long)cnt[1], (long long)cnt[0], jd((double)cnt[1] / (double)cnt[0]).c_str());
    std::printf("%lld of %lld(%s) AlterSV Moves\n", (long long)cnt[3], (long long)cnt[2], jd((double)cnt[3] / (double)cnt[2]).c_str());
    if (!save_space && mbsv_host_write_mcmc_file(h, (prefix + ".mcmc.txt").c_str())) return fail_lib("mcmc.txt");
    if (mbsv_host_write_final_files(h, prefix.c_str())) return fail_lib("writeFinalFiles");
    if (!save_space && mbsv_host_write_state_counts(h, (prefix + ".samplingprobs").c_str())) return fail_lib("writeStateCounts");
    mbsv_host_destroy(h);
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("SAMPLING TIME %s\n", jd(secs).c_str());
    return 0;
}
