/*
 * mbsv_host.h — host-side mirror of the reference driver around the sampler (same library, libmbsv_b200.so).
 *
 * The reference's driver is Java (class MultiBreakSV, src/MultiBreakSV.java); where no JVM exists this C++ mirror
 * plays its part so that the product runs file -> file.  Function names follow the reference's methods:
 *
 *   mbsv_host_get_fragment_alignments   getFragmentAlignments  :551-602  (readExperimentFile :319, readAssignmentFile
 *                                       :333, readClustersFile :402, intersectClustersAndFragments :475)
 *   mbsv_host_construct_cluster_diagrams constructClusterDiagrams :658-739 (+ MultiBreakSVDiagram.java:56-139)
 *   mbsv_host_precompute_priors         precomputePriors :604-656   (logprobcov tables)
 *   mbsv_host_fill_non_singleton_array  fillNonSingletonArray :748-771
 *   mbsv_host_matchfile_initial         initializeMatrix's --matchfile branch :1030-1085
 *   mbsv_host_run_mcmc                  runMCMC :773-877             (-> mbsv_problem_create + mbsv_run)
 *   mbsv_host_write_final_files         writeFinalFiles :1712-1859   (.finalbreakpoints.longburnin, .finalassignment.longburnin)
 *   mbsv_host_write_mcmc_file           the per-iteration mcmc.txt of runMCMC :779-828
 *   mbsv_host_write_state_counts        writeStateCounts :1934-1949  (.samplingprobs)
 *
 * Java semantics reproduced: Scanner tokenisation, String.split/trim, Java 8+ HashMap<String,..> iteration order
 * (fragments, breakpoints, supports, experiments), Double.toString.
 *
 * A host object holds ONE problem (what one JVM run of the reference sees).  mbsv_host_batch_* concatenates several
 * hosts (e.g. one per subset-N.clusters file of generateIndependentSubproblems.pl) into one mbsv_problem_desc;
 * mbsv_host_load_partitioned does partition + per-subset loading in memory and in parallel, reading each file once.
 */
#ifndef MBSV_HOST_H
#define MBSV_HOST_H

#include <stdint.h>

#include "mbsv_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mbsv_host mbsv_host;
typedef struct mbsv_host_batch mbsv_host_batch;

int mbsv_host_create(mbsv_host** out);
int mbsv_host_destroy(mbsv_host* h);

/* --clusterfile --assignmentfile --experimentfile [--readclustersfirst]  (MultiBreakSV.java:211-267) */
int mbsv_host_set_files(mbsv_host* h, const char* clusterfile, const char* assignmentfile, const char* experimentfile,
                        int32_t clusters_first);
int mbsv_host_get_fragment_alignments(mbsv_host* h);
int mbsv_host_construct_cluster_diagrams(mbsv_host* h);
int mbsv_host_precompute_priors(mbsv_host* h);
int mbsv_host_fill_non_singleton_array(mbsv_host* h);
/* all four steps above, in main()'s order (MultiBreakSV.java:119-128) */
int mbsv_host_load(mbsv_host* h);

/* The packed problem of this host (pointers stay valid until the host is destroyed or reloaded). */
int mbsv_host_desc(mbsv_host* h, mbsv_problem_desc* out);
/* Names: kind 0 fragment id (fragments.keySet() order), 1 cluster id (breakpoints.keySet() order), 2 experiment label,
   3 ESP of alignment-ESP entry i, 4 warning text (i ignored), 5 reason the input is rejected (empty = accepted). */
const char* mbsv_host_name(mbsv_host* h, int32_t kind, int32_t i);
/* fragment indices in the sorted order writeFinalFiles uses (getFragmentIDs :1703-1709) */
int mbsv_host_sorted_fragments(mbsv_host* h, int32_t* out);

/* runMCMC: one chain, seed 123456, burnin/skip derived from num_iters as the reference does (:186-193) unless
   params is given.  Results are kept inside the host for the writers; `results` (may be NULL) receives a view. */
int mbsv_host_run_mcmc(mbsv_host* h, int32_t num_iters, double perr, int32_t mode, int32_t device, int32_t save_space,
                       const int32_t* init_assign, mbsv_results* results);
/* After mbsv_host_run_mcmc: counters[5] = {naive proposals, naive accepted, AlterSV proposals, AlterSV accepted,
   fragment reassignments} (proposalCounts/moveCounts, MultiBreakSV.java:842-843) and the initial / final logprob (:839). */
int mbsv_host_run_summary(mbsv_host* h, int64_t* counters, double* initial_logprob, double* final_logprob);
/* --matchfile (MultiBreakSV.java:1030-1085): initial assignment per fragment (keySet() order, n_frag entries) from a
   file of ESPs; counts (may be NULL) = {fragments set, set to error, with a listed ESP but not set, ESPs read}. */
int mbsv_host_matchfile_initial(mbsv_host* h, const char* matchfile, int32_t* init_assign, int32_t* counts);
/* Inject results computed elsewhere (tests without a GPU; batched runs): counts are copied. */
int mbsv_host_set_results(mbsv_host* h, int32_t num_iters, int32_t burnin, const uint32_t* aln_count,
                          const uint32_t* frag_err_count, const uint32_t* cluster_hist);

int mbsv_host_write_final_files(mbsv_host* h, const char* prefix);
int mbsv_host_write_mcmc_file(mbsv_host* h, const char* path);
int mbsv_host_write_state_counts(mbsv_host* h, const char* path);

/* Batch: S hosts -> one descriptor (subproblem s = host s). */
int mbsv_host_batch_create(mbsv_host* const* hosts, int32_t n, mbsv_host_batch** out);
int mbsv_host_batch_desc(mbsv_host_batch* b, mbsv_problem_desc* out);
int mbsv_host_batch_destroy(mbsv_host_batch* b);
/* writeFinalFiles (MultiBreakSV.java:1712-1859) for a batch whose names were kept: <prefix>.finalbreakpoints.longburnin and
   <prefix>.finalassignment.longburnin with one header and the rows of every subproblem in batch order, i.e. the
   concatenation of the per-subset outputs of the reference workflow (rows are ragged across subproblems: each prints
   its own maxSupport columns).  The tallies are those of mbsv_run on the batch descriptor; burnin = the window length. */
int mbsv_host_batch_write_final_files(mbsv_host_batch* b, int32_t burnin, const uint32_t* aln_count,
                                      const uint32_t* frag_err_count, const uint32_t* cluster_hist, const char* prefix);
/* out[i] = subset number (mbsv_host_load_partitioned) or running index (mbsv_host_batch_create) of subproblem i */
int mbsv_host_batch_numbers(mbsv_host_batch* b, int32_t* out);

/* Partition + load in one call, in memory: the batch holds one subproblem per subset of generateIndependentSubproblems.pl,
   each packed exactly as a reference run on `subset-k.clusters --readclustersfirst` with the full assignments file would
   see it (doc/MultiBreakSV-Manual.txt:464-473) -- but the assignments file is read once, not once per subset, and no
   subset file is written.  Subsets without any fragment are skipped.  flags bit 0: a subset whose HashMap order cannot be
   reproduced (a bin Java would treeify, see MBSV_ERR_UNSUPPORTED) keeps the plain-list order instead of failing the load
   -- its chain is valid but not the trajectory a JVM would produce; bit 1: keep the names (fragment, cluster, ESP) for
   mbsv_host_batch_write_final_files.  info (may be NULL) = {skipped subsets, subsets with
   such an order}.  Returns the number of subproblems (> 0) or a negative status. */
int mbsv_host_load_partitioned(const char* clusterfile, const char* assignmentfile, const char* experimentfile,
                               int32_t flags, mbsv_host_batch** out, int32_t* info);

/* generateIndependentSubproblems.pl (src/generateIndependentSubproblems.pl): connected components of the bipartite graph
   long-read id (`longread_(\d+)` inside each ESP, :58) <-> cluster id over the non-dotted rows of a clusters file.
   Subset k (1-based) is the component holding the numerically smallest long-read id not in subsets 1..k-1 (:85,107-113).
   Writes <outdir>/subset-k.clusters (full lines, file order; dotted maximal-cluster rows are dropped exactly as the
   script drops them) when outdir is not NULL; row_subset[i] (may be NULL, length max_rows) receives the subset of
   physical line i of the file or 0 for dropped lines.  Returns the number of subsets or a negative status.
   Union-find: O(rows * alpha) instead of the script's O(|subset|^2) scans (:130-141). */
int mbsv_host_partition(const char* clusterfile, const char* outdir, int32_t* row_subset, int32_t max_rows);

/* Java's Double.toString (shortest repr, JDK 19+ behaviour) into buf; returns the length. */
int mbsv_java_double_to_string(double v, char* buf, int32_t buflen);
/* keySet() order of a java.util.HashMap<String,?> after inserting the n NUL-separated keys (test hook). */
int mbsv_java_hashmap_order(const char* keys, int32_t n, int32_t* out_order, int32_t* out_capacity);
/* keySet() order of a java.util.HashMap<String,?> after a sequence of puts and removes, tree bins included (JDK 8+:
   treeifyBin / TreeNode.putTreeVal / removeTreeNode / split / moveRootToFront, csrc/java_hashmap.hpp).  keys = n
   NUL-terminated strings; ops[k] = id + 1 puts key id (at most once), -(id + 1) removes it.  flags bit 0: always the
   literal emulation instead of the closed form.  out_info[0] = 1 if some bin is treeified at the end (literal) or ever
   was (closed form), out_info[1] = final capacity.  Returns the number of ids written to out_order. */
int mbsv_java_hashmap_replay(const char* keys, int32_t n, const int32_t* ops, int32_t n_ops, int32_t flags,
                             int32_t* out_order, int32_t* out_info);

#ifdef __cplusplus
}
#endif
#endif /* MBSV_HOST_H */
