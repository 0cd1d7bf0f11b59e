// ORACLE — TEST INFRASTRUCTURE ONLY (see jsemantics.hpp header).
//
// Restatement of the reference's input side, i.e. everything main() does before runMCMC:
//   MultiBreakSV.java:319-331  readExperimentFile
//   MultiBreakSV.java:333-400  readAssignmentFile
//   MultiBreakSV.java:402-473  readClustersFile
//   MultiBreakSV.java:475-550  intersectClustersAndFragments
//   MultiBreakSV.java:551-602  getFragmentAlignments (ordering of the above, --readclustersfirst)
//   MultiBreakSV.java:658-746  constructClusterDiagrams  (+ MultiBreakSVDiagram.java:56-139)
//   MultiBreakSV.java:604-656  precomputePriors (logprobcov table; prior[] == 0)
//   MultiBreakSV.java:748-771  fillNonSingletonArray
// and a packing of the resulting object graph into flat index arrays ("Packed"), with the
// HashMap iteration orders the sampler depends on made explicit.
#pragma once
#include <algorithm>
#include <cstdio>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "jsemantics.hpp"

namespace orc {

struct RefAlignment {
    std::vector<std::string> esps;   // FragmentAlignment.alignments
    int numerrors = 0, len = 0;
    int exp_node = -1;               // node id in the experiments map
};
struct RefFragment {
    std::string id;
    std::vector<RefAlignment> alns;  // Fragment.fragmentalignments
};
struct RefDiagram {
    std::string cid;
    bool is_cluster = true;
    std::vector<std::string> leaves;               // leaves.keySet(), insertion order (order is immaterial)
    std::vector<std::vector<std::string>> roots;   // DiagramNode.esps of each root, file order
};

// Flat form of one problem (what one JVM run of the reference sees).  All indices are local.
struct Packed {
    int F = 0, A = 0, K = 0, C = 0, E = 0, P = 0, max_support = 0;
    // fragments, in fragments.keySet() iteration order
    std::vector<std::string> frag_id;
    std::vector<int> frag_aln_ptr;                 // [F+1]
    std::vector<int> sorted_frag;                  // fragment indices in Collections.sort order (:1703-1709)
    // alignments (index within a fragment == AssignmentIndex after the intersect filter, quirk Q8)
    std::vector<int> aln_numerrors, aln_len, aln_exp;
    std::vector<int> aln_esp_ptr;                  // [A+1]
    std::vector<std::string> aln_esp_name;         // [K]
    std::vector<int> aln_esp_id;                   // [K] id of the ESP string
    std::vector<int> aln_esp_rawcluster;           // [K] cluster index of esp2cid.get(esp)
    std::vector<int> aln_esp_cluster;              // [K] same, or -1 when selecting this entry adds nothing
    std::vector<int> aln_esp_leaf;                 // [K] leaf slot inside that cluster, or -1
    // clusters, in breakpoints.keySet() iteration order
    std::vector<std::string> cluster_id;
    std::vector<uint8_t> cluster_kind;             // 0 = simple cluster, 1 = connected component
    std::vector<int> supports_order;               // cluster indices in A.supports.keySet() order
    std::vector<int> cluster_leaf_ptr;             // [C+1]
    std::vector<int> leaf_esp_id;                  // per leaf slot
    std::vector<int> leaf_owner;                   // per leaf slot: fragment index of esp2fragmentID, or -1
    std::vector<int> cluster_root_ptr;             // [C+1]
    std::vector<int> root_leaf_ptr;                // [R+1]
    std::vector<int> root_leaf;                    // leaf slot (local to the cluster) of each root ESP
    // AlterSV (bpsWithNonSingletonSupport) in construction order
    std::vector<int> sv_cluster;
    std::vector<int> sv_frag_ptr;                  // [nsv+1]
    std::vector<int> sv_frag;                      // distinct single-alignment fragments toggled by the move
    std::vector<int> sv_numchanged;                // numchanged as the reference counts it (:1311)
    // experiments, in experiments.keySet() iteration order
    std::vector<std::string> exp_label;
    std::vector<double> exp_pseq, exp_lambda;
    std::vector<double> logprobcov;                // [E][max_support+1], [.][0] = 0.0
    std::vector<std::string> esp_names;            // [P] id -> ESP string
    // diagnostics
    bool tree_bin = false;                         // a HashMap bin would have been treeified (order unpinned)
    std::string unsupported;                       // non-empty: input the reference itself mishandles
    std::vector<int> hashmap_caps;                 // capacities {fragments, breakpoints, supports, experiments}
};

struct RefModel {
    // inputs
    std::string clusterfile, assignmentfile, experimentfile;
    bool clusters_first = false;
    // parsed state (names follow the reference's fields)
    JHashOrder experiments;
    std::vector<double> exp_pseq, exp_lambda;            // by experiments node id
    JHashOrder fragments;
    std::vector<RefFragment> frag_by_node;               // by fragments node id
    std::unordered_map<std::string, int> esp2fragment;   // esp -> fragments node id (last writer wins)
    std::unordered_map<std::string, std::string> esp2cid;
    JHashOrder breakpoints;
    std::vector<char> bp_is_cluster;                     // by breakpoints node id
    int max_support = 0;
    std::unordered_map<std::string, RefDiagram> diagrams;
    std::vector<std::string> diagram_order;              // construction order (immaterial)
    std::vector<std::string> bps_with_nonsingleton_support;
    std::string warning;

    static std::string slurp(const std::string& path) {
        std::ifstream in(path, std::ios::binary);
        if (!in) throw std::runtime_error("cannot open " + path);
        std::stringstream ss;
        ss << in.rdbuf();
        return ss.str();
    }
    static int parse_int(const std::string& s) {
        // Integer.parseInt / Scanner.nextInt for plain decimal integers.
        if (s.empty()) throw std::runtime_error("NumberFormatException: empty");
        size_t i = 0;
        bool neg = false;
        if (s[0] == '-' || s[0] == '+') { neg = s[0] == '-'; i = 1; }
        if (i >= s.size()) throw std::runtime_error("NumberFormatException: " + s);
        long long v = 0;
        for (; i < s.size(); ++i) {
            if (s[i] < '0' || s[i] > '9') throw std::runtime_error("NumberFormatException: " + s);
            v = v * 10 + (s[i] - '0');
            if (v > 2147483648LL) throw std::runtime_error("NumberFormatException: " + s);
        }
        v = neg ? -v : v;
        if (v > 2147483647LL || v < -2147483648LL) throw std::runtime_error("NumberFormatException: " + s);
        return (int)v;
    }
    static double parse_double(const std::string& s) {
        char* end = nullptr;
        double v = std::strtod(s.c_str(), &end);
        if (end == s.c_str() || *end != 0) throw std::runtime_error("InputMismatchException: " + s);
        return v;
    }

    // MultiBreakSV.java:319-331
    void readExperimentFile() {
        JScanner scan{slurp(experimentfile)};
        if (!scan.hasNextLine()) throw std::runtime_error("NoSuchElementException: empty experiment file");
        scan.nextLine();
        while (scan.hasNext()) {
            std::string label = scan.next();
            if (!scan.hasNext()) throw std::runtime_error("NoSuchElementException in experiment file");
            double p = parse_double(scan.next());
            if (!scan.hasNext()) throw std::runtime_error("NoSuchElementException in experiment file");
            double l = parse_double(scan.next());
            int id = experiments.put(label);
            if ((int)exp_pseq.size() <= id) { exp_pseq.resize(id + 1); exp_lambda.resize(id + 1); }
            exp_pseq[id] = p;
            exp_lambda[id] = l;
        }
    }

    // MultiBreakSV.java:333-400
    void readAssignmentFile(std::unordered_set<std::string>& allesps) {
        JScanner scan{slurp(assignmentfile)};
        if (!scan.hasNextLine()) throw std::runtime_error("NoSuchElementException: empty assignment file");
        scan.nextLine();
        while (scan.hasNext()) {
            std::string fragmentID = scan.next();
            if (!scan.hasNext()) throw std::runtime_error("Error parsing line. Check assignmentfile format.");
            std::vector<std::string> esps = java_split(scan.next(), ",");
            if (!scan.hasNext()) throw std::runtime_error("Error parsing line. Check assignmentfile format.");
            int numerrors = parse_int(scan.next());
            if (!scan.hasNext()) throw std::runtime_error("Error parsing line. Check assignmentfile format.");
            int length = parse_int(scan.next());
            if (!scan.hasNext()) throw std::runtime_error("Error parsing line. Check assignmentfile format.");
            std::string explabel = scan.next();
            if (!experiments.contains(explabel))
                throw std::runtime_error("ERROR: Experiment Label " + explabel + " not in Experiments file.");
            if (clusters_first) {
                bool espInClust = false;
                for (const auto& e : esps) if (allesps.count(e)) { espInClust = true; break; }
                if (!espInClust) continue;
            }
            int fid = fragments.put(fragmentID);
            if ((int)frag_by_node.size() <= fid) { frag_by_node.resize(fid + 1); frag_by_node[fid].id = fragmentID; }
            RefAlignment fa;
            fa.esps = esps;
            fa.numerrors = numerrors;
            fa.len = length;
            fa.exp_node = experiments.find(explabel);
            frag_by_node[fid].alns.push_back(fa);
            for (const auto& e : esps) { allesps.insert(e); esp2fragment[e] = fid; }
        }
    }

    // MultiBreakSV.java:402-473
    void readClustersFile(std::unordered_set<std::string>& allesps) {
        max_support = 0;
        JHashOrder clust;
        std::vector<char> clust_val;
        JScanner scan{slurp(clusterfile)};
        bool emitwarning = false;
        while (scan.hasNext()) {
            std::vector<std::string> row = java_split(scan.nextLine(), "\t");
            if (row[0] == "cID") continue;
            std::string cID = row[0];
            if (cID.find('.') != std::string::npos) continue;
            if (row.size() < 2) throw std::runtime_error("ArrayIndexOutOfBounds: clusters row has < 2 columns");
            int n = parse_int(row[1]);
            if (max_support < n) max_support = n;
            if (row.size() < 5) throw std::runtime_error("ArrayIndexOutOfBounds: clusters row has < 5 columns");
            std::vector<std::string> esplist = java_split(row[4], ",");
            bool someESP = false;
            for (auto& e : esplist) {
                e = java_trim(e);
                if (!clusters_first && !allesps.count(e)) { emitwarning = true; continue; }
                someESP = true;
                esp2cid[e] = cID;
                if (clusters_first) allesps.insert(e);
            }
            if (!clusters_first && !someESP) continue;
            int id = clust.put(cID);
            if ((int)clust_val.size() <= id) clust_val.resize(id + 1);
            clust_val[id] = (row[2] == "-1") ? 0 : 1;   // false = conn comp, true = cluster
        }
        if (emitwarning) warning = "WARNING! Some fragments in clusters are NOT in ESP file";
        for (int id : clust.order()) {
            int b = breakpoints.put(clust.key(id));
            if ((int)bp_is_cluster.size() <= b) bp_is_cluster.resize(b + 1);
            bp_is_cluster[b] = clust_val[id];
        }
        if (clust.tree_bin) breakpoints.tree_bin = true;
    }

    // MultiBreakSV.java:475-550
    void intersectClustersAndFragments() {
        std::vector<std::string> toRemove;
        for (int fid : fragments.order()) {
            RefFragment& fr = frag_by_node[fid];
            std::vector<RefAlignment> tmp;
            for (const auto& al : fr.alns) {
                size_t tmpnum = 0;
                for (const auto& e : al.esps) if (esp2cid.count(e)) ++tmpnum;
                if (tmpnum == al.esps.size()) tmp.push_back(al);
            }
            if (tmp.empty()) toRemove.push_back(fr.id);
            else fr.alns = tmp;
        }
        for (const auto& f : toRemove) fragments.remove(f);
        // clusters that contain no ESP of any surviving alignment are dropped
        std::unordered_set<std::string> used;
        for (int fid : fragments.order())
            for (const auto& al : frag_by_node[fid].alns)
                for (const auto& e : al.esps) used.insert(esp2cid.at(e));
        std::vector<std::string> bpRemove;
        for (int b : breakpoints.order())
            if (!used.count(breakpoints.key(b))) bpRemove.push_back(breakpoints.key(b));
        for (const auto& b : bpRemove) breakpoints.remove(b);
    }

    // MultiBreakSV.java:551-574
    void getFragmentAlignments() {
        readExperimentFile();
        std::unordered_set<std::string> allesps;
        if (clusters_first) { readClustersFile(allesps); readAssignmentFile(allesps); }
        else { readAssignmentFile(allesps); readClustersFile(allesps); }
        intersectClustersAndFragments();
    }

    // MultiBreakSV.java:658-746 + MultiBreakSVDiagram.java:56-139
    void constructClusterDiagrams() {
        JScanner scan{slurp(clusterfile)};
        auto advance = [&](std::vector<std::string>& row) -> bool {
            if (!scan.hasNext()) return false;
            row = java_split(scan.nextLine(), "\t");
            return true;
        };
        std::vector<std::string> row;
        bool have = advance(row);
        while (have) {
            if (row[0] == "cID") { have = advance(row); continue; }
            if (row.size() < 5) throw std::runtime_error("ArrayIndexOutOfBounds: clusters row has < 5 columns");
            std::string cID = row[0], local = row[2], names = row[4];
            if (!breakpoints.contains(cID)) { have = advance(row); continue; }
            RefDiagram d;
            d.cid = cID;
            auto add_leaf = [&](const std::string& e) {
                if (std::find(d.leaves.begin(), d.leaves.end(), e) == d.leaves.end()) d.leaves.push_back(e);
            };
            if (local != "-1") {
                d.is_cluster = true;
                std::vector<std::string> nm = java_split(names, ", ");
                d.roots.push_back(nm);
                for (const auto& e : nm) add_leaf(e);
                have = advance(row);
            } else {
                d.is_cluster = false;
                have = advance(row);
                while (have && row[0].compare(0, cID.size(), cID) == 0 && row[0].size() >= cID.size()) {
                    if (row.size() < 5) throw std::runtime_error("ArrayIndexOutOfBounds: sub-cluster row");
                    std::vector<std::string> nm = java_split(row[4], ", ");
                    d.roots.push_back(nm);
                    for (const auto& e : nm) add_leaf(e);
                    have = advance(row);
                }
            }
            for (const auto& e : d.leaves) {
                auto it = esp2cid.find(e);
                if (it != esp2cid.end()) it->second = cID;
            }
            if (!diagrams.count(cID)) diagram_order.push_back(cID);
            diagrams[cID] = d;
        }
    }

    // MultiBreakSV.java:748-771
    void fillNonSingletonArray() {
        bps_with_nonsingleton_support.clear();
        for (int b : breakpoints.order()) {
            const std::string& cID = breakpoints.key(b);
            int uniqueCounts = 0;
            for (const auto& leaf : diagrams.at(cID).leaves) {
                auto it = esp2fragment.find(leaf);
                if (it == esp2fragment.end()) continue;
                const RefFragment& fr = frag_by_node[it->second];
                if (fragments.contains(fr.id) && fr.alns.size() == 1) ++uniqueCounts;
            }
            if (uniqueCounts > 1) bps_with_nonsingleton_support.push_back(cID);
        }
    }

    // MultiBreakSV.java:1644-1649
    static double logPoisson(int k, double lambdaD) {
        double val = -lambdaD + k * j_log(lambdaD);
        for (int i = 1; i <= k; ++i) val -= j_log((double)i);
        return val;
    }

    void load() {
        getFragmentAlignments();
        constructClusterDiagrams();
        // every kept breakpoint needs a diagram (diagrams.get(cID) is dereferenced at :758)
        for (int b : breakpoints.order())
            if (!diagrams.count(breakpoints.key(b)))
                throw std::runtime_error("NullPointerException: no diagram for " + breakpoints.key(b));
        fillNonSingletonArray();
    }

    Packed pack() const {
        Packed p;
        p.tree_bin = experiments.tree_bin || fragments.tree_bin || breakpoints.tree_bin;
        // experiments
        std::vector<int> eorder = experiments.order();
        std::unordered_map<int, int> enode2idx;
        p.E = (int)eorder.size();
        for (int i = 0; i < p.E; ++i) {
            enode2idx[eorder[i]] = i;
            p.exp_label.push_back(experiments.key(eorder[i]));
            p.exp_pseq.push_back(exp_pseq[eorder[i]]);
            p.exp_lambda.push_back(exp_lambda[eorder[i]]);
        }
        p.max_support = max_support;
        // MultiBreakSV.java:612-614 (logprobcov[0] stays 0.0, the Java array default)
        p.logprobcov.assign((size_t)p.E * (max_support + 1), 0.0);
        for (int e = 0; e < p.E; ++e)
            for (int k = 1; k <= max_support; ++k)
                p.logprobcov[(size_t)e * (max_support + 1) + k] = logPoisson(k, p.exp_lambda[e]);
        // clusters
        std::vector<int> border = breakpoints.order();
        p.C = (int)border.size();
        std::unordered_map<std::string, int> cid2idx;
        for (int i = 0; i < p.C; ++i) {
            const std::string& cid = breakpoints.key(border[i]);
            cid2idx[cid] = i;
            p.cluster_id.push_back(cid);
            p.cluster_kind.push_back(bp_is_cluster[border[i]] ? 0 : 1);
        }
        {   // A.supports is filled by iterating breakpoints (MultiBreakSVAssignmentMatrix.java:53-66)
            JHashOrder sup;
            for (int i = 0; i < p.C; ++i) sup.put(p.cluster_id[i]);
            for (int id : sup.order()) p.supports_order.push_back(cid2idx.at(sup.key(id)));
            if (sup.tree_bin) p.tree_bin = true;
            p.hashmap_caps = {fragments.capacity(), breakpoints.capacity(), sup.capacity(), experiments.capacity()};
        }
        // fragments
        std::vector<int> forder = fragments.order();
        p.F = (int)forder.size();
        std::unordered_map<int, int> fnode2idx;
        for (int i = 0; i < p.F; ++i) fnode2idx[forder[i]] = i;
        // ESP ids
        std::unordered_map<std::string, int> espid;
        auto esp_id = [&](const std::string& e) {
            auto it = espid.find(e);
            if (it != espid.end()) return it->second;
            int id = (int)p.esp_names.size();
            espid.emplace(e, id);
            p.esp_names.push_back(e);
            return id;
        };
        // diagrams -> leaves / roots
        p.cluster_leaf_ptr.push_back(0);
        p.cluster_root_ptr.push_back(0);
        p.root_leaf_ptr.push_back(0);
        std::unordered_map<int, int> esp_leaf_count;                 // esp id -> number of diagrams it is a leaf of
        std::vector<std::unordered_map<int, int>> leaf_slot(p.C);    // per cluster: esp id -> local leaf slot
        for (int c = 0; c < p.C; ++c) {
            const RefDiagram& d = diagrams.at(p.cluster_id[c]);
            if (d.is_cluster != (p.cluster_kind[c] == 0)) p.unsupported = "cluster kind mismatch for " + d.cid;
            for (const auto& leaf : d.leaves) {
                int id = esp_id(leaf);
                leaf_slot[c][id] = (int)p.leaf_esp_id.size() - p.cluster_leaf_ptr[c];
                p.leaf_esp_id.push_back(id);
                ++esp_leaf_count[id];
                int owner = -1;
                auto it = esp2fragment.find(leaf);
                if (it != esp2fragment.end() && fragments.contains(frag_by_node[it->second].id))
                    owner = fnode2idx.at(it->second);
                p.leaf_owner.push_back(owner);
            }
            p.cluster_leaf_ptr.push_back((int)p.leaf_esp_id.size());
            for (const auto& r : d.roots) {
                std::unordered_set<int> seen;
                for (const auto& e : r) {
                    int id = esp_id(e);
                    if (!seen.insert(id).second)
                        p.unsupported = "ESP " + e + " listed twice in one row of cluster " + d.cid;
                    p.root_leaf.push_back(leaf_slot[c].at(id));
                }
                p.root_leaf_ptr.push_back((int)p.root_leaf.size());
            }
            p.cluster_root_ptr.push_back((int)p.root_leaf_ptr.size() - 1);
            int nleaf = p.cluster_leaf_ptr[c + 1] - p.cluster_leaf_ptr[c];
            if (nleaf > max_support)
                p.unsupported = "cluster " + d.cid + " has more ESPs than its declared count (logprobcov overflow)";
        }
        for (const auto& kv : esp_leaf_count)
            if (kv.second > 1) p.unsupported = "ESP " + p.esp_names[kv.first] + " is a leaf of more than one cluster";
        // alignments
        p.frag_aln_ptr.push_back(0);
        p.aln_esp_ptr.push_back(0);
        for (int i = 0; i < p.F; ++i) {
            const RefFragment& fr = frag_by_node[forder[i]];
            p.frag_id.push_back(fr.id);
            for (const auto& al : fr.alns) {
                p.aln_numerrors.push_back(al.numerrors);
                p.aln_len.push_back(al.len);
                p.aln_exp.push_back(enode2idx.at(al.exp_node));
                std::unordered_set<int> seen;
                for (const auto& e : al.esps) {
                    int id = esp_id(e);
                    if (!seen.insert(id).second)
                        p.unsupported = "ESP " + e + " listed twice in one alignment of " + fr.id;
                    int c = cid2idx.at(esp2cid.at(e));
                    p.aln_esp_name.push_back(e);
                    p.aln_esp_id.push_back(id);
                    p.aln_esp_rawcluster.push_back(c);
                    auto ls = leaf_slot[c].find(id);
                    bool contributes = ls != leaf_slot[c].end() &&
                                       p.leaf_owner[p.cluster_leaf_ptr[c] + ls->second] == i;
                    p.aln_esp_cluster.push_back(contributes ? c : -1);
                    p.aln_esp_leaf.push_back(contributes ? ls->second : -1);
                }
                p.aln_esp_ptr.push_back((int)p.aln_esp_id.size());
            }
            p.frag_aln_ptr.push_back((int)p.aln_numerrors.size());
        }
        p.A = (int)p.aln_numerrors.size();
        p.K = (int)p.aln_esp_id.size();
        p.P = (int)p.esp_names.size();
        // sorted fragment ids (Collections.sort == String.compareTo; byte order for ASCII)
        p.sorted_frag.resize(p.F);
        for (int i = 0; i < p.F; ++i) p.sorted_frag[i] = i;
        std::sort(p.sorted_frag.begin(), p.sorted_frag.end(),
                  [&](int a, int b) { return p.frag_id[a] < p.frag_id[b]; });
        // AlterSV lists (MultiBreakSV.java:1303-1330)
        p.sv_frag_ptr.push_back(0);
        for (const auto& cid : bps_with_nonsingleton_support) {
            int c = cid2idx.at(cid);
            p.sv_cluster.push_back(c);
            int numchanged = 0;
            std::vector<int> uniq;
            for (int l = p.cluster_leaf_ptr[c]; l < p.cluster_leaf_ptr[c + 1]; ++l) {
                int f = p.leaf_owner[l];
                if (f >= 0 && p.frag_aln_ptr[f + 1] - p.frag_aln_ptr[f] == 1) {
                    ++numchanged;
                    if (std::find(uniq.begin(), uniq.end(), f) == uniq.end()) uniq.push_back(f);
                }
            }
            p.sv_numchanged.push_back(numchanged);
            for (int f : uniq) p.sv_frag.push_back(f);
            p.sv_frag_ptr.push_back((int)p.sv_frag.size());
        }
        return p;
    }
};

// Literal MultiBreakSVAssignmentMatrix.setBreakpoints (:264-414) for one cluster, from ESP identities
// (leaves -> owning fragment -> "is this key in the assigned alignment"), used to cross-check the
// count-based support bookkeeping.  Returns supports[exp] (sorted ascending).
static inline std::vector<std::vector<int>> literal_set_breakpoints(const Packed& p, int c, const int* assign) {
    std::vector<std::vector<int>> esps(p.E);
    int numselected = 0;
    for (int l = p.cluster_leaf_ptr[c]; l < p.cluster_leaf_ptr[c + 1]; ++l) {
        int f = p.leaf_owner[l];
        if (f < 0) continue;
        if (assign[f] != -1) {
            int al = p.frag_aln_ptr[f] + assign[f];
            for (int j = p.aln_esp_ptr[al]; j < p.aln_esp_ptr[al + 1]; ++j)
                if (p.aln_esp_id[j] == p.leaf_esp_id[l]) { esps[p.aln_exp[al]].push_back(p.leaf_esp_id[l]); ++numselected; }
        }
    }
    std::vector<std::vector<int>> sup(p.E);
    if (numselected == 0) return sup;
    for (int x = 0; x < p.E; ++x) {
        const std::vector<int>& sel = esps[x];
        std::vector<char> found(sel.size(), 0);
        std::vector<int> maximal;
        for (int r = p.cluster_root_ptr[c]; r < p.cluster_root_ptr[c + 1]; ++r) maximal.push_back(r);
        auto index_of = [&](int id) { for (size_t i = 0; i < sel.size(); ++i) if (sel[i] == id) return (int)i; return -1; };
        bool allfound = false;
        while (!allfound) {
            int maxcount = 0, maxind = -1;
            for (size_t i = 0; i < maximal.size(); ++i) {
                int cur = 0;
                for (int j = p.root_leaf_ptr[maximal[i]]; j < p.root_leaf_ptr[maximal[i] + 1]; ++j) {
                    int id = p.leaf_esp_id[p.cluster_leaf_ptr[c] + p.root_leaf[j]];
                    int ix = index_of(id);
                    if (ix >= 0 && !found[ix]) ++cur;
                }
                if (cur > maxcount) { maxcount = cur; maxind = (int)i; }
            }
            if (maxind == -1) {
                allfound = true;
                for (char f : found) if (!f) allfound = false;
                if (!allfound) throw std::runtime_error("Error: set cover failed to cover");
                continue;
            }
            for (int j = p.root_leaf_ptr[maximal[maxind]]; j < p.root_leaf_ptr[maximal[maxind] + 1]; ++j) {
                int ix = index_of(p.leaf_esp_id[p.cluster_leaf_ptr[c] + p.root_leaf[j]]);
                if (ix >= 0) found[ix] = 1;
            }
            sup[x].push_back(maxcount);
            maximal.erase(maximal.begin() + maxind);
            allfound = true;
            for (char f : found) if (!f) allfound = false;
        }
        std::sort(sup[x].begin(), sup[x].end());
    }
    return sup;
}

}  // namespace orc
