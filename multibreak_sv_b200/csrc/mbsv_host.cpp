// Host-side mirror of the reference driver (include/mbsv_host.h): parses the three input files with the
// reference's semantics, packs them into mbsv_problem_desc, runs the sampler through the C ABI and writes the
// reference's output files.  No sampling happens here: mbsv_host_run_mcmc calls mbsv_problem_create / mbsv_run.
#include "../../include/mbsv_host.h"
#include "java_hashmap.hpp"

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <chrono>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace {

}  // namespace
extern "C" void mbsv_internal_set_error(const char* msg);   // mbsv_core.cu
namespace {
int hfail(int code, const std::string& msg) { mbsv_internal_set_error(msg.c_str()); return code; }

// ---------------------------------------------------------------------------------------------------------
// Java semantics
// ---------------------------------------------------------------------------------------------------------
uint32_t java_hash(std::string_view s) {               // String.hashCode then HashMap.hash spreading
    uint32_t h = 0;
    for (unsigned char c : s) h = 31u * h + c;
    return h ^ (h >> 16);
}

// Iteration order of a java.util.HashMap<String,?> (Java 8+) without materialising buckets: within a bucket
// entries stay in insertion order (tail insertion; resize splits preserve order; removal unlinks), so the
// keySet() order is "by bucket index at the final capacity, then by insertion".  The capacity is tracked through
// every put: doubling when size exceeds 0.75 capacity, and treeifyBin()'s extra doubling when a bin receives its
// 9th entry while the table is smaller than 64.  A bin that is treeified at capacity >= 64 reorders its `next` links
// (red-black tree, root moved to the front): `tree_bin` is raised and order() replays the map's puts and removes through
// the literal emulation of java_hashmap.hpp instead of the closed form.
struct JavaMapOrder {
    std::vector<uint32_t> hash;
    std::vector<char> alive;
    std::vector<int> bin;     // live entries per bucket
    std::vector<std::string> keys;   // (replay of a treeified map)
    std::vector<int> ops;            // id + 1 = put, -(id + 1) = remove
    int cap = 0, thr = 0, size = 0;
    bool tree_bin = false;
    void resize() {
        if (cap == 0) { cap = 16; thr = 12; }
        else { cap *= 2; thr *= 2; }
        bin.assign(cap, 0);
        for (size_t i = 0; i < hash.size(); ++i) if (alive[i]) ++bin[hash[i] & (cap - 1)];
    }
    int put(std::string_view key) {                   // key must be new
        if (cap == 0) resize();
        const uint32_t h = java_hash(key);
        hash.push_back(h);
        alive.push_back(1);
        keys.emplace_back(key);
        ops.push_back((int)hash.size());
        const int before = bin[h & (cap - 1)]++;
        if (before >= 8) { if (cap < 64) resize(); else tree_bin = true; }
        if (++size > thr) resize();
        return (int)hash.size() - 1;
    }
    void remove(int id) {
        if (!alive[id]) return;
        alive[id] = 0;
        --bin[hash[id] & (cap - 1)];
        --size;
        ops.push_back(-(id + 1));
    }
    std::vector<int> order() const {
        if (tree_bin) {
            mbsv_java::HashMapLiteral m;
            for (int op : ops) { if (op > 0) m.put(keys[op - 1]); else m.remove(-op - 1); }
            return m.order();
        }
        std::vector<int> o;
        for (size_t i = 0; i < hash.size(); ++i) if (alive[i]) o.push_back((int)i);
        const uint32_t m = cap ? (uint32_t)cap - 1 : 0;
        std::stable_sort(o.begin(), o.end(), [&](int a, int b) { return (hash[a] & m) < (hash[b] & m); });
        return o;
    }
};

bool java_ws(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13) || (c >= 28 && c <= 31); }

struct Scanner {                                      // java.util.Scanner over an in-memory file
    std::string buf;
    size_t pos = 0;
    bool hasNext() const { size_t p = pos; while (p < buf.size() && java_ws((unsigned char)buf[p])) ++p; return p < buf.size(); }
    bool hasNextLine() const { return pos < buf.size(); }
    std::string next() {
        while (pos < buf.size() && java_ws((unsigned char)buf[pos])) ++pos;
        const size_t a = pos;
        while (pos < buf.size() && !java_ws((unsigned char)buf[pos])) ++pos;
        return buf.substr(a, pos - a);
    }
    std::string nextLine() { return std::string(nextLineView()); }
    // the same tokens as views into `buf` (no allocation; valid while the Scanner lives)
    std::string_view nextView() {
        while (pos < buf.size() && java_ws((unsigned char)buf[pos])) ++pos;
        const size_t a = pos;
        while (pos < buf.size() && !java_ws((unsigned char)buf[pos])) ++pos;
        return std::string_view(buf).substr(a, pos - a);
    }
    std::string_view nextLineView() {
        const size_t a = pos;
        while (pos < buf.size() && buf[pos] != '\n' && buf[pos] != '\r') ++pos;
        const std::string_view line = std::string_view(buf).substr(a, pos - a);
        if (pos < buf.size()) pos += (buf[pos] == '\r' && pos + 1 < buf.size() && buf[pos + 1] == '\n') ? 2 : 1;
        return line;
    }
};

// String.split(literal) into views (trailing empty strings removed; no separator -> the string itself)
void split_views(std::string_view s, std::string_view sep, std::vector<std::string_view>& out) {
    out.clear();
    size_t p = 0, q;
    bool any = false;
    while ((q = s.find(sep, p)) != std::string_view::npos) { any = true; out.push_back(s.substr(p, q - p)); p = q + sep.size(); }
    if (!any) { out.push_back(s); return; }
    out.push_back(s.substr(p));
    while (!out.empty() && out.back().empty()) out.pop_back();
}
std::string_view trim_view(std::string_view s) {
    size_t a = 0, b = s.size();
    while (a < b && (unsigned char)s[a] <= ' ') ++a;
    while (b > a && (unsigned char)s[b - 1] <= ' ') --b;
    return s.substr(a, b - a);
}

// Interned strings (ESP names, fragment ids: millions at the 5M-fragment scale): NUL-terminated in one arena, found
// through an open-addressing index.  Ids are dense in insertion order.
struct StrTable {
    struct Entry { uint64_t off; uint32_t len; uint32_t tag; };
    std::vector<char> pool;
    std::vector<Entry> ent;
    std::vector<uint64_t> slot;         // open addressing: (tag << 32) | (id + 1), 0 = empty; the tag spares a visit to
                                        // `ent` and `pool` (two more cache misses) for probes that cannot match
    static uint64_t hash(std::string_view s) {
        uint64_t h = 1469598103934665603ULL;
        for (unsigned char c : s) { h ^= c; h *= 1099511628211ULL; }
        return h ^ (h >> 29);
    }
    int size() const { return (int)ent.size(); }
    std::string_view view(int id) const { return std::string_view(pool.data() + ent[id].off, ent[id].len); }
    const char* c_str(int id) const { return pool.data() + ent[id].off; }
    std::string str(int id) const { return std::string(view(id)); }
    bool same(int id, std::string_view s) const {
        return ent[id].len == s.size() && std::memcmp(pool.data() + ent[id].off, s.data(), s.size()) == 0;
    }
    int find(std::string_view s) const {
        if (slot.empty()) return -1;
        const size_t mask = slot.size() - 1;
        const uint64_t h = hash(s);
        const uint32_t tag = (uint32_t)(h >> 32);
        for (size_t i = h & mask;; i = (i + 1) & mask) {
            const uint64_t v = slot[i];
            if (v == 0) return -1;
            if ((uint32_t)(v >> 32) == tag && same((int)(uint32_t)v - 1, s)) return (int)(uint32_t)v - 1;
        }
    }
    int intern(std::string_view s, bool* added = nullptr) {
        if ((ent.size() + 1) * 2 > slot.size()) grow();
        const size_t mask = slot.size() - 1;
        const uint64_t h = hash(s);
        const uint32_t tag = (uint32_t)(h >> 32);
        size_t i = h & mask;
        for (;; i = (i + 1) & mask) {
            const uint64_t v = slot[i];
            if (v == 0) break;
            if ((uint32_t)(v >> 32) == tag && same((int)(uint32_t)v - 1, s)) { if (added) *added = false; return (int)(uint32_t)v - 1; }
        }
        const int id = (int)ent.size();
        slot[i] = ((uint64_t)tag << 32) | (uint32_t)(id + 1);
        ent.push_back(Entry{pool.size(), (uint32_t)s.size(), tag});
        pool.insert(pool.end(), s.begin(), s.end());
        pool.push_back(0);
        if (added) *added = true;
        return id;
    }
    void grow() {
        const size_t n = slot.empty() ? 16 : slot.size() * 2;
        slot.assign(n, 0);
        for (int id = 0; id < (int)ent.size(); ++id) {
            const uint64_t h = hash(view(id));
            size_t i = h & (n - 1);
            while (slot[i] != 0) i = (i + 1) & (n - 1);
            slot[i] = ((uint64_t)ent[id].tag << 32) | (uint32_t)(id + 1);
        }
    }
};
bool parse_int(std::string_view s, int& out) {
    if (s.empty()) return false;
    size_t i = (s[0] == '-' || s[0] == '+') ? 1 : 0;
    if (i >= s.size()) return false;
    long long v = 0;
    for (; i < s.size(); ++i) {
        if (s[i] < '0' || s[i] > '9') return false;
        v = v * 10 + (s[i] - '0');
        if (v > 2147483648LL) return false;
    }
    if (s[0] == '-') v = -v;
    if (v > 2147483647LL || v < -2147483648LL) return false;
    out = (int)v;
    return true;
}
bool parse_double(const std::string& s, double& out) {
    char* end = nullptr;
    out = std::strtod(s.c_str(), &end);
    return end != s.c_str() && *end == 0;
}
bool slurp(const std::string& path, std::string& out) {
    std::FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    out.clear();
    if (std::fseek(f, 0, SEEK_END) == 0) {
        const long n = std::ftell(f);
        std::rewind(f);
        if (n > 0) out.reserve((size_t)n);
    }
    char buf[1 << 16];
    size_t got;
    while ((got = std::fread(buf, 1, sizeof buf, f)) > 0) out.append(buf, got);
    const bool ok = !std::ferror(f);
    std::fclose(f);
    return ok;
}

// Double.toString: shortest digits that round-trip, decimal for 1e-3 <= |v| < 1e7, else d.dddE[-]n
std::string java_double(double v) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v > 0 ? "Infinity" : "-Infinity";
    if (v == 0) return std::signbit(v) ? "-0.0" : "0.0";
    char tmp[64];
    auto res = std::to_chars(tmp, tmp + sizeof tmp, std::fabs(v), std::chars_format::scientific);
    std::string sci(tmp, res.ptr);                     // d[.ddd]e[+-]xx
    const size_t epos = sci.find('e');
    std::string digits;
    for (size_t i = 0; i < epos; ++i) if (sci[i] != '.') digits.push_back(sci[i]);
    const int exp10 = std::atoi(sci.c_str() + epos + 1);
    std::string out = std::signbit(v) ? "-" : "";
    const double a = std::fabs(v);
    if (a >= 1e-3 && a < 1e7) {
        if (exp10 >= 0) {
            std::string ip = digits.substr(0, std::min<size_t>(digits.size(), exp10 + 1));
            while ((int)ip.size() < exp10 + 1) ip.push_back('0');
            std::string fp = digits.size() > (size_t)exp10 + 1 ? digits.substr(exp10 + 1) : "0";
            out += ip + "." + fp;
        } else {
            out += "0." + std::string(-exp10 - 1, '0') + digits;
        }
    } else {
        out += digits.substr(0, 1) + "." + (digits.size() > 1 ? digits.substr(1) : "0") + "E" + std::to_string(exp10);
    }
    return out;
}

struct Alignment { std::vector<int> esps; int numerrors = 0, len = 0, exp = -1; };   // FragmentAlignment
struct Fragment { std::string id; std::vector<Alignment> alns; };                    // Fragment
struct Diagram { bool is_cluster = true; std::vector<int> leaves; std::vector<std::vector<int>> roots; bool built = false; };

}  // namespace

struct mbsv_host {
    std::string clusterfile, assignmentfile, experimentfile;
    bool clusters_first = false;
    // in-memory text instead of the files (mbsv_host_load_partitioned feeds every subset through the same readers)
    const std::string* text_clusters = nullptr; const std::string* text_assignments = nullptr; const std::string* text_experiments = nullptr;
    // ---- parsed state (field names follow the reference) ----
    std::vector<std::string> exp_label; std::vector<double> exp_pseq, exp_lambda;   // by experiment id
    std::unordered_map<std::string, int> exp_id; JavaMapOrder experiments;
    StrTable esp_name;                        // ESP id <-> name
    std::vector<Fragment> frags; StrTable frag_id_of; JavaMapOrder fragments;   // fragment id == its index in frag_id_of
    std::vector<int> esp2fragment;            // ESP id -> fragment id (last writer), -1
    std::vector<char> in_allesps;
    std::vector<int> esp2cid;                 // ESP id -> cluster id (as in `clust`), -1
    std::vector<std::string> cl_name; StrTable cl_id_of; std::vector<char> cl_is_cluster;   // cluster id == index in cl_id_of
    JavaMapOrder clust, breakpoints;          // cluster id == insertion id in `clust`
    std::vector<int> bp_of_cl, cl_of_bp;      // cluster id <-> breakpoints entry id (-1: none)
    int max_support = 0;
    std::vector<Diagram> diagrams;            // by cluster id
    std::string warning, unsupported;
    // ---- packed problem ----
    std::vector<int> frag_order, cl_order;    // fragment ids / cluster ids in keySet() order
    std::vector<int32_t> sub_frag_ptr, sub_cluster_ptr, sub_sv_ptr, frag_aln_ptr, aln_numerrors, aln_len, aln_exp, aln_esp_ptr,
        aln_esp_cluster, aln_esp_leaf, supports_order, cluster_hist_ptr, cluster_leaf_ptr, leaf_owner, cluster_root_ptr,
        root_leaf_ptr, root_leaf, sv_cluster, sv_frag_ptr, sv_frag, sv_numchanged, sorted_frag;
    std::vector<int> aln_esp_espid;
    std::vector<uint8_t> cluster_kind;
    std::vector<double> pk_pseq, pk_lambda, logprobcov;
    bool packed = false, priors = false, sv_filled = false;
    // ---- results ----
    int num_iters = 0, burnin = 0, skip = 1;
    std::vector<uint32_t> r_aln, r_frag, r_hist;
    std::vector<int32_t> r_init, r_trace_frag, r_trace_assign;
    std::vector<double> r_trace_logprob;
    int64_t r_counters[5] = {0, 0, 0, 0, 0};
    double r_final_logprob = 0.0, r_initial_logprob = 0.0;
    bool have_results = false, have_trace = false;

    int esp(std::string_view name) {
        bool added;
        const int id = esp_name.intern(name, &added);
        if (added) { esp2fragment.push_back(-1); in_allesps.push_back(0); esp2cid.push_back(-1); }
        return id;
    }
    void clear() { *this = mbsv_host(); }
};

struct mbsv_host_batch {
    std::vector<int32_t> sub_frag_ptr, sub_cluster_ptr, sub_sv_ptr, frag_aln_ptr, aln_numerrors, aln_len, aln_exp, aln_esp_ptr,
        aln_esp_cluster, aln_esp_leaf, supports_order, cluster_hist_ptr, cluster_leaf_ptr, leaf_owner, cluster_root_ptr,
        root_leaf_ptr, root_leaf, sv_cluster, sv_frag_ptr, sv_frag, sv_numchanged;
    std::vector<uint8_t> cluster_kind;
    std::vector<double> pseq, lambda, logprobcov;
    int max_support = 0, n_exp = 0;
    // for the writers (every host appended so far contributed them; mbsv_host_load_partitioned: flags bit 1)
    bool have_names = true;
    std::vector<char> name_pool;                               // NUL-terminated strings
    std::vector<uint64_t> frag_name, cl_name, esp_name;        // offsets: per fragment / cluster / alignment-ESP entry
    std::vector<int32_t> sorted_frag;                          // per subproblem: its fragments in getFragmentIDs() order (global index)
    std::vector<int32_t> sub_max_support, sub_number;          // the subset's own maxSupport; its subset number (or running index)
};

namespace {

bool fetch(const std::string* text, const std::string& path, std::string& out) {
    if (text) { out = *text; return true; }
    return slurp(path, out);
}

// MultiBreakSV.java:319-331
int read_experiment_file(mbsv_host* h) {
    Scanner sc;
    if (!fetch(h->text_experiments, h->experimentfile, sc.buf)) return hfail(MBSV_ERR_IO, "cannot open " + h->experimentfile);
    if (!sc.hasNextLine()) return hfail(MBSV_ERR_IO, "empty experiment file");
    sc.nextLine();
    while (sc.hasNext()) {
        const std::string label = sc.next();
        double p, l;
        if (!sc.hasNext() || !parse_double(sc.next(), p) || !sc.hasNext() || !parse_double(sc.next(), l))
            return hfail(MBSV_ERR_IO, "InputMismatchException in experiment file");
        auto it = h->exp_id.find(label);
        if (it == h->exp_id.end()) {
            h->exp_id.emplace(label, (int)h->exp_label.size());
            h->exp_label.push_back(label); h->exp_pseq.push_back(p); h->exp_lambda.push_back(l);
            h->experiments.put(label);
        } else { h->exp_pseq[it->second] = p; h->exp_lambda[it->second] = l; }
    }
    return 0;
}

// MultiBreakSV.java:333-400
int read_assignment_file(mbsv_host* h) {
    Scanner sc;
    if (!fetch(h->text_assignments, h->assignmentfile, sc.buf)) return hfail(MBSV_ERR_IO, "cannot open " + h->assignmentfile);
    if (!sc.hasNextLine()) return hfail(MBSV_ERR_IO, "empty assignment file");
    sc.nextLineView();
    std::vector<std::string_view> esps;
    std::vector<int> ids;
    while (sc.hasNext()) {
        const std::string_view fragmentID = sc.nextView();
        if (!sc.hasNext()) return hfail(MBSV_ERR_IO, "Error parsing line. Check assignmentfile format.");
        split_views(sc.nextView(), ",", esps);
        int numerrors, length;
        if (!sc.hasNext() || !parse_int(sc.nextView(), numerrors) || !sc.hasNext() || !parse_int(sc.nextView(), length) || !sc.hasNext())
            return hfail(MBSV_ERR_IO, "Error parsing line. Check assignmentfile format.");
        const std::string explabel(sc.nextView());
        auto ex = h->exp_id.find(explabel);
        if (ex == h->exp_id.end()) return hfail(MBSV_ERR_IO, "ERROR: Experiment Label " + explabel + " not in Experiments file.");
        ids.clear();
        for (const auto& e : esps) ids.push_back(h->esp(e));
        if (h->clusters_first) {
            bool any = false;
            for (int id : ids) if (h->in_allesps[id]) { any = true; break; }
            if (!any) continue;
        }
        bool added;
        const int fid = h->frag_id_of.intern(fragmentID, &added);
        if (added) {
            h->frags.push_back(Fragment{std::string(fragmentID), {}});
            h->fragments.put(fragmentID);
        }
        Alignment al;
        al.esps = ids; al.numerrors = numerrors; al.len = length; al.exp = ex->second;
        h->frags[fid].alns.push_back(std::move(al));
        for (int id : ids) { h->in_allesps[id] = 1; h->esp2fragment[id] = fid; }
    }
    return 0;
}

// MultiBreakSV.java:402-473
int read_clusters_file(mbsv_host* h) {
    Scanner sc;
    if (!fetch(h->text_clusters, h->clusterfile, sc.buf)) return hfail(MBSV_ERR_IO, "cannot open " + h->clusterfile);
    h->max_support = 0;
    bool warn = false;
    std::vector<std::string_view> row, names;
    while (sc.hasNext()) {
        split_views(sc.nextLineView(), "\t", row);
        if (row[0] == "cID") continue;
        const std::string cID(row[0]);
        if (cID.find('.') != std::string::npos) continue;
        int n;
        if (row.size() < 2) return hfail(MBSV_ERR_IO, "ArrayIndexOutOfBounds: clusters row has < 2 columns");
        if (!parse_int(row[1], n)) return hfail(MBSV_ERR_IO, "NumberFormatException: " + std::string(row[1]));
        if (h->max_support < n) h->max_support = n;
        if (row.size() < 5) return hfail(MBSV_ERR_IO, "ArrayIndexOutOfBounds: clusters row has < 5 columns");
        int cid = h->cl_id_of.find(cID);
        auto ensure = [&]() {                       // clust.put(cID, ...) — a repeated cID keeps its slot
            if (cid >= 0) return;
            cid = (int)h->cl_name.size();
            h->cl_id_of.intern(cID);
            h->cl_name.push_back(cID); h->cl_is_cluster.push_back(1); h->bp_of_cl.push_back(-1);
            h->diagrams.emplace_back();
            h->clust.put(cID);
        };
        bool some = false;
        split_views(row[4], ",", names);
        for (const auto& raw : names) {
            const int id = h->esp(trim_view(raw));
            if (!h->clusters_first && !h->in_allesps[id]) { warn = true; continue; }
            some = true;
            ensure();
            h->esp2cid[id] = cid;
            if (h->clusters_first) h->in_allesps[id] = 1;
        }
        if (!h->clusters_first && !some) continue;
        ensure();
        h->cl_is_cluster[cid] = row[2] == "-1" ? 0 : 1;
    }
    if (warn) h->warning = "WARNING! Some fragments in clusters are NOT in ESP file";
    // breakpoints are filled by iterating `clust` (:464-470)
    for (int cid : h->clust.order()) {
        h->bp_of_cl[cid] = h->breakpoints.put(h->cl_name[cid]);
        h->cl_of_bp.push_back(cid);
    }
    return 0;
}

// MultiBreakSV.java:475-550
void intersect_clusters_and_fragments(mbsv_host* h) {
    // the reference collects the names in a HashSet (`toRemove`) and removes them in THAT set's iteration order
    // (:481-516, :525-547); the order of the removals only matters for a map with a treeified bin
    auto remove_in_hashset_order = [](JavaMapOrder& map, const std::vector<int>& ids) {
        JavaMapOrder to_remove;
        for (int id : ids) to_remove.put(map.keys[id]);
        for (int k : to_remove.order()) map.remove(ids[k]);
    };
    std::vector<int> drop;
    for (int fe : h->fragments.order()) {
        Fragment& fr = h->frags[fe];
        std::vector<Alignment> keep;
        for (const Alignment& al : fr.alns) {
            bool all = true;
            for (int e : al.esps) if (h->esp2cid[e] < 0) { all = false; break; }
            if (all) keep.push_back(al);
        }
        if (keep.empty()) drop.push_back(fe); else fr.alns.swap(keep);
    }
    remove_in_hashset_order(h->fragments, drop);
    std::vector<char> used(h->cl_name.size(), 0);
    for (int fe : h->fragments.order())
        for (const Alignment& al : h->frags[fe].alns)
            for (int e : al.esps) used[h->esp2cid[e]] = 1;
    drop.clear();
    for (int be : h->breakpoints.order())
        if (!used[h->cl_of_bp[be]]) { drop.push_back(be); h->bp_of_cl[h->cl_of_bp[be]] = -1; }
    remove_in_hashset_order(h->breakpoints, drop);
}

// MultiBreakSV.java:658-746 + MultiBreakSVDiagram.java:56-139
int construct_cluster_diagrams(mbsv_host* h) {
    Scanner sc;
    if (!fetch(h->text_clusters, h->clusterfile, sc.buf)) return hfail(MBSV_ERR_IO, "cannot open " + h->clusterfile);
    std::vector<std::string_view> row, list_items;
    auto advance = [&]() -> bool { if (!sc.hasNext()) return false; split_views(sc.nextLineView(), "\t", row); return true; };
    bool have = advance();
    std::vector<int> seen_stamp;              // ESP id -> last cluster row group that listed it (leaves are a set)
    int stamp = 0;
    while (have) {
        if (row[0] == "cID") { have = advance(); continue; }
        if (row.size() < 5) return hfail(MBSV_ERR_IO, "ArrayIndexOutOfBounds: clusters row has < 5 columns");
        const std::string cID(row[0]);
        const bool simple = row[2] != "-1";
        const std::string_view names = row[4];
        const int cid = h->cl_id_of.find(cID);
        if (cid < 0 || h->bp_of_cl[cid] < 0) { have = advance(); continue; }
        Diagram d;
        d.built = true;
        ++stamp;
        auto add_root = [&](std::string_view list) {
            std::vector<int> r;
            split_views(list, ", ", list_items);
            for (const auto& e : list_items) {
                const int id = h->esp(e);
                r.push_back(id);
                if ((size_t)id >= seen_stamp.size()) seen_stamp.resize(std::max<size_t>(h->esp_name.size(), (size_t)id + 1), 0);
                if (seen_stamp[id] != stamp) { seen_stamp[id] = stamp; d.leaves.push_back(id); }
            }
            d.roots.push_back(std::move(r));
        };
        if (simple) {
            d.is_cluster = true;
            add_root(names);
            have = advance();
        } else {
            d.is_cluster = false;
            have = advance();
            while (have && row[0].compare(0, cID.size(), cID) == 0) {
                if (row.size() < 5) return hfail(MBSV_ERR_IO, "ArrayIndexOutOfBounds: sub-cluster row");
                add_root(row[4]);
                have = advance();
            }
        }
        for (int e : d.leaves) if (h->esp2cid[e] >= 0) h->esp2cid[e] = cid;
        h->diagrams[cid] = std::move(d);
    }
    for (int be : h->breakpoints.order())
        if (!h->diagrams[h->cl_of_bp[be]].built) return hfail(MBSV_ERR_IO, "NullPointerException: no diagram for " + h->cl_name[h->cl_of_bp[be]]);
    return 0;
}

// Flatten into mbsv_problem_desc arrays (same layout as the oracle's packing; compared integer-for-integer in tests).
int pack(mbsv_host* h) {
    const std::vector<int> eorder = h->experiments.order();
    const int E = (int)eorder.size();
    std::vector<int> exp_idx(h->exp_label.size(), -1);
    h->pk_pseq.clear(); h->pk_lambda.clear();
    for (int i = 0; i < E; ++i) { exp_idx[eorder[i]] = i; h->pk_pseq.push_back(h->exp_pseq[eorder[i]]); h->pk_lambda.push_back(h->exp_lambda[eorder[i]]); }
    // clusters in breakpoints.keySet() order
    h->cl_order.clear();
    for (int be : h->breakpoints.order()) h->cl_order.push_back(h->cl_of_bp[be]);
    const int C = (int)h->cl_order.size();
    std::vector<int> cl_idx(h->cl_name.size(), -1);
    for (int i = 0; i < C; ++i) cl_idx[h->cl_order[i]] = i;
    h->cluster_kind.assign(C, 0);
    for (int i = 0; i < C; ++i) h->cluster_kind[i] = h->cl_is_cluster[h->cl_order[i]] ? 0 : 1;
    {   // A.supports is a fresh HashMap filled by iterating breakpoints (MultiBreakSVAssignmentMatrix.java:53-66)
        JavaMapOrder sup;
        for (int i = 0; i < C; ++i) sup.put(h->cl_name[h->cl_order[i]]);
        h->supports_order.clear();
        for (int id : sup.order()) h->supports_order.push_back(id);
    }
    // fragments in fragments.keySet() order
    h->frag_order = h->fragments.order();
    const int F = (int)h->frag_order.size();
    std::vector<int> frag_idx(h->frags.size(), -1);
    for (int i = 0; i < F; ++i) frag_idx[h->frag_order[i]] = i;
    // leaves / roots
    h->cluster_leaf_ptr.assign(1, 0); h->leaf_owner.clear(); h->cluster_root_ptr.assign(1, 0); h->root_leaf_ptr.assign(1, 0);
    h->root_leaf.clear();
    // An ESP is a leaf of at most one cluster in accepted input (anything else is flagged below), so "slot of ESP e among
    // the leaves of cluster i" is two flat arrays instead of a map per cluster; `mark` detects repeats within one list.
    const size_t nesp = h->esp_name.size();
    std::vector<int> leaf_count(nesp, 0), leaf_cluster(nesp, -1), leaf_slot(nesp, -1), mark(nesp, 0);
    int stamp = 0;
    for (int i = 0; i < C; ++i) {
        const Diagram& d = h->diagrams[h->cl_order[i]];
        if (d.is_cluster != (h->cluster_kind[i] == 0)) h->unsupported = "cluster kind mismatch for " + h->cl_name[h->cl_order[i]];
        for (int e : d.leaves) {
            leaf_cluster[e] = i;
            leaf_slot[e] = (int)h->leaf_owner.size() - h->cluster_leaf_ptr[i];
            ++leaf_count[e];
            const int fe = h->esp2fragment[e];
            h->leaf_owner.push_back(fe >= 0 && h->fragments.alive[fe] ? frag_idx[fe] : -1);
        }
        h->cluster_leaf_ptr.push_back((int)h->leaf_owner.size());
        for (const auto& r : d.roots) {
            ++stamp;
            for (int e : r) {
                if (mark[e] == stamp) h->unsupported = "ESP " + h->esp_name.str(e) + " listed twice in one row of cluster " + h->cl_name[h->cl_order[i]];
                mark[e] = stamp;
                // (every ESP of a root was added to this diagram's leaves; leaf_cluster[e] != i only for rejected input)
                h->root_leaf.push_back(leaf_cluster[e] == i ? leaf_slot[e] : 0);
            }
            h->root_leaf_ptr.push_back((int)h->root_leaf.size());
        }
        h->cluster_root_ptr.push_back((int)h->root_leaf_ptr.size() - 1);
        if (h->cluster_leaf_ptr[i + 1] - h->cluster_leaf_ptr[i] > h->max_support)
            h->unsupported = "cluster " + h->cl_name[h->cl_order[i]] + " has more ESPs than its declared count (logprobcov overflow)";
    }
    for (size_t e = 0; e < leaf_count.size(); ++e)
        if (leaf_count[e] > 1) h->unsupported = "ESP " + h->esp_name.str((int)e) + " is a leaf of more than one cluster";
    // alignments
    h->frag_aln_ptr.assign(1, 0); h->aln_esp_ptr.assign(1, 0);
    h->aln_numerrors.clear(); h->aln_len.clear(); h->aln_exp.clear(); h->aln_esp_cluster.clear(); h->aln_esp_leaf.clear();
    h->aln_esp_espid.clear();
    h->aln_esp_cluster.reserve(nesp); h->aln_esp_leaf.reserve(nesp); h->aln_esp_espid.reserve(nesp);
    for (int i = 0; i < F; ++i) {
        const Fragment& fr = h->frags[h->frag_order[i]];
        for (const Alignment& al : fr.alns) {
            h->aln_numerrors.push_back(al.numerrors); h->aln_len.push_back(al.len); h->aln_exp.push_back(exp_idx[al.exp]);
            ++stamp;
            for (int e : al.esps) {
                if (mark[e] == stamp) h->unsupported = "ESP " + h->esp_name.str(e) + " listed twice in one alignment of " + fr.id;
                mark[e] = stamp;
                const int c = cl_idx[h->esp2cid[e]];
                const bool is_leaf = c >= 0 && leaf_cluster[e] == c;
                const bool contributes = is_leaf && h->leaf_owner[h->cluster_leaf_ptr[c] + leaf_slot[e]] == i;
                h->aln_esp_cluster.push_back(contributes ? c : -1);
                h->aln_esp_leaf.push_back(contributes ? leaf_slot[e] : -1);
                h->aln_esp_espid.push_back(e);
            }
            h->aln_esp_ptr.push_back((int)h->aln_esp_cluster.size());
        }
        h->frag_aln_ptr.push_back((int)h->aln_numerrors.size());
    }
    h->sorted_frag.resize(F);
    for (int i = 0; i < F; ++i) h->sorted_frag[i] = i;
    std::sort(h->sorted_frag.begin(), h->sorted_frag.end(),
              [&](int a, int b) { return h->frags[h->frag_order[a]].id < h->frags[h->frag_order[b]].id; });
    const std::vector<int32_t> nleaf_ptr = h->cluster_leaf_ptr;
    h->cluster_hist_ptr.assign(1, 0);
    for (int i = 0; i < C; ++i) h->cluster_hist_ptr.push_back(h->cluster_hist_ptr.back() + (nleaf_ptr[i + 1] - nleaf_ptr[i]) + 1);
    h->sub_frag_ptr = {0, F};
    h->sub_cluster_ptr = {0, C};
    h->packed = true;
    return 0;
}

// MultiBreakSV.java:604-656: logprobcov[k] = logPoisson(k; lambdaD), k = 1..maxSupport (prior[] == 0)
int precompute_priors(mbsv_host* h) {
    const int E = (int)h->pk_pseq.size();
    h->logprobcov.assign((size_t)E * (h->max_support + 1), 0.0);
    for (int x = 0; x < E; ++x)
        if (int rc = mbsv_fill_logprobcov(h->pk_lambda[x], h->max_support, h->logprobcov.data() + (size_t)x * (h->max_support + 1))) return rc;
    h->priors = true;
    return 0;
}

// MultiBreakSV.java:748-771 and the toggle lists of svMove :1303-1330
void fill_non_singleton_array(mbsv_host* h) {
    const int C = (int)h->cl_order.size();
    h->sv_cluster.clear(); h->sv_frag.clear(); h->sv_numchanged.clear(); h->sv_frag_ptr.assign(1, 0);
    for (int c = 0; c < C; ++c) {
        int unique = 0;
        std::vector<int> uniq;
        for (int l = h->cluster_leaf_ptr[c]; l < h->cluster_leaf_ptr[c + 1]; ++l) {
            const int f = h->leaf_owner[l];
            if (f >= 0 && h->frag_aln_ptr[f + 1] - h->frag_aln_ptr[f] == 1) {
                ++unique;
                if (std::find(uniq.begin(), uniq.end(), f) == uniq.end()) uniq.push_back(f);
            }
        }
        if (unique > 1) {
            h->sv_cluster.push_back(c);
            h->sv_numchanged.push_back(unique);
            for (int f : uniq) h->sv_frag.push_back(f);
            h->sv_frag_ptr.push_back((int)h->sv_frag.size());
        }
    }
    h->sub_sv_ptr = {0, (int)h->sv_cluster.size()};
    h->sv_filled = true;
}

template <typename T> const T* ptr_or_null(const std::vector<T>& v) { return v.empty() ? nullptr : v.data(); }

}  // namespace

extern "C" {

int mbsv_host_create(mbsv_host** out) {
    if (!out) return hfail(MBSV_ERR_ARG, "out is NULL");
    *out = new (std::nothrow) mbsv_host();
    return *out ? MBSV_OK : hfail(MBSV_ERR_OOM, "allocation failed");
}
int mbsv_host_destroy(mbsv_host* h) { delete h; return MBSV_OK; }

int mbsv_host_set_files(mbsv_host* h, const char* clusterfile, const char* assignmentfile, const char* experimentfile,
                        int32_t clusters_first) {
    if (!h || !clusterfile || !assignmentfile || !experimentfile) return hfail(MBSV_ERR_ARG, "Cluster, assignment and experiment files are required.");
    h->clear();
    h->clusterfile = clusterfile; h->assignmentfile = assignmentfile; h->experimentfile = experimentfile;
    h->clusters_first = clusters_first != 0;
    return MBSV_OK;
}

namespace {
struct Lap {   // MBSV_DEBUG_TIMING=1: phase times of the host mirror on stderr
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void operator()(const char* what) {
        static const bool on = std::getenv("MBSV_DEBUG_TIMING") != nullptr;
        const auto n = std::chrono::steady_clock::now();
        const double dt = std::chrono::duration<double>(n - t).count();
        if (on && dt >= 0.005) std::fprintf(stderr, "[mbsv host] %-28s %.3f s\n", what, dt);   // (not for the many tiny subsets)
        t = n;
    }
};
}  // namespace

int mbsv_host_get_fragment_alignments(mbsv_host* h) {
    if (!h) return hfail(MBSV_ERR_ARG, "host is NULL");
    int rc;
    Lap lap;
    if ((rc = read_experiment_file(h))) return rc;
    if (h->clusters_first) {
        if ((rc = read_clusters_file(h))) return rc;
        lap("readClustersFile");
        if ((rc = read_assignment_file(h))) return rc;
        lap("readAssignmentFile");
    } else {
        if ((rc = read_assignment_file(h))) return rc;
        lap("readAssignmentFile");
        if ((rc = read_clusters_file(h))) return rc;
        lap("readClustersFile");
    }
    intersect_clusters_and_fragments(h);
    lap("intersectClustersAndFragments");
    return MBSV_OK;
}
int mbsv_host_construct_cluster_diagrams(mbsv_host* h) {
    if (!h) return hfail(MBSV_ERR_ARG, "host is NULL");
    Lap lap;
    if (int rc = construct_cluster_diagrams(h)) return rc;
    lap("constructClusterDiagrams");
    const int rc = pack(h);
    lap("pack");
    return rc;
}
int mbsv_host_precompute_priors(mbsv_host* h) {
    if (!h || !h->packed) return hfail(MBSV_ERR_ARG, "construct the cluster diagrams first");
    return precompute_priors(h);
}
int mbsv_host_fill_non_singleton_array(mbsv_host* h) {
    if (!h || !h->packed) return hfail(MBSV_ERR_ARG, "construct the cluster diagrams first");
    fill_non_singleton_array(h);
    return MBSV_OK;
}
int mbsv_host_load(mbsv_host* h) {
    int rc;
    if ((rc = mbsv_host_get_fragment_alignments(h)) || (rc = mbsv_host_construct_cluster_diagrams(h)) ||
        (rc = mbsv_host_precompute_priors(h)) || (rc = mbsv_host_fill_non_singleton_array(h))) return rc;
    return MBSV_OK;
}

int mbsv_host_desc(mbsv_host* h, mbsv_problem_desc* d) {
    if (!h || !d || !h->packed || !h->priors || !h->sv_filled) return hfail(MBSV_ERR_ARG, "mbsv_host_load first");
    if (!h->unsupported.empty()) return hfail(MBSV_ERR_UNSUPPORTED, h->unsupported);
    std::memset(d, 0, sizeof *d);
    d->abi_version = MBSV_ABI_VERSION;
    d->n_sub = 1;
    d->n_frag = (int)h->frag_order.size(); d->n_aln = (int)h->aln_numerrors.size(); d->n_aln_esp = (int)h->aln_esp_cluster.size();
    d->n_cluster = (int)h->cl_order.size(); d->n_exp = (int)h->pk_pseq.size(); d->max_support = h->max_support;
    d->n_sv = (int)h->sv_cluster.size(); d->n_sv_frag = (int)h->sv_frag.size(); d->n_leaf = (int)h->leaf_owner.size();
    d->n_root = (int)h->root_leaf_ptr.size() - 1; d->n_root_leaf = (int)h->root_leaf.size(); d->n_hist = h->cluster_hist_ptr.back();
    d->sub_frag_ptr = h->sub_frag_ptr.data(); d->sub_cluster_ptr = h->sub_cluster_ptr.data(); d->sub_sv_ptr = h->sub_sv_ptr.data();
    d->frag_aln_ptr = h->frag_aln_ptr.data(); d->aln_numerrors = ptr_or_null(h->aln_numerrors); d->aln_len = ptr_or_null(h->aln_len);
    d->aln_exp = ptr_or_null(h->aln_exp); d->aln_esp_ptr = h->aln_esp_ptr.data(); d->aln_esp_cluster = ptr_or_null(h->aln_esp_cluster);
    d->aln_esp_leaf = ptr_or_null(h->aln_esp_leaf); d->cluster_kind = ptr_or_null(h->cluster_kind);
    d->supports_order = ptr_or_null(h->supports_order); d->cluster_hist_ptr = h->cluster_hist_ptr.data();
    d->cluster_leaf_ptr = h->cluster_leaf_ptr.data(); d->leaf_owner = ptr_or_null(h->leaf_owner);
    d->cluster_root_ptr = h->cluster_root_ptr.data(); d->root_leaf_ptr = h->root_leaf_ptr.data(); d->root_leaf = ptr_or_null(h->root_leaf);
    d->sv_cluster = ptr_or_null(h->sv_cluster); d->sv_frag_ptr = h->sv_frag_ptr.data(); d->sv_frag = ptr_or_null(h->sv_frag);
    d->sv_numchanged = ptr_or_null(h->sv_numchanged); d->exp_pseq = h->pk_pseq.data(); d->exp_lambda = h->pk_lambda.data();
    d->logprobcov = h->logprobcov.data();
    return MBSV_OK;
}

const char* mbsv_host_name(mbsv_host* h, int32_t kind, int32_t i) {
    if (!h) return nullptr;
    switch (kind) {
        case 0: return (i >= 0 && i < (int)h->frag_order.size()) ? h->frags[h->frag_order[i]].id.c_str() : nullptr;
        case 1: return (i >= 0 && i < (int)h->cl_order.size()) ? h->cl_name[h->cl_order[i]].c_str() : nullptr;
        case 2: { const auto o = h->experiments.order(); return (i >= 0 && i < (int)o.size()) ? h->exp_label[o[i]].c_str() : nullptr; }
        case 3: return (i >= 0 && i < (int)h->aln_esp_espid.size()) ? h->esp_name.c_str(h->aln_esp_espid[i]) : nullptr;
        case 4: return h->warning.c_str();
        case 5: return h->unsupported.c_str();
        default: return nullptr;
    }
}
int mbsv_host_sorted_fragments(mbsv_host* h, int32_t* out) {
    if (!h || !out || !h->packed) return hfail(MBSV_ERR_ARG, "mbsv_host_load first");
    for (size_t i = 0; i < h->sorted_frag.size(); ++i) out[i] = h->sorted_frag[i];
    return (int)h->sorted_frag.size();
}

int mbsv_host_set_results(mbsv_host* h, int32_t num_iters, int32_t burnin, const uint32_t* aln_count,
                          const uint32_t* frag_err_count, const uint32_t* cluster_hist) {
    if (!h || !h->packed || !aln_count || !frag_err_count || !cluster_hist) return hfail(MBSV_ERR_ARG, "bad argument");
    h->num_iters = num_iters; h->burnin = burnin;
    h->skip = std::max(1, num_iters / 1000);
    h->r_aln.assign(aln_count, aln_count + h->aln_numerrors.size());
    h->r_frag.assign(frag_err_count, frag_err_count + h->frag_order.size());
    h->r_hist.assign(cluster_hist, cluster_hist + h->cluster_hist_ptr.back());
    h->have_results = true;
    h->have_trace = false;
    return MBSV_OK;
}

int mbsv_host_run_mcmc(mbsv_host* h, int32_t num_iters, double perr, int32_t mode, int32_t device, int32_t save_space,
                       const int32_t* init_assign, mbsv_results* results) {
    mbsv_problem_desc d;
    if (int rc = mbsv_host_desc(h, &d)) return rc;
    if (num_iters < 0) return hfail(MBSV_ERR_ARG, "--numiterations must be a positive integer.");
    if (!(perr >= 0 && perr <= 1)) return hfail(MBSV_ERR_ARG, "--perr must be a value between 0 and 1.");
    // burnin / skip as the constructor derives them (MultiBreakSV.java:186-193)
    h->num_iters = num_iters;
    h->burnin = num_iters / 10 ? num_iters / 10 : 10;
    h->skip = num_iters / 1000 ? num_iters / 1000 : 1;
    mbsv_problem* prob = nullptr;
    if (int rc = mbsv_problem_create(&d, device, &prob)) return hfail(rc, mbsv_last_error());
    mbsv_run_params rp;
    std::memset(&rp, 0, sizeof rp);
    rp.mode = mode; rp.n_chains = 1; rp.num_iters = num_iters; rp.burnin = h->burnin; rp.perr = perr; rp.seed = 123456;
    rp.device = device; rp.java_int_wrap = 1; rp.init_assign = init_assign; rp.trace = save_space ? 0 : 1; rp.memo_states = 65536;
    h->r_aln.assign(d.n_aln, 0); h->r_frag.assign(d.n_frag, 0); h->r_hist.assign(d.n_hist, 0); h->r_init.assign(d.n_frag, 0);
    std::vector<int32_t> fin(d.n_frag, 0), errv(1, 0);
    std::vector<double> flp(1, 0.0), ilp(1, 0.0);
    std::vector<int64_t> counters(5, 0), totals(6, 0);
    std::vector<uint64_t> rngs(1, 0);
    if (!save_space) { h->r_trace_logprob.assign(num_iters, 0.0); h->r_trace_frag.assign(num_iters, -1); h->r_trace_assign.assign(num_iters, -1); }
    mbsv_results r;
    std::memset(&r, 0, sizeof r);
    r.aln_count = h->r_aln.data(); r.frag_err_count = h->r_frag.data(); r.cluster_hist = h->r_hist.data();
    r.final_assign = fin.data(); r.final_logprob = flp.data(); r.counters = counters.data(); r.rng_state = rngs.data();
    r.error = errv.data(); r.initial_assign = h->r_init.data(); r.initial_logprob = ilp.data(); r.totals = totals.data();
    if (!save_space) { r.trace_logprob = h->r_trace_logprob.data(); r.trace_move_frag = h->r_trace_frag.data(); r.trace_move_assign = h->r_trace_assign.data(); }
    const int rc = mbsv_run(prob, &rp, &r);
    mbsv_problem_destroy(prob);
    if (rc) return hfail(rc, mbsv_last_error());
    if (errv[0]) return hfail(MBSV_ERR_INVARIANT, "the chain stopped on one of the reference's Error conditions (MultiBreakSV.java:1122-1144)");
    for (int i = 0; i < 5; ++i) h->r_counters[i] = counters[i];
    h->r_final_logprob = flp[0]; h->r_initial_logprob = ilp[0];
    h->have_results = true;
    h->have_trace = !save_space;
    if (results) { *results = r; results->final_assign = nullptr; results->final_logprob = nullptr; results->counters = nullptr;
                   results->rng_state = nullptr; results->error = nullptr; results->initial_logprob = nullptr; results->totals = nullptr; }
    return MBSV_OK;
}

int mbsv_host_run_summary(mbsv_host* h, int64_t* counters, double* initial_logprob, double* final_logprob) {
    if (!h || !h->have_results) return hfail(MBSV_ERR_ARG, "no results: mbsv_host_run_mcmc first");
    if (counters) for (int i = 0; i < 5; ++i) counters[i] = h->r_counters[i];
    if (initial_logprob) *initial_logprob = h->r_initial_logprob;
    if (final_logprob) *final_logprob = h->r_final_logprob;
    return MBSV_OK;
}

// The --matchfile branch of initializeMatrix (MultiBreakSV.java:1030-1085): the file is whitespace-separated ESPs.
// Per fragment, the first alignment with `thisval == alignments.size()` is set, where thisval counts at most ONE
// listed ESP (the inner loop breaks at the first hit, :1058-1063) -- so only single-ESP alignments can ever be set;
// every other fragment starts "unmapped".  counts = {numset, numerr, numbad, ESPs read} as the reference prints them.
int mbsv_host_matchfile_initial(mbsv_host* h, const char* matchfile, int32_t* init_assign, int32_t* counts) {
    if (!h || !h->packed || !matchfile || !init_assign) return hfail(MBSV_ERR_ARG, "mbsv_host_load first");
    Scanner sc;
    if (!slurp(matchfile, sc.buf)) return hfail(MBSV_ERR_IO, std::string("cannot open ") + matchfile);
    std::unordered_map<std::string, char> listed;
    int nread = 0;
    while (sc.hasNext()) { listed.emplace(sc.next(), 1); ++nread; }
    int numset = 0, numerr = 0, numbad = 0;
    const int F = (int)h->frag_order.size();
    for (int f = 0; f < F; ++f) {
        init_assign[f] = -1;
        bool foundall = false, foundany = false;
        for (int a = h->frag_aln_ptr[f]; a < h->frag_aln_ptr[f + 1] && !foundall; ++a) {
            int thisval = 0;
            // FragmentAlignment.contains(esp) is membership of the alignment's ESP list; which listed ESP hits first
            // does not matter, only whether any does
            for (int j = h->aln_esp_ptr[a]; j < h->aln_esp_ptr[a + 1]; ++j)
                if (listed.count(h->esp_name.str(h->aln_esp_espid[j]))) { thisval = 1; foundany = true; break; }
            if (thisval == h->aln_esp_ptr[a + 1] - h->aln_esp_ptr[a]) { init_assign[f] = a - h->frag_aln_ptr[f]; ++numset; foundall = true; }
        }
        if (!foundany && !foundall) ++numerr;
        if (foundany && !foundall) ++numbad;
    }
    if (counts) { counts[0] = numset; counts[1] = numerr; counts[2] = numbad; counts[3] = nread; }
    return MBSV_OK;
}

// MultiBreakSV.java:1754-1792
int mbsv_host_write_final_files(mbsv_host* h, const char* prefix) {
    if (!h || !prefix || !h->have_results) return hfail(MBSV_ERR_ARG, "no results to write");
    const int C = (int)h->cl_order.size(), F = (int)h->frag_order.size();
    {
        std::ofstream w(std::string(prefix) + ".finalbreakpoints.longburnin");
        if (!w) return hfail(MBSV_ERR_IO, "cannot write breakpoints file");
        w << "ClusterID\tNumIters\tSupport0\tSupport1\tSupport2...\n";
        for (int c = 0; c < C; ++c) {
            w << h->cl_name[h->cl_order[c]] << "\t" << h->burnin;
            const uint32_t* hist = h->r_hist.data() + h->cluster_hist_ptr[c];
            const int bins = h->cluster_hist_ptr[c + 1] - h->cluster_hist_ptr[c];
            double tot = 0;
            for (int j = 1; j <= h->max_support; ++j) tot = tot + (j < bins ? (double)hist[j] : 0.0);
            w << "\t" << java_double((double)h->burnin - tot);           // bin 0 is derived (:1766-1771)
            for (int j = 1; j <= h->max_support; ++j) w << "\t" << java_double(j < bins ? (double)hist[j] : 0.0);
            w << "\n";
        }
    }
    {
        std::ofstream w(std::string(prefix) + ".finalassignment.longburnin");
        if (!w) return hfail(MBSV_ERR_IO, "cannot write assignment file");
        w << "FragmentID\tAssignmentIndex\tAvgAssignment\tDiscordantPairs\n";
        for (int k = 0; k < F; ++k) {
            const int f = h->sorted_frag[k];
            const std::string& id = h->frags[h->frag_order[f]].id;
            w << id << "\t-1\t" << java_double((double)h->r_frag[f] / h->burnin) << "\terror\n";
            for (int a = h->frag_aln_ptr[f]; a < h->frag_aln_ptr[f + 1]; ++a) {
                w << id << "\t" << (a - h->frag_aln_ptr[f]) << "\t" << java_double((double)h->r_aln[a] / h->burnin) << "\t";
                for (int j = h->aln_esp_ptr[a]; j < h->aln_esp_ptr[a + 1]; ++j)
                    w << (j > h->aln_esp_ptr[a] ? "," : "") << h->esp_name.view(h->aln_esp_espid[j]);
                w << "\n";
            }
        }
    }
    return MBSV_OK;
}

// runMCMC's writer (:779-828): Iteration, LogProb, NumBreakpoints (sum of supports.get(cID).size() = C*E every
// `skip` iterations, else 0), MadeMove?
int mbsv_host_write_mcmc_file(mbsv_host* h, const char* path) {
    if (!h || !path || !h->have_trace) return hfail(MBSV_ERR_ARG, "no trace (run without --savespace)");
    std::ofstream w(path);
    if (!w) return hfail(MBSV_ERR_IO, "cannot write mcmc file");
    w << "Iteration\tLogProb\tNumBreakpoints\tMadeMove?\n";
    const long long numbps = (long long)h->cl_order.size() * (long long)h->pk_pseq.size();
    for (int i = 0; i < h->num_iters; ++i)
        w << i << "\t" << java_double(h->r_trace_logprob[i]) << "\t" << (i % h->skip == 0 ? numbps : 0) << "\t"
          << (h->r_trace_frag[i] != -1 ? 1 : 0) << "\n";
    return MBSV_OK;
}

// writeStateCounts (:1934-1949): counts/logprobs are keyed by the tab-joined assignment string in sorted-fragment
// order (MultiBreakSVAssignmentMatrix.java:465-472) and iterated in HashMap order.
int mbsv_host_write_state_counts(mbsv_host* h, const char* path) {
    if (!h || !path || !h->have_trace) return hfail(MBSV_ERR_ARG, "no trace (run without --savespace)");
    const int F = (int)h->frag_order.size();
    std::vector<int32_t> cur = h->r_init;
    JavaMapOrder counts;
    std::unordered_map<std::string, int> id_of;
    std::vector<std::string> keys;
    std::vector<long long> cnt;
    std::vector<double> lp;
    for (int i = 0; i < h->num_iters; ++i) {
        const int mf = h->r_trace_frag[i];
        if (mf >= 0) cur[mf] = h->r_trace_assign[i];
        else if (mf <= -2) {
            const int k = -mf - 2;
            for (int q = h->sv_frag_ptr[k]; q < h->sv_frag_ptr[k + 1]; ++q) cur[h->sv_frag[q]] = cur[h->sv_frag[q]] == -1 ? 0 : -1;
        }
        std::string key = std::to_string(cur[h->sorted_frag[0]]);
        for (int k = 1; k < F; ++k) key += "\t" + std::to_string(cur[h->sorted_frag[k]]);
        auto it = id_of.find(key);
        if (it == id_of.end()) {
            id_of.emplace(key, (int)keys.size());
            keys.push_back(key); cnt.push_back(1); lp.push_back(h->r_trace_logprob[i]);
            counts.put(key);
        } else ++cnt[it->second];
    }
    std::ofstream w(path);
    if (!w) return hfail(MBSV_ERR_IO, "cannot write state counts");
    w << "SampledProb\tLogProb";
    for (int k = 0; k < F; ++k) w << "\t" << h->frags[h->frag_order[h->sorted_frag[k]]].id;
    w << "\n";
    for (int id : counts.order()) w << java_double((double)cnt[id] / (double)h->num_iters) << "\t" << java_double(lp[id]) << "\t" << keys[id] << "\n";
    return MBSV_OK;
}

static int partition_lines(const std::string& buf, std::vector<std::string_view>& lines, std::vector<int>& line_subset);

// one more subproblem at the end of the batch (the host may be reused or destroyed afterwards)
static int batch_append(mbsv_host_batch* b, const mbsv_host* h, bool keep_names = true, int number = 0) {
    if (!h || !h->packed || !h->priors || !h->sv_filled) return hfail(MBSV_ERR_ARG, "every host must be loaded");
    if (!h->unsupported.empty()) return hfail(MBSV_ERR_UNSUPPORTED, h->unsupported);
    if (b->sub_frag_ptr.empty()) {
        b->n_exp = (int)h->pk_pseq.size();
        b->pseq = h->pk_pseq; b->lambda = h->pk_lambda;
        b->sub_frag_ptr = {0}; b->sub_cluster_ptr = {0}; b->sub_sv_ptr = {0}; b->frag_aln_ptr = {0}; b->aln_esp_ptr = {0};
        b->cluster_hist_ptr = {0}; b->cluster_leaf_ptr = {0}; b->cluster_root_ptr = {0}; b->root_leaf_ptr = {0}; b->sv_frag_ptr = {0};
    } else if (h->pk_pseq != b->pseq || h->pk_lambda != b->lambda) {
        return hfail(MBSV_ERR_ARG, "hosts of one batch must share the experiments file");
    }
    b->max_support = std::max(b->max_support, h->max_support);
    const int f0 = b->sub_frag_ptr.back(), c0 = b->sub_cluster_ptr.back(), a0 = b->frag_aln_ptr.back(), k0 = b->aln_esp_ptr.back();
    const int l0 = b->cluster_leaf_ptr.back(), r0 = b->cluster_root_ptr.back(), rl0 = b->root_leaf_ptr.back();
    const int h0 = b->cluster_hist_ptr.back(), q0 = b->sv_frag_ptr.back();
    const int F = (int)h->frag_order.size(), C = (int)h->cl_order.size();
    for (int f = 1; f <= F; ++f) b->frag_aln_ptr.push_back(a0 + h->frag_aln_ptr[f]);
    b->aln_numerrors.insert(b->aln_numerrors.end(), h->aln_numerrors.begin(), h->aln_numerrors.end());
    b->aln_len.insert(b->aln_len.end(), h->aln_len.begin(), h->aln_len.end());
    b->aln_exp.insert(b->aln_exp.end(), h->aln_exp.begin(), h->aln_exp.end());
    for (size_t a = 1; a < h->aln_esp_ptr.size(); ++a) b->aln_esp_ptr.push_back(k0 + h->aln_esp_ptr[a]);
    for (int c : h->aln_esp_cluster) b->aln_esp_cluster.push_back(c < 0 ? -1 : c0 + c);
    b->aln_esp_leaf.insert(b->aln_esp_leaf.end(), h->aln_esp_leaf.begin(), h->aln_esp_leaf.end());
    b->cluster_kind.insert(b->cluster_kind.end(), h->cluster_kind.begin(), h->cluster_kind.end());
    for (int c : h->supports_order) b->supports_order.push_back(c0 + c);
    for (int c = 1; c <= C; ++c) {
        b->cluster_hist_ptr.push_back(h0 + h->cluster_hist_ptr[c]);
        b->cluster_leaf_ptr.push_back(l0 + h->cluster_leaf_ptr[c]);
        b->cluster_root_ptr.push_back(r0 + h->cluster_root_ptr[c]);
    }
    for (int o : h->leaf_owner) b->leaf_owner.push_back(o < 0 ? -1 : f0 + o);
    for (size_t r = 1; r < h->root_leaf_ptr.size(); ++r) b->root_leaf_ptr.push_back(rl0 + h->root_leaf_ptr[r]);
    b->root_leaf.insert(b->root_leaf.end(), h->root_leaf.begin(), h->root_leaf.end());
    for (int c : h->sv_cluster) b->sv_cluster.push_back(c0 + c);
    for (size_t k = 1; k < h->sv_frag_ptr.size(); ++k) b->sv_frag_ptr.push_back(q0 + h->sv_frag_ptr[k]);
    for (int f : h->sv_frag) b->sv_frag.push_back(f0 + f);
    b->sv_numchanged.insert(b->sv_numchanged.end(), h->sv_numchanged.begin(), h->sv_numchanged.end());
    b->sub_frag_ptr.push_back(f0 + F); b->sub_cluster_ptr.push_back(c0 + C);
    b->sub_sv_ptr.push_back(b->sub_sv_ptr.back() + (int)h->sv_cluster.size());
    b->sub_max_support.push_back(h->max_support);
    b->sub_number.push_back(number > 0 ? number : (int)b->sub_max_support.size());
    if (!keep_names) b->have_names = false;
    if (b->have_names) {
        auto put = [&](std::string_view v) { const uint64_t o = b->name_pool.size(); b->name_pool.insert(b->name_pool.end(), v.begin(), v.end()); b->name_pool.push_back(0); return o; };
        for (int f = 0; f < F; ++f) b->frag_name.push_back(put(h->frags[h->frag_order[f]].id));
        for (int c = 0; c < C; ++c) b->cl_name.push_back(put(h->cl_name[h->cl_order[c]]));
        for (int e : h->aln_esp_espid) b->esp_name.push_back(put(h->esp_name.view(e)));
        for (int f : h->sorted_frag) b->sorted_frag.push_back(f0 + f);
    }
    return MBSV_OK;
}
// logprobcov is shared by the batch: one table up to the largest maxSupport (values do not depend on the table length)
static void batch_finish(mbsv_host_batch* b) {
    b->logprobcov.assign((size_t)b->n_exp * (b->max_support + 1), 0.0);
    for (int x = 0; x < b->n_exp; ++x) mbsv_fill_logprobcov(b->lambda[x], b->max_support, b->logprobcov.data() + (size_t)x * (b->max_support + 1));
}

// dst += src (all subproblems of src, in order)
static int batch_merge(mbsv_host_batch* d, const mbsv_host_batch* s) {
    if (s->sub_frag_ptr.size() < 2) return MBSV_OK;
    if (d->sub_frag_ptr.empty()) { *d = *s; return MBSV_OK; }
    if (s->pseq != d->pseq || s->lambda != d->lambda) return hfail(MBSV_ERR_ARG, "hosts of one batch must share the experiments file");
    d->max_support = std::max(d->max_support, s->max_support);
    const int f0 = d->sub_frag_ptr.back(), c0 = d->sub_cluster_ptr.back(), a0 = d->frag_aln_ptr.back(), k0 = d->aln_esp_ptr.back();
    const int l0 = d->cluster_leaf_ptr.back(), r0 = d->cluster_root_ptr.back(), rl0 = d->root_leaf_ptr.back();
    const int h0 = d->cluster_hist_ptr.back(), q0 = d->sv_frag_ptr.back(), v0 = d->sub_sv_ptr.back();
    auto cat = [](std::vector<int32_t>& a, const std::vector<int32_t>& b, int add, size_t skip, bool keep_neg) {
        a.reserve(a.size() + b.size());
        for (size_t i = skip; i < b.size(); ++i) a.push_back(keep_neg && b[i] < 0 ? b[i] : b[i] + add);
    };
    cat(d->sub_frag_ptr, s->sub_frag_ptr, f0, 1, false); cat(d->sub_cluster_ptr, s->sub_cluster_ptr, c0, 1, false);
    cat(d->sub_sv_ptr, s->sub_sv_ptr, v0, 1, false); cat(d->frag_aln_ptr, s->frag_aln_ptr, a0, 1, false);
    cat(d->aln_numerrors, s->aln_numerrors, 0, 0, false); cat(d->aln_len, s->aln_len, 0, 0, false); cat(d->aln_exp, s->aln_exp, 0, 0, false);
    cat(d->aln_esp_ptr, s->aln_esp_ptr, k0, 1, false); cat(d->aln_esp_cluster, s->aln_esp_cluster, c0, 0, true);
    cat(d->aln_esp_leaf, s->aln_esp_leaf, 0, 0, false);
    d->cluster_kind.insert(d->cluster_kind.end(), s->cluster_kind.begin(), s->cluster_kind.end());
    cat(d->supports_order, s->supports_order, c0, 0, false); cat(d->cluster_hist_ptr, s->cluster_hist_ptr, h0, 1, false);
    cat(d->cluster_leaf_ptr, s->cluster_leaf_ptr, l0, 1, false); cat(d->leaf_owner, s->leaf_owner, f0, 0, true);
    cat(d->cluster_root_ptr, s->cluster_root_ptr, r0, 1, false); cat(d->root_leaf_ptr, s->root_leaf_ptr, rl0, 1, false);
    cat(d->root_leaf, s->root_leaf, 0, 0, false); cat(d->sv_cluster, s->sv_cluster, c0, 0, false);
    cat(d->sv_frag_ptr, s->sv_frag_ptr, q0, 1, false); cat(d->sv_frag, s->sv_frag, f0, 0, false);
    cat(d->sv_numchanged, s->sv_numchanged, 0, 0, false);
    cat(d->sub_max_support, s->sub_max_support, 0, 0, false); cat(d->sub_number, s->sub_number, 0, 0, false);
    if (!s->have_names) d->have_names = false;
    if (d->have_names) {
        const uint64_t p0 = d->name_pool.size();
        d->name_pool.insert(d->name_pool.end(), s->name_pool.begin(), s->name_pool.end());
        for (uint64_t o : s->frag_name) d->frag_name.push_back(p0 + o);
        for (uint64_t o : s->cl_name) d->cl_name.push_back(p0 + o);
        for (uint64_t o : s->esp_name) d->esp_name.push_back(p0 + o);
        cat(d->sorted_frag, s->sorted_frag, f0, 0, false);
    }
    return MBSV_OK;
}

int mbsv_host_batch_create(mbsv_host* const* hosts, int32_t n, mbsv_host_batch** out) {
    if (!hosts || n < 1 || !out) return hfail(MBSV_ERR_ARG, "bad argument");
    mbsv_host_batch* b = new (std::nothrow) mbsv_host_batch();
    if (!b) return hfail(MBSV_ERR_OOM, "allocation failed");
    for (int i = 0; i < n; ++i)
        if (const int rc = batch_append(b, hosts[i])) { delete b; return rc; }
    batch_finish(b);
    *out = b;
    return MBSV_OK;
}

// The large-scale workflow in one call: partition the clusters file (generateIndependentSubproblems.pl), then load every
// subset exactly as one reference run on `subset-k.clusters --readclustersfirst` would see it -- without writing the subset
// files and without re-reading the assignments file once per subset (doc/MultiBreakSV-Manual.txt:464-473 does both).
// Every assignments tuple is routed to the subset(s) that list one of its ESPs; each subset then goes through the same
// readers on its own slice of text, so orders, filtering and quirks are those of the per-file path (tested equal).
// Subsets left without any fragment are skipped (n_skipped).  Returns the number of subproblems in the batch.
int mbsv_host_load_partitioned(const char* clusterfile, const char* assignmentfile, const char* experimentfile,
                               int32_t flags, mbsv_host_batch** out, int32_t* info) {
    if (!clusterfile || !assignmentfile || !experimentfile || !out) return hfail(MBSV_ERR_ARG, "bad argument");
    std::string cbuf, abuf, ebuf;
    if (!slurp(clusterfile, cbuf)) return hfail(MBSV_ERR_IO, std::string("cannot open ") + clusterfile);
    if (!slurp(assignmentfile, abuf)) return hfail(MBSV_ERR_IO, std::string("cannot open ") + assignmentfile);
    if (!slurp(experimentfile, ebuf)) return hfail(MBSV_ERR_IO, std::string("cannot open ") + experimentfile);
    Lap lap;
    std::vector<std::string_view> lines;
    std::vector<int> line_subset;
    const int S = partition_lines(cbuf, lines, line_subset);
    lap("partition");
    // subset texts of the clusters file (file order) and ESP -> subset (names as readClustersFile splits them: "," + trim)
    std::vector<std::string> ctext((size_t)S + 1), atext((size_t)S + 1);
    // allesps of every subset, as one table: ESP -> the subset whose allesps holds it (a second, third ... subset goes to
    // `more`; that needs an ESP shared by unlinked rows and does not happen with long-read-named ESPs)
    StrTable esps;
    std::vector<int> esp_subset;
    std::unordered_map<int, std::vector<int>> more;
    auto add_member = [&](std::string_view name, int k) {
        bool added;
        const int id = esps.intern(name, &added);
        if (added) { esp_subset.push_back(k); return; }
        if (esp_subset[id] == k) return;
        std::vector<int>& v = more[id];
        if (std::find(v.begin(), v.end(), k) == v.end()) v.push_back(k);
    };
    std::vector<std::string_view> row, names;
    for (size_t i = 0; i < lines.size(); ++i) {
        const int k = line_subset[i];
        if (k <= 0) continue;
        ctext[k].append(lines[i]); ctext[k].push_back('\n');
        split_views(lines[i], "\t", row);
        if (row.size() < 5) continue;
        split_views(row[4], ",", names);
        for (const auto& raw : names) add_member(trim_view(raw), k);
    }
    lap("subset cluster texts");
    // route the assignments tuples (token-based, like readAssignmentFile :342-373): a tuple is kept by a subset when one of
    // its ESPs is in that subset's allesps -- and a kept tuple adds ALL its ESPs to allesps (:386), so a later tuple can be
    // kept through an ESP no cluster lists; one forward pass in file order reproduces that
    Scanner sc;
    sc.buf.swap(abuf);
    if (!sc.hasNextLine()) return hfail(MBSV_ERR_IO, "empty assignment file");
    const std::string header = std::string(sc.nextLineView()) + "\n";
    std::vector<std::string_view> tuple_esps;
    std::vector<int> targets;
    while (sc.hasNext()) {
        const std::string_view frag = sc.nextView();
        if (!sc.hasNext()) return hfail(MBSV_ERR_IO, "Error parsing line. Check assignmentfile format.");
        const std::string_view list = sc.nextView();
        if (!sc.hasNext()) return hfail(MBSV_ERR_IO, "Error parsing line. Check assignmentfile format.");
        const std::string_view nerr = sc.nextView();
        if (!sc.hasNext()) return hfail(MBSV_ERR_IO, "Error parsing line. Check assignmentfile format.");
        const std::string_view len = sc.nextView();
        if (!sc.hasNext()) return hfail(MBSV_ERR_IO, "Error parsing line. Check assignmentfile format.");
        const std::string_view label = sc.nextView();
        split_views(list, ",", tuple_esps);
        targets.clear();
        auto want = [&](int k) { if (std::find(targets.begin(), targets.end(), k) == targets.end()) targets.push_back(k); };
        for (const auto& e : tuple_esps) {
            const int id = esps.find(e);
            if (id < 0) continue;
            want(esp_subset[id]);
            if (!more.empty()) { auto it = more.find(id); if (it != more.end()) for (int k : it->second) want(k); }
        }
        for (int k : targets) {
            std::string& t = atext[k];
            if (t.empty()) t = header;
            t.append(frag); t.push_back('\t'); t.append(list); t.push_back('\t'); t.append(nerr); t.push_back('\t');
            t.append(len); t.push_back('\t'); t.append(label); t.push_back('\n');
            for (const auto& e : tuple_esps) add_member(e, k);
        }
    }
    lap("route assignments");
    // the subsets are independent: worker threads load contiguous ranges into partial batches, merged in subset order
    mbsv_host_batch* b = new (std::nothrow) mbsv_host_batch();
    if (!b) return hfail(MBSV_ERR_OOM, "allocation failed");
    int nthreads = (int)std::max(1u, std::min({std::thread::hardware_concurrency(), 32u, (unsigned)((S + 255) / 256)}));
    if (const char* env = std::getenv("MBSV_HOST_THREADS")) nthreads = std::max(1, std::min(64, std::atoi(env)));   // (tests)
    std::vector<mbsv_host_batch> part(nthreads);
    std::vector<int> t_skipped(nthreads, 0), t_approx(nthreads, 0), t_rc(nthreads, MBSV_OK);
    std::vector<std::string> t_msg(nthreads);
    // ranges balanced by text size (a few subsets hold most of the data)
    std::vector<int> bound(nthreads + 1, S + 1);
    {
        size_t total = 0;
        for (int k = 1; k <= S; ++k) total += atext[k].size() + ctext[k].size();
        size_t acc = 0;
        int t = 0;
        bound[0] = 1;
        for (int k = 1; k <= S; ++k) {
            acc += atext[k].size() + ctext[k].size();
            while (t + 1 < nthreads && acc * nthreads >= total * (size_t)(t + 1)) bound[++t] = k + 1;
        }
        for (int u = t + 1; u <= nthreads; ++u) bound[u] = S + 1;
    }
    auto worker = [&](int t) {
        mbsv_host* h = new (std::nothrow) mbsv_host();
        if (!h) { t_rc[t] = MBSV_ERR_OOM; t_msg[t] = "allocation failed"; return; }
        for (int k = bound[t]; k < bound[t + 1] && t_rc[t] == MBSV_OK; ++k) {
            if (atext[k].empty()) { ++t_skipped[t]; continue; }
            h->clear();
            h->clusterfile = "subset-" + std::to_string(k) + ".clusters"; h->assignmentfile = assignmentfile; h->experimentfile = experimentfile;
            h->clusters_first = true;
            h->text_clusters = &ctext[k]; h->text_assignments = &atext[k]; h->text_experiments = &ebuf;
            int rc = mbsv_host_load(h);
            if (rc == MBSV_OK && (h->experiments.tree_bin || h->fragments.tree_bin || h->breakpoints.tree_bin)) ++t_approx[t];   // (counted: treeified maps)
            if (rc == MBSV_OK && h->frag_order.empty()) { ++t_skipped[t]; continue; }
            if (rc == MBSV_OK) rc = batch_append(&part[t], h, (flags & 2) != 0, k);
            if (rc != MBSV_OK) { t_rc[t] = rc; t_msg[t] = "subset " + std::to_string(k) + ": " + mbsv_last_error(); }
            std::string().swap(ctext[k]); std::string().swap(atext[k]);
        }
        delete h;
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < nthreads; ++t) th.emplace_back(worker, t);
        worker(0);
        for (auto& x : th) x.join();
    }
    lap("load subsets");
    int skipped = 0, approx = 0, rc = MBSV_OK;
    for (int t = 0; t < nthreads && rc == MBSV_OK; ++t) {
        if (t_rc[t] != MBSV_OK) { rc = hfail(t_rc[t], t_msg[t]); break; }
        skipped += t_skipped[t]; approx += t_approx[t];
        rc = batch_merge(b, &part[t]);
        part[t] = mbsv_host_batch();
    }
    lap("merge");
    if (rc != MBSV_OK) { delete b; return rc; }
    if (b->sub_frag_ptr.size() < 2) { delete b; return hfail(MBSV_ERR_ARG, "no subset has any fragment"); }
    batch_finish(b);
    if (info) { info[0] = skipped; info[1] = approx; }
    *out = b;
    return (int)b->sub_frag_ptr.size() - 1;
}
int mbsv_host_batch_desc(mbsv_host_batch* b, mbsv_problem_desc* d) {
    if (!b || !d) return hfail(MBSV_ERR_ARG, "bad argument");
    std::memset(d, 0, sizeof *d);
    d->abi_version = MBSV_ABI_VERSION;
    d->n_sub = (int)b->sub_frag_ptr.size() - 1;
    d->n_frag = b->sub_frag_ptr.back(); d->n_aln = (int)b->aln_numerrors.size(); d->n_aln_esp = (int)b->aln_esp_cluster.size();
    d->n_cluster = b->sub_cluster_ptr.back(); d->n_exp = b->n_exp; d->max_support = b->max_support;
    d->n_sv = (int)b->sv_cluster.size(); d->n_sv_frag = (int)b->sv_frag.size(); d->n_leaf = (int)b->leaf_owner.size();
    d->n_root = (int)b->root_leaf_ptr.size() - 1; d->n_root_leaf = (int)b->root_leaf.size(); d->n_hist = b->cluster_hist_ptr.back();
    d->sub_frag_ptr = b->sub_frag_ptr.data(); d->sub_cluster_ptr = b->sub_cluster_ptr.data(); d->sub_sv_ptr = b->sub_sv_ptr.data();
    d->frag_aln_ptr = b->frag_aln_ptr.data(); d->aln_numerrors = ptr_or_null(b->aln_numerrors); d->aln_len = ptr_or_null(b->aln_len);
    d->aln_exp = ptr_or_null(b->aln_exp); d->aln_esp_ptr = b->aln_esp_ptr.data(); d->aln_esp_cluster = ptr_or_null(b->aln_esp_cluster);
    d->aln_esp_leaf = ptr_or_null(b->aln_esp_leaf); d->cluster_kind = ptr_or_null(b->cluster_kind);
    d->supports_order = ptr_or_null(b->supports_order); d->cluster_hist_ptr = b->cluster_hist_ptr.data();
    d->cluster_leaf_ptr = b->cluster_leaf_ptr.data(); d->leaf_owner = ptr_or_null(b->leaf_owner);
    d->cluster_root_ptr = b->cluster_root_ptr.data(); d->root_leaf_ptr = b->root_leaf_ptr.data(); d->root_leaf = ptr_or_null(b->root_leaf);
    d->sv_cluster = ptr_or_null(b->sv_cluster); d->sv_frag_ptr = b->sv_frag_ptr.data(); d->sv_frag = ptr_or_null(b->sv_frag);
    d->sv_numchanged = ptr_or_null(b->sv_numchanged); d->exp_pseq = b->pseq.data(); d->exp_lambda = b->lambda.data();
    d->logprobcov = b->logprobcov.data();
    return MBSV_OK;
}
int mbsv_host_batch_destroy(mbsv_host_batch* b) { delete b; return MBSV_OK; }

// writeFinalFiles (:1712-1859) for a whole batch: one pair of files, one header, the rows of every subproblem in batch
// order -- what concatenating the per-subset outputs of the reference workflow gives.  Each subproblem prints its own
// number of Support columns (its maxSupport), so rows are ragged across subproblems, as those per-subset files are.
int mbsv_host_batch_write_final_files(mbsv_host_batch* b, int32_t burnin, const uint32_t* aln_count,
                                      const uint32_t* frag_err_count, const uint32_t* cluster_hist, const char* prefix) {
    if (!b || !aln_count || !frag_err_count || !cluster_hist || !prefix || burnin <= 0) return hfail(MBSV_ERR_ARG, "bad argument");
    if (!b->have_names || b->frag_name.size() != (size_t)b->sub_frag_ptr.back())
        return hfail(MBSV_ERR_ARG, "the batch was built without names (mbsv_host_load_partitioned flags bit 1)");
    const int S = (int)b->sub_frag_ptr.size() - 1;
    const char* pool = b->name_pool.data();
    {
        std::ofstream w(std::string(prefix) + ".finalbreakpoints.longburnin");
        if (!w) return hfail(MBSV_ERR_IO, "cannot write breakpoints file");
        w << "ClusterID\tNumIters\tSupport0\tSupport1\tSupport2...\n";
        for (int s = 0; s < S; ++s)
            for (int c = b->sub_cluster_ptr[s]; c < b->sub_cluster_ptr[s + 1]; ++c) {
                w << (pool + b->cl_name[c]) << "\t" << burnin;
                const uint32_t* hist = cluster_hist + b->cluster_hist_ptr[c];
                const int bins = b->cluster_hist_ptr[c + 1] - b->cluster_hist_ptr[c];
                const int ms = b->sub_max_support[s];
                double tot = 0;
                for (int j = 1; j <= ms; ++j) tot = tot + (j < bins ? (double)hist[j] : 0.0);
                w << "\t" << java_double((double)burnin - tot);
                for (int j = 1; j <= ms; ++j) w << "\t" << java_double(j < bins ? (double)hist[j] : 0.0);
                w << "\n";
            }
    }
    {
        std::ofstream w(std::string(prefix) + ".finalassignment.longburnin");
        if (!w) return hfail(MBSV_ERR_IO, "cannot write assignment file");
        w << "FragmentID\tAssignmentIndex\tAvgAssignment\tDiscordantPairs\n";
        for (size_t k = 0; k < b->sorted_frag.size(); ++k) {
            const int f = b->sorted_frag[k];
            const char* id = pool + b->frag_name[f];
            w << id << "\t-1\t" << java_double((double)frag_err_count[f] / burnin) << "\terror\n";
            for (int a = b->frag_aln_ptr[f]; a < b->frag_aln_ptr[f + 1]; ++a) {
                w << id << "\t" << (a - b->frag_aln_ptr[f]) << "\t" << java_double((double)aln_count[a] / burnin) << "\t";
                for (int j = b->aln_esp_ptr[a]; j < b->aln_esp_ptr[a + 1]; ++j) w << (j > b->aln_esp_ptr[a] ? "," : "") << (pool + b->esp_name[j]);
                w << "\n";
            }
        }
    }
    return MBSV_OK;
}
/* subset number (mbsv_host_load_partitioned) or running index of every subproblem of the batch */
int mbsv_host_batch_numbers(mbsv_host_batch* b, int32_t* out) {
    if (!b || !out) return hfail(MBSV_ERR_ARG, "bad argument");
    for (size_t i = 0; i < b->sub_number.size(); ++i) out[i] = b->sub_number[i];
    return (int)b->sub_number.size();
}

// Connected components of the clusters file as generateIndependentSubproblems.pl finds them: lines = the physical lines
// of buf, line_subset[i] = 1-based subset of line i (0: dropped).  Returns the number of subsets.
static int partition_lines(const std::string& buf, std::vector<std::string_view>& lines, std::vector<int>& line_subset) {
    // physical lines (Perl: while (<CLUST>) { chomp; ... }), as views into buf
    lines.clear();
    for (size_t p = 0; p < buf.size();) {
        size_t q = buf.find('\n', p);
        if (q == std::string::npos) q = buf.size();
        lines.push_back(std::string_view(buf).substr(p, q - p));
        p = q + 1;
    }
    StrTable cids;                                            // cluster ids in order of first appearance
    std::vector<int> cid_last_line;                           // $fulllines{$cid}: the last row of a cID wins
    std::vector<int> cid_first_node;                          // a long read of the first row of the cID
    StrTable lr_node;                                         // long-read id string -> node
    std::vector<long long> lr_value;                          // numeric value for the `<=>` sort
    std::vector<int> parent;
    auto find = [&](int x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; };
    auto node_of = [&](std::string_view lr) {
        bool added;
        const int id = lr_node.intern(lr, &added);
        if (added) {
            parent.push_back(id);
            long long v = 0;                                  // strtoll of a digit string (empty: 0)
            for (char ch : lr) v = v * 10 + (ch - '0');
            lr_value.push_back(v);
        }
        return id;
    };
    std::vector<int> line_node(lines.size(), -1);            // a long read of the row (all of them get united)
    std::vector<int> line_cid(lines.size(), -1);
    std::string last_lr;                                     // Perl's $1 keeps its previous value when the match fails
    std::vector<std::string_view> row;
    for (size_t i = 0; i < lines.size(); ++i) {
        split_views(lines[i], "\t", row);
        const std::string_view cid = row[0];
        if (cid.find('.') != std::string_view::npos) continue;    // maximal clusters are skipped (:52)
        bool added;
        const int ci = cids.intern(cid, &added);
        if (added) { cid_last_line.push_back((int)i); cid_first_node.push_back(-1); } else cid_last_line[ci] = (int)i;
        if (row.size() < 5) continue;
        int first = -1;
        // split(/,\s*/, $row[4])
        const std::string_view list = row[4];
        for (size_t p = 0; p <= list.size();) {
            size_t q = list.find(',', p);
            if (q == std::string_view::npos) q = list.size();
            const std::string_view esp = list.substr(p, q - p);
            p = q + 1;
            while (p < list.size() && (list[p] == ' ' || list[p] == '\t')) ++p;
            const size_t m = esp.find("longread_");
            size_t d = m == std::string_view::npos ? std::string_view::npos : m + 9, e = d;
            while (e != std::string_view::npos && e < esp.size() && esp[e] >= '0' && esp[e] <= '9') ++e;
            if (d != std::string_view::npos && e > d) last_lr.assign(esp.substr(d, e - d));
            const int n = node_of(last_lr);
            if (first < 0) first = n; else parent[find(n)] = find(first);
            if (q == list.size()) break;
        }
        line_node[i] = first;
        line_cid[i] = ci;
        // a repeated cID only keeps its last row, but its long reads were linked through every row it appeared in
        // (the script's %longreads accumulates over all rows while %clusters is overwritten): link rows of one cID
        if (first >= 0) {
            if (cid_first_node[ci] < 0) cid_first_node[ci] = first;
            else parent[find(first)] = find(cid_first_node[ci]);
        }
    }
    // subsets numbered by their numerically smallest long read
    const int L = (int)parent.size();
    std::vector<long long> comp_min(L, -1);
    for (int n = 0; n < L; ++n) {
        const int r = find(n);
        if (comp_min[r] < 0 || lr_value[n] < comp_min[r]) comp_min[r] = lr_value[n];
    }
    std::vector<int> roots;
    for (int n = 0; n < L; ++n) if (find(n) == n) roots.push_back(n);
    std::sort(roots.begin(), roots.end(), [&](int a, int b) { return comp_min[a] < comp_min[b]; });
    std::vector<int> subset_of_root(L, 0);
    for (size_t k = 0; k < roots.size(); ++k) subset_of_root[roots[k]] = (int)k + 1;
    line_subset.assign(lines.size(), 0);
    for (const int i : cid_last_line)
        if (line_node[i] >= 0) line_subset[i] = subset_of_root[find(line_node[i])];
    return (int)roots.size();
}

int mbsv_host_partition(const char* clusterfile, const char* outdir, int32_t* row_subset, int32_t max_rows) {
    if (!clusterfile) return hfail(MBSV_ERR_ARG, "clusters file is required");
    std::string buf;
    if (!slurp(clusterfile, buf)) return hfail(MBSV_ERR_IO, std::string("cannot open ") + clusterfile);
    std::vector<std::string_view> lines;
    std::vector<int> line_subset;
    const int n_subsets = partition_lines(buf, lines, line_subset);
    if (row_subset) for (int i = 0; i < max_rows && i < (int)lines.size(); ++i) row_subset[i] = line_subset[i];
    if (outdir) {
        std::vector<std::ofstream> out;
        // at most a few thousand files are kept open at once: write subset by subset
        std::vector<std::vector<int>> members((size_t)n_subsets + 1);
        for (size_t i = 0; i < lines.size(); ++i) if (line_subset[i] > 0) members[line_subset[i]].push_back((int)i);
        for (size_t k = 1; k < members.size(); ++k) {
            std::ofstream w(std::string(outdir) + "/subset-" + std::to_string(k) + ".clusters");
            if (!w) return hfail(MBSV_ERR_IO, "cannot write subset file");
            for (int i : members[k]) w << lines[i] << "\n";
        }
    }
    return n_subsets;
}

int mbsv_java_double_to_string(double v, char* buf, int32_t buflen) {
    const std::string s = java_double(v);
    if (buf && buflen > 0) { std::strncpy(buf, s.c_str(), buflen - 1); buf[buflen - 1] = 0; }
    return (int)s.size();
}
int mbsv_java_hashmap_order(const char* keys, int32_t n, int32_t* out_order, int32_t* out_capacity) {
    JavaMapOrder m;
    const char* p = keys;
    for (int i = 0; i < n; ++i) { std::string k(p); p += k.size() + 1; m.put(k); }
    int i = 0;
    for (int id : m.order()) out_order[i++] = id;
    if (out_capacity) *out_capacity = m.cap;
    return i;
}
// keySet() order after a sequence of puts and removes: keys = n NUL-terminated strings; ops[k] = id + 1 puts key id (every
// key at most once), -(id + 1) removes it.  out_order receives the ids in iteration order; returns their number.
// flags bit 0: always use the literal emulation (java_hashmap.hpp) instead of the closed form; out_info[0] = 1 if a bin was
// treeified, out_info[1] = final capacity.
int mbsv_java_hashmap_replay(const char* keys, int32_t n, const int32_t* ops, int32_t n_ops, int32_t flags, int32_t* out_order,
                             int32_t* out_info) {
    std::vector<std::string> ks;
    const char* p = keys;
    for (int i = 0; i < n; ++i) { ks.emplace_back(p); p += ks.back().size() + 1; }
    std::vector<int> order;
    int treeified = 0, cap = 0;
    if (flags & 1) {
        mbsv_java::HashMapLiteral m;
        std::vector<int> id_of(n, -1);                        // key id -> node id (insertion index)
        for (int k = 0; k < n_ops; ++k) {
            const int id = ops[k] > 0 ? ops[k] - 1 : -ops[k] - 1;
            if (id < 0 || id >= n) return -1;
            if (ops[k] > 0) id_of[id] = m.put(ks[id]); else m.remove(id_of[id]);
        }
        std::vector<int> key_of(m.nd.size(), -1);
        for (int i = 0; i < n; ++i) if (id_of[i] >= 0) key_of[id_of[i]] = i;
        for (int e : m.order()) order.push_back(key_of[e]);
        for (const auto& nd : m.nd) if (nd.alive && nd.tree) treeified = 1;
        cap = m.capacity();
    } else {
        JavaMapOrder m;
        std::vector<int> id_of(n, -1);
        for (int k = 0; k < n_ops; ++k) {
            const int id = ops[k] > 0 ? ops[k] - 1 : -ops[k] - 1;
            if (id < 0 || id >= n) return -1;
            if (ops[k] > 0) id_of[id] = m.put(ks[id]); else m.remove(id_of[id]);
        }
        std::vector<int> key_of(m.hash.size(), -1);
        for (int i = 0; i < n; ++i) if (id_of[i] >= 0) key_of[id_of[i]] = i;
        for (int e : m.order()) order.push_back(key_of[e]);
        treeified = m.tree_bin;
        cap = m.cap;
    }
    for (size_t i = 0; i < order.size(); ++i) out_order[i] = order[i];
    if (out_info) { out_info[0] = treeified; out_info[1] = cap; }
    return (int)order.size();
}

}  // extern "C"
