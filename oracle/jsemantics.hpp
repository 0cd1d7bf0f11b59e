// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under multibreak_sv_b200/ may include, link or call
// this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
//
// JDK semantics the reference sampler leans on, restated on the CPU.  The JDK is a third-party
// dependency that is absent from /root/reference (build.xml pins no version; the manual says "tested on
// version 1.7", north_star asks for Panama FFM => JDK 22+).  We restate the *Java 8+* behaviour:
//   * java.util.Random          — bit-specified by its Javadoc (48-bit LCG)
//   * StrictMath.log / exp      — fdlibm 5.3 e_log.c / e_exp.c algorithms (Java's StrictMath contract)
//   * java.util.HashMap<String> — key iteration order (table doubling, tail insertion, order-preserving
//                                 split on resize, treeifyBin()'s resize below capacity 64)
//   * String.hashCode, String.split, String.trim, Scanner tokenisation
// Call sites in the reference: MultiBreakSV.java:78,1008,1013,1106,1115,1148,1198,1248,1400,1406 (Random),
// :1146 (exp), :1219-1271,1342-1343,1645-1693 (log), MultiBreakSVAssignmentMatrix.java:223 (log).
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

namespace orc {

// ------------------------------------------------------------------------------------------------
// fdlibm 5.3 log/exp (the algorithm Java's StrictMath specifies).  Compiled with -ffp-contract=off.
// ------------------------------------------------------------------------------------------------
static inline uint64_t d2u(double d) { uint64_t u; std::memcpy(&u, &d, 8); return u; }
static inline double u2d(uint64_t u) { double d; std::memcpy(&d, &u, 8); return d; }
static inline int32_t hi_word(double d) { return (int32_t)(d2u(d) >> 32); }
static inline uint32_t lo_word(double d) { return (uint32_t)d2u(d); }
static inline double with_hi(double d, int32_t hi) { return u2d(((uint64_t)(uint32_t)hi << 32) | lo_word(d)); }

static inline double j_log(double x) {
    const double ln2_hi = u2d(0x3fe62e42fee00000ULL), ln2_lo = u2d(0x3dea39ef35793c76ULL);
    const double two54 = u2d(0x4350000000000000ULL);
    const double Lg1 = u2d(0x3FE5555555555593ULL), Lg2 = u2d(0x3FD999999997FA04ULL),
                 Lg3 = u2d(0x3FD2492494229359ULL), Lg4 = u2d(0x3FCC71C51D8E78AFULL),
                 Lg5 = u2d(0x3FC7466496CB03DEULL), Lg6 = u2d(0x3FC39A09D078C69FULL),
                 Lg7 = u2d(0x3FC2F112DF3E5244ULL);
    const double zero = 0.0;
    int32_t hx = hi_word(x);
    uint32_t lx = lo_word(x);
    int32_t k = 0;
    if (hx < 0x00100000) {                       // x < 2**-1022
        if (((hx & 0x7fffffff) | lx) == 0) return -two54 / zero;  // log(+-0) = -inf
        if (hx < 0) return (x - x) / zero;       // log(-#) = NaN
        k -= 54; x *= two54; hx = hi_word(x);    // subnormal, scale up
    }
    if (hx >= 0x7ff00000) return x + x;
    k += (hx >> 20) - 1023;
    hx &= 0x000fffff;
    int32_t i = (hx + 0x95f64) & 0x100000;
    x = with_hi(x, hx | (i ^ 0x3ff00000));       // normalize x or x/2
    k += (i >> 20);
    double f = x - 1.0;
    double dk, R;
    if ((0x000fffff & (2 + hx)) < 3) {           // |f| < 2**-20
        if (f == zero) {
            if (k == 0) return zero;
            dk = (double)k; return dk * ln2_hi + dk * ln2_lo;
        }
        R = f * f * (0.5 - 0.33333333333333333 * f);
        if (k == 0) return f - R;
        dk = (double)k; return dk * ln2_hi - ((R - dk * ln2_lo) - f);
    }
    double s = f / (2.0 + f);
    dk = (double)k;
    double z = s * s;
    i = hx - 0x6147a;
    double w = z * z;
    int32_t j = 0x6b851 - hx;
    double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    i |= j;
    R = t2 + t1;
    if (i > 0) {
        double hfsq = 0.5 * f * f;
        if (k == 0) return f - (hfsq - s * (hfsq + R));
        return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
    } else {
        if (k == 0) return f - s * (f - R);
        return dk * ln2_hi - ((s * (f - R) - dk * ln2_lo) - f);
    }
}

static inline double j_exp(double x) {
    const double one = 1.0, huge = 1.0e+300;
    const double halF[2] = {0.5, -0.5};
    const double twom1000 = u2d(0x0170000000000000ULL);          // 2**-1000
    const double o_threshold = u2d(0x40862E42FEFA39EFULL);
    const double u_threshold = u2d(0xc0874910D52D3051ULL);
    const double ln2HI[2] = {u2d(0x3fe62e42fee00000ULL), u2d(0xbfe62e42fee00000ULL)};
    const double ln2LO[2] = {u2d(0x3dea39ef35793c76ULL), u2d(0xbdea39ef35793c76ULL)};
    const double invln2 = u2d(0x3ff71547652b82feULL);
    const double P1 = u2d(0x3FC555555555553EULL), P2 = u2d(0xBF66C16C16BEBD93ULL),
                 P3 = u2d(0x3F11566AAF25DE2CULL), P4 = u2d(0xBEBBBD41C5D26BF1ULL),
                 P5 = u2d(0x3E66376972BEA4D0ULL);
    double y, hi = 0.0, lo = 0.0, c, t;
    int32_t k = 0, xsb;
    uint32_t hx = (uint32_t)hi_word(x);
    xsb = (int32_t)((hx >> 31) & 1);
    hx &= 0x7fffffff;
    if (hx >= 0x40862E42) {                       // |x| >= 709.78...
        if (hx >= 0x7ff00000) {
            if (((hx & 0xfffff) | lo_word(x)) != 0) return x + x;  // NaN
            return (xsb == 0) ? x : 0.0;          // exp(+-inf) = {inf,0}
        }
        if (x > o_threshold) return huge * huge;
        if (x < u_threshold) return twom1000 * twom1000;
    }
    if (hx > 0x3fd62e42) {                        // |x| > 0.5 ln2
        if (hx < 0x3FF0A2B2) {                    // and |x| < 1.5 ln2
            hi = x - ln2HI[xsb]; lo = ln2LO[xsb]; k = 1 - xsb - xsb;
        } else {
            k = (int32_t)(invln2 * x + halF[xsb]);
            t = k;
            hi = x - t * ln2HI[0];
            lo = t * ln2LO[0];
        }
        x = hi - lo;
    } else if (hx < 0x3e300000) {                 // |x| < 2**-28
        if (huge + x > one) return one + x;
    } else k = 0;
    t = x * x;
    c = x - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    if (k == 0) return one - ((x * c) / (c - 2.0) - x);
    y = one - ((lo - (x * c) / (2.0 - c)) - hi);
    if (k >= -1021) return with_hi(y, hi_word(y) + (k << 20));
    y = with_hi(y, hi_word(y) + ((k + 1000) << 20));
    return y * twom1000;
}

// ------------------------------------------------------------------------------------------------
// java.util.Random
// ------------------------------------------------------------------------------------------------
struct JRandom {
    static constexpr uint64_t MULT = 0x5DEECE66DULL, ADD = 0xBULL, MASK = (1ULL << 48) - 1;
    uint64_t s;
    long long ncalls = 0;  // number of next() steps, for tests
    explicit JRandom(int64_t seed) : s(((uint64_t)seed ^ MULT) & MASK) {}
    inline int32_t next(int bits) {
        s = (s * MULT + ADD) & MASK;
        ++ncalls;
        return (int32_t)(uint32_t)(s >> (48 - bits));   // Java: (int)(seed >>> (48 - bits))
    }
    inline int32_t nextInt() { return next(32); }
    inline double nextDouble() {
        int64_t a = (int64_t)next(26);
        int64_t b = (int64_t)next(27);
        return (double)((a << 27) + b) * 0x1.0p-53;
    }
    inline int32_t nextInt(int32_t bound) {
        int32_t r = next(31);
        int32_t m = bound - 1;
        if ((bound & m) == 0) {
            r = (int32_t)(((int64_t)bound * (int64_t)r) >> 31);
        } else {
            // for (int u = r; u - (r = u % bound) + m < 0; u = next(31));   (32-bit wrap-around)
            int32_t u = r;
            for (;;) {
                r = u % bound;
                int32_t t = (int32_t)((uint32_t)u - (uint32_t)r + (uint32_t)m);
                if (t < 0) u = next(31); else break;
            }
        }
        return r;
    }
};

// ------------------------------------------------------------------------------------------------
// String helpers
// ------------------------------------------------------------------------------------------------
static inline int32_t java_string_hash(const std::string& s) {
    uint32_t h = 0;
    for (unsigned char c : s) h = 31u * h + (uint32_t)c;   // ASCII assumed (UTF-16 unit == byte)
    return (int32_t)h;
}

// String.split(literal): no limit => trailing empty strings removed; a no-match returns {s}.
static inline std::vector<std::string> java_split(const std::string& s, const std::string& delim) {
    std::vector<std::string> out;
    size_t pos = 0;
    bool matched = false;
    for (;;) {
        size_t q = s.find(delim, pos);
        if (q == std::string::npos) break;
        matched = true;
        out.push_back(s.substr(pos, q - pos));
        pos = q + delim.size();
    }
    if (!matched) { out.push_back(s); return out; }
    out.push_back(s.substr(pos));
    while (!out.empty() && out.back().empty()) out.pop_back();
    return out;
}

static inline std::string java_trim(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && (unsigned char)s[a] <= ' ') ++a;
    while (b > a && (unsigned char)s[b - 1] <= ' ') --b;
    return s.substr(a, b - a);
}

static inline bool java_is_ws(unsigned char c) {
    // Character.isWhitespace over ASCII: \t \n \v \f \r, FS GS RS US, space
    return c == ' ' || (c >= 9 && c <= 13) || (c >= 28 && c <= 31);
}

// A Scanner over a whole file held in memory: next() tokens, nextLine(), hasNext().
struct JScanner {
    std::string buf;
    size_t pos = 0;
    bool hasNext() const {
        size_t p = pos;
        while (p < buf.size() && java_is_ws((unsigned char)buf[p])) ++p;
        return p < buf.size();
    }
    bool hasNextLine() const { return pos < buf.size(); }
    std::string next() {
        while (pos < buf.size() && java_is_ws((unsigned char)buf[pos])) ++pos;
        size_t a = pos;
        while (pos < buf.size() && !java_is_ws((unsigned char)buf[pos])) ++pos;
        return buf.substr(a, pos - a);
    }
    std::string nextLine() {
        size_t a = pos;
        while (pos < buf.size() && buf[pos] != '\n' && buf[pos] != '\r') ++pos;
        std::string line = buf.substr(a, pos - a);
        if (pos < buf.size()) {
            if (buf[pos] == '\r' && pos + 1 < buf.size() && buf[pos + 1] == '\n') pos += 2;
            else pos += 1;
        }
        return line;
    }
};

// ------------------------------------------------------------------------------------------------
// java.util.HashMap<String,?> key-iteration order (Java 8+), literal bucket emulation.
// ------------------------------------------------------------------------------------------------
struct JHashOrder {
    struct Node { std::string key; uint32_t h; bool alive; };
    std::vector<Node> nodes;
    std::vector<std::vector<int>> table;
    std::unordered_map<std::string, int> index;
    int size = 0, threshold = 0;
    bool tree_bin = false;   // a bin would have been treeified: iteration order no longer emulated

    static uint32_t spread(int32_t h) { uint32_t u = (uint32_t)h; return u ^ (u >> 16); }
    int capacity() const { return (int)table.size(); }
    void resize() {
        if (table.empty()) { table.assign(16, {}); threshold = 12; return; }
        size_t ncap = table.size() * 2;
        std::vector<std::vector<int>> nt(ncap);
        for (size_t b = 0; b < table.size(); ++b)
            for (int id : table[b]) nt[nodes[id].h & (ncap - 1)].push_back(id);  // lo/hi split keeps order
        table.swap(nt);
        threshold *= 2;
    }
    int find(const std::string& k) const {
        auto it = index.find(k);
        return it == index.end() ? -1 : it->second;
    }
    bool contains(const std::string& k) const { return find(k) >= 0; }
    // put: returns node id (stable handle).  Existing key => same id, order unchanged.
    int put(const std::string& k) {
        int id = find(k);
        if (id >= 0) return id;
        if (table.empty()) resize();
        uint32_t h = spread(java_string_hash(k));
        id = (int)nodes.size();
        nodes.push_back({k, h, true});
        index.emplace(k, id);
        std::vector<int>& bin = table[h & (table.size() - 1)];
        size_t before = bin.size();
        bin.push_back(id);
        if (before >= 8) {                       // binCount >= TREEIFY_THRESHOLD-1 => treeifyBin()
            if (table.size() < 64) resize();     // MIN_TREEIFY_CAPACITY: resize instead of treeify
            else tree_bin = true;
        }
        if (++size > threshold) resize();
        return id;
    }
    void remove(const std::string& k) {
        int id = find(k);
        if (id < 0) return;
        std::vector<int>& bin = table[nodes[id].h & (table.size() - 1)];
        for (size_t i = 0; i < bin.size(); ++i)
            if (bin[i] == id) { bin.erase(bin.begin() + i); break; }
        nodes[id].alive = false;
        index.erase(k);
        --size;
    }
    std::vector<int> order() const {             // node ids in keySet() iteration order
        std::vector<int> o;
        o.reserve(size);
        for (const auto& bin : table) for (int id : bin) o.push_back(id);
        return o;
    }
    const std::string& key(int id) const { return nodes[id].key; }
};

}  // namespace orc
