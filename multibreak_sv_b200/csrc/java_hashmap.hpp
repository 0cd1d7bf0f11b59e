// Literal emulation of java.util.HashMap<String, ?> (JDK 8 and later) as far as KEY ITERATION ORDER goes, tree bins
// included: putVal / resize / treeifyBin / TreeNode.{treeify, putTreeVal, removeTreeNode, split, untreeify,
// balanceInsertion, balanceDeletion, rotateLeft, rotateRight, moveRootToFront} of the JDK's HashMap.java, node for node.
//
// The reference's trajectory depends on that order (fragments.keySet() at MultiBreakSV.java:999 and :1401,
// breakpoints.keySet() at :750, A.supports.keySet() at MultiBreakSVAssignmentMatrix.java:137).  The closed form in
// mbsv_host.cpp ("by bucket at the final capacity, then by insertion") is exact while every bin is a plain list; a bin
// that collects nine keys at capacity >= 64 becomes a red-black tree whose `next` links are reordered (the tree's root is
// moved to the front of the bin after every insertion and removal, a new node is linked behind its tree parent), so maps
// with such a bin are replayed through this class.
//
// Keys are compared like String.compareTo on byte strings (the inputs are ASCII identifiers); tieBreakOrder is never
// reached because keys are distinct.  There is no JVM in this image: pinned against an independent transliteration of the
// same JDK source in tests/java_hashmap_ref.py.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace mbsv_java {

struct HashMapLiteral {
    static constexpr int TREEIFY_THRESHOLD = 8, UNTREEIFY_THRESHOLD = 6, MIN_TREEIFY_CAPACITY = 64;
    struct Node {
        int32_t hash;                 // spread(String.hashCode()), compared as a signed int inside trees
        int next = -1;
        bool tree = false;            // TreeNode (else plain Node)
        int parent = -1, left = -1, right = -1, prev = -1;
        bool red = false;
        bool alive = true;
    };
    std::vector<Node> nd;
    std::vector<std::string> key;
    std::vector<int> table;           // heads, -1 = empty
    int size = 0, threshold = 0;

    static int32_t string_hash(const std::string& s) {
        uint32_t h = 0;
        for (unsigned char c : s) h = 31u * h + c;
        return (int32_t)h;
    }
    static int32_t spread(int32_t h) { const uint32_t u = (uint32_t)h; return (int32_t)(u ^ (u >> 16)); }
    static int compare_keys(const std::string& a, const std::string& b) {      // String.compareTo
        const size_t n = a.size() < b.size() ? a.size() : b.size();
        for (size_t i = 0; i < n; ++i)
            if (a[i] != b[i]) return (int)(unsigned char)a[i] - (int)(unsigned char)b[i];
        return (int)a.size() - (int)b.size();
    }

    // ---- red-black machinery (HashMap.TreeNode) ----
    int rotate_left(int root, int p) {
        int r, pp, rl;
        if (p != -1 && (r = nd[p].right) != -1) {
            if ((rl = nd[p].right = nd[r].left) != -1) nd[rl].parent = p;
            if ((pp = nd[r].parent = nd[p].parent) == -1) { root = r; nd[r].red = false; }
            else if (nd[pp].left == p) nd[pp].left = r;
            else nd[pp].right = r;
            nd[r].left = p;
            nd[p].parent = r;
        }
        return root;
    }
    int rotate_right(int root, int p) {
        int l, pp, lr;
        if (p != -1 && (l = nd[p].left) != -1) {
            if ((lr = nd[p].left = nd[l].right) != -1) nd[lr].parent = p;
            if ((pp = nd[l].parent = nd[p].parent) == -1) { root = l; nd[l].red = false; }
            else if (nd[pp].right == p) nd[pp].right = l;
            else nd[pp].left = l;
            nd[l].right = p;
            nd[p].parent = l;
        }
        return root;
    }
    int balance_insertion(int root, int x) {
        nd[x].red = true;
        for (int xp, xpp, xppl, xppr;;) {
            if ((xp = nd[x].parent) == -1) { nd[x].red = false; return x; }
            else if (!nd[xp].red || (xpp = nd[xp].parent) == -1) return root;
            if (xp == (xppl = nd[xpp].left)) {
                if ((xppr = nd[xpp].right) != -1 && nd[xppr].red) {
                    nd[xppr].red = false; nd[xp].red = false; nd[xpp].red = true; x = xpp;
                } else {
                    if (x == nd[xp].right) {
                        root = rotate_left(root, x = xp);
                        xpp = (xp = nd[x].parent) == -1 ? -1 : nd[xp].parent;
                    }
                    if (xp != -1) {
                        nd[xp].red = false;
                        if (xpp != -1) { nd[xpp].red = true; root = rotate_right(root, xpp); }
                    }
                }
            } else {
                if (xppl != -1 && nd[xppl].red) {
                    nd[xppl].red = false; nd[xp].red = false; nd[xpp].red = true; x = xpp;
                } else {
                    if (x == nd[xp].left) {
                        root = rotate_right(root, x = xp);
                        xpp = (xp = nd[x].parent) == -1 ? -1 : nd[xp].parent;
                    }
                    if (xp != -1) {
                        nd[xp].red = false;
                        if (xpp != -1) { nd[xpp].red = true; root = rotate_left(root, xpp); }
                    }
                }
            }
        }
    }
    int balance_deletion(int root, int x) {
        for (int xp, xpl, xpr;;) {
            if (x == -1 || x == root) return root;
            else if ((xp = nd[x].parent) == -1) { nd[x].red = false; return x; }
            else if (nd[x].red) { nd[x].red = false; return root; }
            else if ((xpl = nd[xp].left) == x) {
                if ((xpr = nd[xp].right) != -1 && nd[xpr].red) {
                    nd[xpr].red = false; nd[xp].red = true;
                    root = rotate_left(root, xp);
                    xpr = (xp = nd[x].parent) == -1 ? -1 : nd[xp].right;
                }
                if (xpr == -1) x = xp;
                else {
                    int sl = nd[xpr].left, sr = nd[xpr].right;
                    if ((sr == -1 || !nd[sr].red) && (sl == -1 || !nd[sl].red)) { nd[xpr].red = true; x = xp; }
                    else {
                        if (sr == -1 || !nd[sr].red) {
                            if (sl != -1) nd[sl].red = false;
                            nd[xpr].red = true;
                            root = rotate_right(root, xpr);
                            xpr = (xp = nd[x].parent) == -1 ? -1 : nd[xp].right;
                        }
                        if (xpr != -1) {
                            nd[xpr].red = (xp == -1) ? false : nd[xp].red;
                            if ((sr = nd[xpr].right) != -1) nd[sr].red = false;
                        }
                        if (xp != -1) { nd[xp].red = false; root = rotate_left(root, xp); }
                        x = root;
                    }
                }
            } else {   // symmetric
                if (xpl != -1 && nd[xpl].red) {
                    nd[xpl].red = false; nd[xp].red = true;
                    root = rotate_right(root, xp);
                    xpl = (xp = nd[x].parent) == -1 ? -1 : nd[xp].left;
                }
                if (xpl == -1) x = xp;
                else {
                    int sl = nd[xpl].left, sr = nd[xpl].right;
                    if ((sl == -1 || !nd[sl].red) && (sr == -1 || !nd[sr].red)) { nd[xpl].red = true; x = xp; }
                    else {
                        if (sl == -1 || !nd[sl].red) {
                            if (sr != -1) nd[sr].red = false;
                            nd[xpl].red = true;
                            root = rotate_left(root, xpl);
                            xpl = (xp = nd[x].parent) == -1 ? -1 : nd[xp].left;
                        }
                        if (xpl != -1) {
                            nd[xpl].red = (xp == -1) ? false : nd[xp].red;
                            if ((sl = nd[xpl].left) != -1) nd[sl].red = false;
                        }
                        if (xp != -1) { nd[xp].red = false; root = rotate_right(root, xp); }
                        x = root;
                    }
                }
            }
        }
    }
    int root_of(int r) const { for (int p; (p = nd[r].parent) != -1;) r = p; return r; }
    void move_root_to_front(int root) {
        if (root == -1 || table.empty()) return;
        const int index = (int)(((uint32_t)nd[root].hash) & (uint32_t)(table.size() - 1));
        const int first = table[index];
        if (root != first) {
            const int rn = nd[root].next, rp = nd[root].prev;
            table[index] = root;
            if (rn != -1) nd[rn].prev = rp;
            if (rp != -1) nd[rp].next = rn;
            if (first != -1) nd[first].prev = root;
            nd[root].next = first;
            nd[root].prev = -1;
        }
    }
    // dir of inserting key/hash of node x below node p (treeify, putTreeVal): hash as signed int, then String.compareTo
    int dir_of(int x, int p) const {
        const int32_t ph = nd[p].hash, h = nd[x].hash;
        if (ph > h) return -1;
        if (ph < h) return 1;
        const int c = compare_keys(key[x], key[p]);
        return c < 0 ? -1 : 1;                                  // distinct keys: never 0
    }
    void treeify(int head) {                                    // TreeNode.treeify on the `next` list starting at head
        int root = -1;
        for (int x = head, next; x != -1; x = next) {
            next = nd[x].next;
            nd[x].left = nd[x].right = -1;
            if (root == -1) { nd[x].parent = -1; nd[x].red = false; root = x; }
            else {
                for (int p = root;;) {
                    const int dir = dir_of(x, p);
                    const int xp = p;
                    if ((p = (dir <= 0) ? nd[p].left : nd[p].right) == -1) {
                        nd[x].parent = xp;
                        if (dir <= 0) nd[xp].left = x; else nd[xp].right = x;
                        root = balance_insertion(root, x);
                        break;
                    }
                }
            }
        }
        move_root_to_front(root);
    }
    int untreeify(int head) {                                   // plain nodes in `next` order
        for (int q = head; q != -1; q = nd[q].next) { nd[q].tree = false; nd[q].parent = nd[q].left = nd[q].right = nd[q].prev = -1; nd[q].red = false; }
        return head;
    }
    void treeify_bin(int32_t hash) {
        if ((int)table.size() < MIN_TREEIFY_CAPACITY) { resize(); return; }
        const int index = (int)(((uint32_t)hash) & (uint32_t)(table.size() - 1));
        int e = table[index];
        if (e == -1) return;
        int tl = -1;
        for (; e != -1; e = nd[e].next) {
            nd[e].tree = true; nd[e].parent = nd[e].left = nd[e].right = -1; nd[e].red = false;
            nd[e].prev = tl;
            tl = e;
        }
        treeify(table[index]);
    }
    void split(std::vector<int>& ntab, int b, int index, int bit) {
        int loHead = -1, loTail = -1, hiHead = -1, hiTail = -1, lc = 0, hc = 0;
        for (int e = b, next; e != -1; e = next) {
            next = nd[e].next;
            nd[e].next = -1;
            if ((((uint32_t)nd[e].hash) & (uint32_t)bit) == 0) {
                if ((nd[e].prev = loTail) == -1) loHead = e; else nd[loTail].next = e;
                loTail = e; ++lc;
            } else {
                if ((nd[e].prev = hiTail) == -1) hiHead = e; else nd[hiTail].next = e;
                hiTail = e; ++hc;
            }
        }
        // (treeify writes through `table`: the new table is installed before split is called)
        if (loHead != -1) {
            if (lc <= UNTREEIFY_THRESHOLD) ntab[index] = untreeify(loHead);
            else { ntab[index] = loHead; if (hiHead != -1) treeify(loHead); }
        }
        if (hiHead != -1) {
            if (hc <= UNTREEIFY_THRESHOLD) ntab[index + bit] = untreeify(hiHead);
            else { ntab[index + bit] = hiHead; if (loHead != -1) treeify(hiHead); }
        }
    }
    void resize() {
        if (table.empty()) { table.assign(16, -1); threshold = 12; return; }
        const int oldCap = (int)table.size(), newCap = oldCap * 2;
        std::vector<int> old(newCap, -1);
        old.swap(table);                                        // `table` is the new one from here on (as in the JDK)
        threshold *= 2;
        for (int j = 0; j < oldCap; ++j) {
            const int e = old[j];
            if (e == -1) continue;
            if (nd[e].next == -1) table[((uint32_t)nd[e].hash) & (uint32_t)(newCap - 1)] = e;
            else if (nd[e].tree) split(table, e, j, oldCap);
            else {
                int loHead = -1, loTail = -1, hiHead = -1, hiTail = -1;
                for (int q = e, next; q != -1; q = next) {
                    next = nd[q].next;
                    if ((((uint32_t)nd[q].hash) & (uint32_t)oldCap) == 0) { if (loTail == -1) loHead = q; else nd[loTail].next = q; loTail = q; }
                    else { if (hiTail == -1) hiHead = q; else nd[hiTail].next = q; hiTail = q; }
                }
                if (loTail != -1) { nd[loTail].next = -1; table[j] = loHead; }
                if (hiTail != -1) { nd[hiTail].next = -1; table[j + oldCap] = hiHead; }
            }
        }
    }
    // put of a NEW key; returns its node id (= insertion index)
    int put(const std::string& k) {
        if (table.empty()) resize();
        const int32_t h = spread(string_hash(k));
        const int id = (int)nd.size();
        Node n; n.hash = h;
        nd.push_back(n);
        key.push_back(k);
        const int i = (int)(((uint32_t)h) & (uint32_t)(table.size() - 1));
        int p = table[i];
        if (p == -1) table[i] = id;
        else if (nd[p].tree) {
            // putTreeVal: descend from the root, link the new node behind its tree parent in the `next` list
            nd[id].tree = true;
            const int root = nd[p].parent != -1 ? root_of(p) : p;
            for (int q = root;;) {
                const int dir = dir_of(id, q);
                const int xp = q;
                if ((q = (dir <= 0) ? nd[q].left : nd[q].right) == -1) {
                    const int xpn = nd[xp].next;
                    nd[id].next = xpn;
                    if (dir <= 0) nd[xp].left = id; else nd[xp].right = id;
                    nd[xp].next = id;
                    nd[id].parent = nd[id].prev = xp;
                    if (xpn != -1) nd[xpn].prev = id;
                    move_root_to_front(balance_insertion(root, id));
                    break;
                }
            }
        } else {
            for (int binCount = 0;; ++binCount) {
                const int e = nd[p].next;
                if (e == -1) {
                    nd[p].next = id;
                    if (binCount >= TREEIFY_THRESHOLD - 1) treeify_bin(h);
                    break;
                }
                p = e;
            }
        }
        if (++size > threshold) resize();
        return id;
    }
    void remove(int id) {                                        // HashMap.remove(key): removeNode(..., movable = true)
        if (id < 0 || id >= (int)nd.size() || !nd[id].alive || table.empty()) return;
        nd[id].alive = false;
        --size;
        const int index = (int)(((uint32_t)nd[id].hash) & (uint32_t)(table.size() - 1));
        if (!nd[id].tree) {
            int p = table[index];
            if (p == id) table[index] = nd[id].next;
            else { while (nd[p].next != id) p = nd[p].next; nd[p].next = nd[id].next; }
            return;
        }
        // ---- TreeNode.removeTreeNode(map, tab, movable = true) ----
        int first = table[index], root = first, rl;
        const int succ = nd[id].next, pred = nd[id].prev;
        if (pred == -1) table[index] = first = succ; else nd[pred].next = succ;
        if (succ != -1) nd[succ].prev = pred;
        if (first == -1) return;
        if (nd[root].parent != -1) root = root_of(root);
        if (root == -1 || nd[root].right == -1 || (rl = nd[root].left) == -1 || nd[rl].left == -1) {
            table[index] = untreeify(first);                    // too small
            return;
        }
        const int p = id, pl = nd[p].left, pr = nd[p].right;
        int replacement;
        if (pl != -1 && pr != -1) {
            int s = pr, sl;
            while ((sl = nd[s].left) != -1) s = sl;             // successor
            const bool c = nd[s].red; nd[s].red = nd[p].red; nd[p].red = c;
            const int sr = nd[s].right, pp = nd[p].parent;
            if (s == pr) { nd[p].parent = s; nd[s].right = p; }
            else {
                const int sp = nd[s].parent;
                if ((nd[p].parent = sp) != -1) { if (s == nd[sp].left) nd[sp].left = p; else nd[sp].right = p; }
                if ((nd[s].right = pr) != -1) nd[pr].parent = s;
            }
            nd[p].left = -1;
            if ((nd[p].right = sr) != -1) nd[sr].parent = p;
            if ((nd[s].left = pl) != -1) nd[pl].parent = s;
            if ((nd[s].parent = pp) == -1) root = s;
            else if (p == nd[pp].left) nd[pp].left = s;
            else nd[pp].right = s;
            replacement = sr != -1 ? sr : p;
        } else if (pl != -1) replacement = pl;
        else if (pr != -1) replacement = pr;
        else replacement = p;
        if (replacement != p) {
            const int pp = nd[replacement].parent = nd[p].parent;
            if (pp == -1) { root = replacement; nd[root].red = false; }
            else if (p == nd[pp].left) nd[pp].left = replacement;
            else nd[pp].right = replacement;
            nd[p].left = nd[p].right = nd[p].parent = -1;
        }
        const int r = nd[p].red ? root : balance_deletion(root, replacement);
        if (replacement == p) {                                 // detach
            const int pp = nd[p].parent;
            nd[p].parent = -1;
            if (pp != -1) { if (p == nd[pp].left) nd[pp].left = -1; else if (p == nd[pp].right) nd[pp].right = -1; }
        }
        move_root_to_front(r);
    }
    std::vector<int> order() const {                            // keySet() iteration: bins ascending, `next` links
        std::vector<int> o;
        o.reserve(size);
        for (int head : table) for (int e = head; e != -1; e = nd[e].next) o.push_back(e);
        return o;
    }
    int capacity() const { return (int)table.size(); }
};

}  // namespace mbsv_java
