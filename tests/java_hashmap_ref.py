"""java.util.HashMap<String, ?> (JDK 8 and later), transliterated from the JDK's HashMap.java into Python objects: putVal,
resize, treeifyBin, removeNode and TreeNode.{treeify, putTreeVal, removeTreeNode, split, untreeify, balanceInsertion,
balanceDeletion, rotateLeft, rotateRight, moveRootToFront}.  Test infrastructure: an independent second statement of the
algorithm that multibreak_sv_b200/csrc/java_hashmap.hpp implements with index arrays (tests/test_hashmap_tree_bins.py
compares their iteration orders on forged collisions).  No JVM exists in this image, so neither is checked against Java."""
TREEIFY_THRESHOLD, UNTREEIFY_THRESHOLD, MIN_TREEIFY_CAPACITY = 8, 6, 64


def string_hash(s: str) -> int:
    h = 0
    for ch in s.encode():
        h = (31 * h + ch) & 0xFFFFFFFF
    return h - (1 << 32) if h & 0x80000000 else h


def spread(h: int) -> int:
    u = h & 0xFFFFFFFF
    u ^= u >> 16
    return u - (1 << 32) if u & 0x80000000 else u


class Node:
    __slots__ = ("hash", "key", "next", "tree", "parent", "left", "right", "prev", "red")

    def __init__(self, h, key, nxt=None):
        self.hash, self.key, self.next = h, key, nxt
        self.tree = False
        self.parent = self.left = self.right = self.prev = None
        self.red = False


def compare(a: str, b: str) -> int:                       # String.compareTo
    ab, bb = a.encode(), b.encode()
    for x, y in zip(ab, bb):
        if x != y:
            return x - y
    return len(ab) - len(bb)


class JavaHashMap:
    def __init__(self):
        self.table = None
        self.size = 0
        self.threshold = 0

    # ---- TreeNode statics ----
    @staticmethod
    def rotate_left(root, p):
        if p is not None and p.right is not None:
            r = p.right
            rl = p.right = r.left
            if rl is not None:
                rl.parent = p
            pp = r.parent = p.parent
            if pp is None:
                root = r
                r.red = False
            elif pp.left is p:
                pp.left = r
            else:
                pp.right = r
            r.left = p
            p.parent = r
        return root

    @staticmethod
    def rotate_right(root, p):
        if p is not None and p.left is not None:
            l = p.left
            lr = p.left = l.right
            if lr is not None:
                lr.parent = p
            pp = l.parent = p.parent
            if pp is None:
                root = l
                l.red = False
            elif pp.right is p:
                pp.right = l
            else:
                pp.left = l
            l.right = p
            p.parent = l
        return root

    @classmethod
    def balance_insertion(cls, root, x):
        x.red = True
        while True:
            xp = x.parent
            if xp is None:
                x.red = False
                return x
            if not xp.red or xp.parent is None:
                return root
            xpp = xp.parent
            xppl = xpp.left
            if xp is xppl:
                xppr = xpp.right
                if xppr is not None and xppr.red:
                    xppr.red = False
                    xp.red = False
                    xpp.red = True
                    x = xpp
                else:
                    if x is xp.right:
                        x = xp
                        root = cls.rotate_left(root, x)
                        xp = x.parent
                        xpp = None if xp is None else xp.parent
                    if xp is not None:
                        xp.red = False
                        if xpp is not None:
                            xpp.red = True
                            root = cls.rotate_right(root, xpp)
            else:
                if xppl is not None and xppl.red:
                    xppl.red = False
                    xp.red = False
                    xpp.red = True
                    x = xpp
                else:
                    if x is xp.left:
                        x = xp
                        root = cls.rotate_right(root, x)
                        xp = x.parent
                        xpp = None if xp is None else xp.parent
                    if xp is not None:
                        xp.red = False
                        if xpp is not None:
                            xpp.red = True
                            root = cls.rotate_left(root, xpp)

    @classmethod
    def balance_deletion(cls, root, x):
        while True:
            if x is None or x is root:
                return root
            xp = x.parent
            if xp is None:
                x.red = False
                return x
            if x.red:
                x.red = False
                return root
            xpl = xp.left
            if xpl is x:
                xpr = xp.right
                if xpr is not None and xpr.red:
                    xpr.red = False
                    xp.red = True
                    root = cls.rotate_left(root, xp)
                    xp = x.parent
                    xpr = None if xp is None else xp.right
                if xpr is None:
                    x = xp
                else:
                    sl, sr = xpr.left, xpr.right
                    if (sr is None or not sr.red) and (sl is None or not sl.red):
                        xpr.red = True
                        x = xp
                    else:
                        if sr is None or not sr.red:
                            if sl is not None:
                                sl.red = False
                            xpr.red = True
                            root = cls.rotate_right(root, xpr)
                            xp = x.parent
                            xpr = None if xp is None else xp.right
                        if xpr is not None:
                            xpr.red = False if xp is None else xp.red
                            sr = xpr.right
                            if sr is not None:
                                sr.red = False
                        if xp is not None:
                            xp.red = False
                            root = cls.rotate_left(root, xp)
                        x = root
            else:
                if xpl is not None and xpl.red:
                    xpl.red = False
                    xp.red = True
                    root = cls.rotate_right(root, xp)
                    xp = x.parent
                    xpl = None if xp is None else xp.left
                if xpl is None:
                    x = xp
                else:
                    sl, sr = xpl.left, xpl.right
                    if (sl is None or not sl.red) and (sr is None or not sr.red):
                        xpl.red = True
                        x = xp
                    else:
                        if sl is None or not sl.red:
                            if sr is not None:
                                sr.red = False
                            xpl.red = True
                            root = cls.rotate_left(root, xpl)
                            xp = x.parent
                            xpl = None if xp is None else xp.left
                        if xpl is not None:
                            xpl.red = False if xp is None else xp.red
                            sl = xpl.left
                            if sl is not None:
                                sl.red = False
                        if xp is not None:
                            xp.red = False
                            root = cls.rotate_right(root, xp)
                        x = root

    @staticmethod
    def root_of(r):
        while r.parent is not None:
            r = r.parent
        return r

    def move_root_to_front(self, tab, root):
        if root is None or not tab:
            return
        index = (len(tab) - 1) & (root.hash & 0xFFFFFFFF)
        first = tab[index]
        if root is not first:
            tab[index] = root
            rp, rn = root.prev, root.next
            if rn is not None:
                rn.prev = rp
            if rp is not None:
                rp.next = rn
            if first is not None:
                first.prev = root
            root.next = first
            root.prev = None

    @staticmethod
    def direction(h, k, p):
        if p.hash > h:
            return -1
        if p.hash < h:
            return 1
        return -1 if compare(k, p.key) < 0 else 1          # distinct keys

    def treeify(self, tab, head):
        root = None
        x = head
        while x is not None:
            nxt = x.next
            x.left = x.right = None
            if root is None:
                x.parent = None
                x.red = False
                root = x
            else:
                p = root
                while True:
                    d = self.direction(x.hash, x.key, p)
                    xp = p
                    p = p.left if d <= 0 else p.right
                    if p is None:
                        x.parent = xp
                        if d <= 0:
                            xp.left = x
                        else:
                            xp.right = x
                        root = self.balance_insertion(root, x)
                        break
            x = nxt
        self.move_root_to_front(tab, root)

    @staticmethod
    def untreeify(head):
        hd = tl = None
        q = head
        while q is not None:
            p = Node(q.hash, q.key)
            if tl is None:
                hd = p
            else:
                tl.next = p
            tl = p
            q = q.next
        return hd

    def treeify_bin(self, h):
        tab = self.table
        if tab is None or len(tab) < MIN_TREEIFY_CAPACITY:
            self.resize()
            return
        index = (len(tab) - 1) & (h & 0xFFFFFFFF)
        e = tab[index]
        if e is None:
            return
        hd = tl = None
        while e is not None:
            p = Node(e.hash, e.key)
            p.tree = True
            if tl is None:
                hd = p
            else:
                p.prev = tl
                tl.next = p
            tl = p
            e = e.next
        tab[index] = hd
        self.treeify(tab, hd)

    def split(self, tab, b, index, bit):
        lo_head = lo_tail = hi_head = hi_tail = None
        lc = hc = 0
        e = b
        while e is not None:
            nxt = e.next
            e.next = None
            if (e.hash & bit) == 0:
                e.prev = lo_tail
                if lo_tail is None:
                    lo_head = e
                else:
                    lo_tail.next = e
                lo_tail = e
                lc += 1
            else:
                e.prev = hi_tail
                if hi_tail is None:
                    hi_head = e
                else:
                    hi_tail.next = e
                hi_tail = e
                hc += 1
            e = nxt
        if lo_head is not None:
            if lc <= UNTREEIFY_THRESHOLD:
                tab[index] = self.untreeify(lo_head)
            else:
                tab[index] = lo_head
                if hi_head is not None:
                    self.treeify(tab, lo_head)
        if hi_head is not None:
            if hc <= UNTREEIFY_THRESHOLD:
                tab[index + bit] = self.untreeify(hi_head)
            else:
                tab[index + bit] = hi_head
                if lo_head is not None:
                    self.treeify(tab, hi_head)

    def resize(self):
        old = self.table
        if old is None:
            self.table = [None] * 16
            self.threshold = 12
            return
        old_cap = len(old)
        new_cap = old_cap * 2
        self.threshold *= 2
        tab = self.table = [None] * new_cap
        for j in range(old_cap):
            e = old[j]
            if e is None:
                continue
            old[j] = None
            if e.next is None:
                tab[(e.hash & 0xFFFFFFFF) & (new_cap - 1)] = e
            elif e.tree:
                self.split(tab, e, j, old_cap)
            else:
                lo_head = lo_tail = hi_head = hi_tail = None
                while e is not None:
                    nxt = e.next
                    if (e.hash & old_cap) == 0:
                        if lo_tail is None:
                            lo_head = e
                        else:
                            lo_tail.next = e
                        lo_tail = e
                    else:
                        if hi_tail is None:
                            hi_head = e
                        else:
                            hi_tail.next = e
                        hi_tail = e
                    e = nxt
                if lo_tail is not None:
                    lo_tail.next = None
                    tab[j] = lo_head
                if hi_tail is not None:
                    hi_tail.next = None
                    tab[j + old_cap] = hi_head

    def put(self, key: str):                              # a NEW key
        if self.table is None:
            self.resize()
        h = spread(string_hash(key))
        tab = self.table
        i = (len(tab) - 1) & (h & 0xFFFFFFFF)
        p = tab[i]
        if p is None:
            tab[i] = Node(h, key)
        elif p.tree:
            root = self.root_of(p) if p.parent is not None else p
            q = root
            while True:
                d = self.direction(h, key, q)
                xp = q
                q = q.left if d <= 0 else q.right
                if q is None:
                    xpn = xp.next
                    x = Node(h, key, xpn)
                    x.tree = True
                    if d <= 0:
                        xp.left = x
                    else:
                        xp.right = x
                    xp.next = x
                    x.parent = x.prev = xp
                    if xpn is not None:
                        xpn.prev = x
                    self.move_root_to_front(tab, self.balance_insertion(root, x))
                    break
        else:
            bin_count = 0
            while True:
                e = p.next
                if e is None:
                    p.next = Node(h, key)
                    if bin_count >= TREEIFY_THRESHOLD - 1:
                        self.treeify_bin(h)
                    break
                p = e
                bin_count += 1
        self.size += 1
        if self.size > self.threshold:
            self.resize()

    def remove(self, key: str):                           # HashMap.remove(key): removeNode(..., movable = true)
        tab = self.table
        h = spread(string_hash(key))
        index = (len(tab) - 1) & (h & 0xFFFFFFFF)
        p = tab[index]
        node, prev = None, None
        e = p
        while e is not None:
            if e.hash == h and e.key == key:
                node = e
                break
            prev = e
            e = e.next
        if node is None:
            return
        self.size -= 1
        if not node.tree:
            if node is p:
                tab[index] = node.next
            else:
                prev.next = node.next
            return
        self.remove_tree_node(tab, node)

    def remove_tree_node(self, tab, me):
        index = (len(tab) - 1) & (me.hash & 0xFFFFFFFF)
        first = tab[index]
        root = first
        succ, pred = me.next, me.prev
        if pred is None:
            tab[index] = first = succ
        else:
            pred.next = succ
        if succ is not None:
            succ.prev = pred
        if first is None:
            return
        if root.parent is not None:
            root = self.root_of(root)
        if root is None or root.right is None or root.left is None or root.left.left is None:
            tab[index] = self.untreeify(first)
            return
        p, pl, pr = me, me.left, me.right
        if pl is not None and pr is not None:
            s = pr
            while s.left is not None:
                s = s.left
            s.red, p.red = p.red, s.red
            sr = s.right
            pp = p.parent
            if s is pr:
                p.parent = s
                s.right = p
            else:
                sp = s.parent
                p.parent = sp
                if sp is not None:
                    if s is sp.left:
                        sp.left = p
                    else:
                        sp.right = p
                s.right = pr
                if pr is not None:
                    pr.parent = s
            p.left = None
            p.right = sr
            if sr is not None:
                sr.parent = p
            s.left = pl
            if pl is not None:
                pl.parent = s
            s.parent = pp
            if pp is None:
                root = s
            elif p is pp.left:
                pp.left = s
            else:
                pp.right = s
            replacement = sr if sr is not None else p
        elif pl is not None:
            replacement = pl
        elif pr is not None:
            replacement = pr
        else:
            replacement = p
        if replacement is not p:
            pp = replacement.parent = p.parent
            if pp is None:
                root = replacement
                root.red = False
            elif p is pp.left:
                pp.left = replacement
            else:
                pp.right = replacement
            p.left = p.right = p.parent = None
        r = root if p.red else self.balance_deletion(root, replacement)
        if replacement is p:
            pp = p.parent
            p.parent = None
            if pp is not None:
                if p is pp.left:
                    pp.left = None
                elif p is pp.right:
                    pp.right = None
        self.move_root_to_front(tab, r)

    def keys(self):
        out = []
        for head in self.table or []:
            e = head
            while e is not None:
                out.append(e.key)
                e = e.next
        return out

    def has_tree_bin(self):
        return any(head is not None and head.tree for head in self.table or [])
