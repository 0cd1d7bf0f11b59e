"""java.util.HashMap tree bins: a bin that collects nine keys at capacity >= 64 becomes a red-black tree and its iteration
order stops being "insertion order" (the root is moved to the front, new nodes are linked behind their tree parent).  The
reference's trajectory depends on that order (fragments.keySet(), MultiBreakSV.java:999,1401), so the host mirror replays
such maps through a literal emulation (csrc/java_hashmap.hpp).  Checked here against an independent transliteration of the
JDK source (tests/java_hashmap_ref.py) on forged collisions: equal hashCodes ("Aa"/"BB" blocks), equal buckets with
different hashes, resizes that split and untreeify bins, removals that rebalance them."""
import ctypes as C
import itertools
import random

import numpy as np
import pytest

from java_hashmap_ref import JavaHashMap, spread, string_hash


@pytest.fixture(scope="module")
def replay(mbsv):
    L = mbsv.load_library()
    L.mbsv_java_hashmap_replay.argtypes = [C.c_char_p, C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.c_int32,
                                           C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.mbsv_java_hashmap_replay.restype = C.c_int

    def run(keys, ops, literal):
        blob = b"".join(k.encode() + b"\0" for k in keys)
        o = np.ascontiguousarray(ops, np.int32)
        out = np.zeros(len(keys), np.int32)
        info = np.zeros(2, np.int32)
        n = L.mbsv_java_hashmap_replay(blob, len(keys), o.ctypes.data_as(C.POINTER(C.c_int32)), len(o), int(literal),
                                       out.ctypes.data_as(C.POINTER(C.c_int32)), info.ctypes.data_as(C.POINTER(C.c_int32)))
        assert n >= 0
        return [keys[i] for i in out[:n]], bool(info[0]), int(info[1])
    return run


def reference(keys, ops):
    m = JavaHashMap()
    for op in ops:
        if op > 0:
            m.put(keys[op - 1])
        else:
            m.remove(keys[-op - 1])
    return m.keys(), m.has_tree_bin(), len(m.table)


def same_hashcode(n, prefix="c"):
    """n distinct strings with one String.hashCode: "Aa" and "BB" hash alike, so do equal-length sequences of them."""
    out = []
    for blocks in itertools.product(("Aa", "BB"), repeat=7):
        out.append(prefix + "".join(blocks))
        if len(out) == n:
            break
    assert len({string_hash(k) for k in out}) == 1
    return out


def same_bucket(n, bucket, mask, prefix="longread_"):
    """n fragment-style names whose spread hash lands in `bucket` under `mask` (different hashes)."""
    out, i = [], 0
    while len(out) < n:
        k = f"{prefix}{i}"
        if (spread(string_hash(k)) & mask) == bucket:
            out.append(k)
        i += 1
    return out


def test_treeified_bin_changes_the_iteration_order(replay):
    keys = same_hashcode(12) + [f"filler_{i}" for i in range(40)]     # 52 keys: capacity 64 -> 128 on the way
    ops = list(range(1, len(keys) + 1))
    # fillers first (capacity reaches 64), then the colliding keys: the 9th one treeifies the bin
    ops = ops[12:] + ops[:12]
    want, tree, cap = reference(keys, ops)
    assert tree and cap == 128
    for literal in (True, False):
        got, treeified, gcap = replay(keys, ops, literal)
        assert got == want and treeified and gcap == cap
    coll = [k for k in want if k in set(keys[:12])]
    assert coll != keys[:12]                      # not insertion order any more: a plain-list model would be wrong here


@pytest.mark.parametrize("seed", range(12))
def test_random_puts_and_removes_match_the_transliteration(replay, seed):
    rng = random.Random(seed)
    a = same_hashcode(rng.randint(9, 40), prefix=rng.choice(["c", "longread_", "x9"]))
    b = same_bucket(rng.randint(9, 30), rng.randrange(64), 63)
    c = same_bucket(rng.randint(0, 20), rng.randrange(256), 255, prefix="frag")
    fill = [f"k{rng.randrange(10**9)}" for _ in range(rng.randint(20, 400))]
    keys = list(dict.fromkeys(a + b + c + fill))
    order = list(range(len(keys)))
    rng.shuffle(order)
    if seed % 3 == 0:                             # colliding keys late: the table is large when the bin fills up
        order.sort(key=lambda i: keys[i] in set(a + b))
    ops, live = [], []
    for i in order:
        ops.append(i + 1)
        live.append(i)
        if rng.random() < 0.25 and live:
            j = live.pop(rng.randrange(len(live)))
            ops.append(-(j + 1))
    # remove most of one colliding family at the end: rebalancing, then untreeify
    for i in [i for i in live if keys[i] in set(a)][:-3]:
        ops.append(-(i + 1))
    for upto in sorted({len(ops), len(ops) * 2 // 3, len(ops) // 3}):
        want, tree, cap = reference(keys, ops[:upto])
        for literal in (True, False):
            got, _, gcap = replay(keys, ops[:upto], literal)
            assert got == want, (seed, upto, literal)
            assert gcap == cap


def test_plain_maps_still_take_the_closed_form(replay):
    rng = random.Random(5)
    keys = [f"longread_{rng.randrange(10**7)}" for _ in range(3000)]
    keys = list(dict.fromkeys(keys))
    ops = list(range(1, len(keys) + 1)) + [-(i + 1) for i in range(0, len(keys), 7)]
    want, tree, cap = reference(keys, ops)
    closed, treeified, gcap = replay(keys, ops, False)
    literal, _, _ = replay(keys, ops, True)
    assert not tree and not treeified and closed == want == literal and gcap == cap


def test_loader_accepts_files_whose_fragment_map_treeifies(mbsv, tmp_path):
    """Fragment IDs forged into one bucket: the load used to be refused (MBSV_ERR_UNSUPPORTED); now the fragment order is the
    emulated red-black order."""
    from multibreak_sv_b200 import host
    coll = same_bucket(11, 17, 63)
    other = [f"longread_{900000 + i}" for i in range(30)]
    names = other + coll                          # capacity 64 is reached before the 9th colliding ID arrives
    with open(tmp_path / "experiments.txt", "w") as fh:
        fh.write("ExperimentLabel\tPseq\tLambdaD\npacbio\t0.15\t3\n")
    with open(tmp_path / "assignments.txt", "w") as fh:
        fh.write("FragmentID\tDiscordantPairs\tNumErrors\tBasesInAlignment\tExperimentLabel\n")
        for n in names:
            fh.write(f"{n}\t{n}_0.0-1.0\t30\t5000\tpacbio\n")
    with open(tmp_path / "gasv.clusters", "w") as fh:
        for i, n in enumerate(names):
            fh.write(f"c{i + 1}\t1\t100.0\tD\t{n}_0.0-1.0\t1\t1\t1, 2, 3, 4\n")
    sv = host.MultiBreakSV(["--clusterfile", str(tmp_path / "gasv.clusters"), "--assignmentfile",
                            str(tmp_path / "assignments.txt"), "--experimentfile", str(tmp_path / "experiments.txt"),
                            "--numiterations", "100", "--perr", "0.001"]).load()
    want, tree, _ = reference(names, list(range(1, len(names) + 1)))
    assert tree
    assert sv.names(0, len(names)) == want
