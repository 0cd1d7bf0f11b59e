"""Synthetic MultiBreak-SV inputs of the shapes BASELINE.json names (the reference ships no generator).

Two outputs:
  * generate(...)             -> PackedProblem: the flat CSR a host hands to mbsv_problem_create, built with
                                 vectorised numpy so that the 5M-fragment configuration is practical;
  * write_reference_files(...) -> the same problem as the three text files of the reference's formats
                                 (doc/MultiBreakSV-Manual.txt:203-227; MultiBreakSV.java:319-473), with IDs that
                                 satisfy generateIndependentSubproblems.pl's `longread_(\\d+)` (:58) and the
                                 ESP naming of M5toMBSV.py:274.

Shapes: experiments `illumina 0.01 10`, `pacbio 0.15 3`; perr 0.001; paired-end alignments have
len 200 and one ESP; long-read alignments have len ~ LogN(ln 5000, 0.5) clipped to [1000, 30000] and 1/2/3 ESPs
with probability .75/.20/.05; numerrors ~ Binomial(len, pseq) for the true mapping and Binomial(len, 1.5 pseq)
for decoys; one "true SV" cluster per ~lambda_D fragments plus small decoy clusters.
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass

import numpy as np

from ._abi import PackedProblem

EXPERIMENTS = {"illumina": (0.01, 10.0), "pacbio": (0.15, 3.0)}


@dataclass
class SynthConfig:
    name: str
    n_frag: int
    pe_frac: float            # fraction of paired-end (illumina) fragments
    sub_sizes: str            # "geom5" | "zipf2" | "single"
    sub_max: int = 50_000
    seed: int = 20260101
    n_chains: int = 1         # the chain count BASELINE.json quotes for the config


CONFIGS = {
    # BASELINE.json configs[1..4]
    "cfg2": SynthConfig("cfg2-illumina-pe-100k", 100_000, 1.0, "geom5", seed=20260102, n_chains=1),
    "cfg3": SynthConfig("cfg3-pacbio-200k", 200_000, 0.0, "geom5", seed=20260103, n_chains=16),
    "cfg4": SynthConfig("cfg4-mixed-5M", 5_000_000, 0.7, "zipf2", seed=20260104, n_chains=64),
    "cfg5": SynthConfig("cfg5-giant-500k", 500_000, 0.7, "single", seed=20260105, n_chains=64),
}


def _hi_lo(x: float) -> tuple[int, int]:
    u = struct.unpack("<Q", struct.pack("<d", x))[0]
    hi = u >> 32
    return (hi - (1 << 32) if hi & 0x80000000 else hi), u & 0xFFFFFFFF


def _from_hi_lo(hi: int, lo: int) -> float:
    return struct.unpack("<d", struct.pack("<Q", ((hi & 0xFFFFFFFF) << 32) | (lo & 0xFFFFFFFF)))[0]


def jlog(x: float) -> float:
    """fdlibm 5.3 __ieee754_log (what java.lang.StrictMath.log specifies) in plain Python doubles: the generator fills
    Experiment.logprobcov (MultiBreakSV.java:612-614,1644-1649) itself, bit for bit like mbsv_fill_logprobcov, without
    calling into the library it generates inputs for (tests/test_abi_cpu.py compares the two)."""
    ln2_hi, ln2_lo = _from_hi_lo(0x3fe62e42, 0xfee00000), _from_hi_lo(0x3dea39ef, 0x35793c76)
    two54 = _from_hi_lo(0x43500000, 0)
    Lg = [_from_hi_lo(h, l) for h, l in ((0x3FE55555, 0x55555593), (0x3FD99999, 0x9997FA04), (0x3FD24924, 0x94229359),
                                          (0x3FCC71C5, 0x1D8E78AF), (0x3FC74664, 0x96CB03DE), (0x3FC39A09, 0xD078C69F),
                                          (0x3FC2F112, 0xDF3E5244))]
    hx, lx = _hi_lo(x)
    k = 0
    if hx < 0x00100000:
        if ((hx & 0x7fffffff) | lx) == 0:
            return float("-inf")
        if hx < 0:
            return float("nan")
        k -= 54
        x *= two54
        hx, lx = _hi_lo(x)
    if hx >= 0x7ff00000:
        return x + x
    k += (hx >> 20) - 1023
    hx &= 0x000fffff
    i = (hx + 0x95f64) & 0x100000
    x = _from_hi_lo(hx | (i ^ 0x3ff00000), lx)
    k += i >> 20
    f = x - 1.0
    dk = float(k)
    if (0x000fffff & (2 + hx)) < 3:
        if f == 0.0:
            return 0.0 if k == 0 else dk * ln2_hi + dk * ln2_lo
        R = f * f * (0.5 - 0.33333333333333333 * f)
        return f - R if k == 0 else dk * ln2_hi - ((R - dk * ln2_lo) - f)
    s_ = f / (2.0 + f)
    z = s_ * s_
    i = hx - 0x6147a
    w = z * z
    j = 0x6b851 - hx
    t1 = w * (Lg[1] + w * (Lg[3] + w * Lg[5]))
    t2 = z * (Lg[0] + w * (Lg[2] + w * (Lg[4] + w * Lg[6])))
    i |= j
    R = t2 + t1
    if i > 0:
        hfsq = 0.5 * f * f
        return f - (hfsq - s_ * (hfsq + R)) if k == 0 else dk * ln2_hi - ((hfsq - (s_ * (hfsq + R) + dk * ln2_lo)) - f)
    return f - s_ * (f - R) if k == 0 else dk * ln2_hi - ((s_ * (f - R) - dk * ln2_lo) - f)


def logprobcov_table(lam: float, max_support: int) -> np.ndarray:
    """logprobcov[k] = logPoisson(k; lambda), k = 1..max_support, [0] = 0 (MultiBreakSV.java:612-614,1644-1649)."""
    out = np.zeros(max_support + 1, np.float64)
    ll = jlog(lam)
    logs = [0.0] + [jlog(float(i)) for i in range(1, max_support + 1)]
    for k in range(1, max_support + 1):
        val = -lam + k * ll
        for i in range(1, k + 1):
            val -= logs[i]
        out[k] = val
    return out


def _sub_sizes(rng: np.random.Generator, cfg: SynthConfig, n_frag: int) -> np.ndarray:
    if cfg.sub_sizes == "single":
        return np.array([n_frag], np.int64)
    sizes = []
    total = 0
    while total < n_frag:
        n = max(1024, int((n_frag - total) / 4))
        if cfg.sub_sizes == "geom5":
            s = rng.geometric(1.0 / 5.0, n)            # 1 + Geom, mean 5
        else:                                          # Zipf(alpha = 2) truncated to [1, sub_max]
            s = rng.zipf(2.0, n)
            s = s[s <= cfg.sub_max]
        sizes.append(s.astype(np.int64))
        total += int(s.sum())
    s = np.concatenate(sizes)
    cut = int(np.searchsorted(np.cumsum(s), n_frag)) + 1
    s = s[:cut]
    s[-1] -= int(s.sum()) - n_frag
    return s[s > 0]


def _ragged_arange(counts: np.ndarray) -> np.ndarray:
    """[0..c0), [0..c1), ... concatenated."""
    ptr = np.concatenate([[0], np.cumsum(counts)])
    return np.arange(ptr[-1]) - np.repeat(ptr[:-1], counts)


@dataclass
class Skeleton:
    """The part of a synthetic batch that fixes the work per subproblem: subproblem sizes, and for every
    fragment its platform and number of candidate alignments.  Cheap to draw for N x 5M fragments, so every
    rank can draw the same global skeleton, shard it (sharding.lpt_shard) and build only its own share."""
    sizes: np.ndarray       # [S] fragments per subproblem
    frag_is_pe: np.ndarray  # [F]
    nal: np.ndarray         # [F] alignments per fragment

    @property
    def weights(self) -> np.ndarray:
        """fragment x alignment count of each subproblem (SURVEY §8e LPT weight, per chain)."""
        ptr = np.concatenate([[0], np.cumsum(self.sizes)])
        csum = np.concatenate([[0], np.cumsum(self.nal)])
        return (csum[ptr[1:]] - csum[ptr[:-1]]).astype(np.int64)

    def select(self, subs: np.ndarray) -> "Skeleton":
        ptr = np.concatenate([[0], np.cumsum(self.sizes)])
        sz = self.sizes[subs]
        idx = np.repeat(ptr[subs], sz) + _ragged_arange(sz)
        return Skeleton(sz, self.frag_is_pe[idx], self.nal[idx])


def draw_skeleton(cfg: SynthConfig | str, n_frag: int | None = None, seed: int | None = None) -> Skeleton:
    if isinstance(cfg, str):
        cfg = CONFIGS[cfg]
    F = int(cfg.n_frag if n_frag is None else n_frag)
    rng = np.random.default_rng(cfg.seed if seed is None else seed)
    sizes = _sub_sizes(rng, cfg, F)
    frag_is_pe = rng.random(F) < cfg.pe_frac
    # alignments per fragment: PE 1+Poisson(1); long reads 1+min(15, Geom(0.25)) (<= 16)
    nal = np.where(frag_is_pe, 1 + rng.poisson(1.0, F), 1 + np.minimum(15, rng.geometric(0.25, F))).astype(np.int64)
    return Skeleton(sizes, frag_is_pe, nal)


def generate(cfg: SynthConfig | str, n_frag: int | None = None, seed: int | None = None,
             shuffle_orders: bool = False, skeleton: Skeleton | None = None, conncomp_frac: float = 0.0) -> PackedProblem:
    """Build the CSR of a synthetic batch of independent subproblems.

    shuffle_orders: permute supports_order inside each subproblem (exercises the summation-order dependence
    of REPLAY_EXACT); the benchmark keeps the identity.
    skeleton: build exactly these subproblems (a shard of a larger skeleton) instead of drawing new ones.
    conncomp_frac: fraction of the clusters with >= 3 ESPs that become GASV connected components (localization -1)
    with 2-4 overlapping maximal sub-clusters (tests only: built with a Python loop).
    """
    if isinstance(cfg, str):
        cfg = CONFIGS[cfg]
    if skeleton is None:
        skeleton = draw_skeleton(cfg, n_frag, seed)
    rng = np.random.default_rng((cfg.seed if seed is None else seed) + 7919)
    sizes, frag_is_pe, nal = skeleton.sizes, skeleton.frag_is_pe, skeleton.nal
    F = int(sizes.sum())
    S = len(sizes)
    sub_frag_ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    frag_sub = np.repeat(np.arange(S), sizes)

    # experiments in "experiments.keySet() order": index 0 = illumina, 1 = pacbio when both are present
    if cfg.pe_frac >= 1.0:
        labels = ["illumina"]
    elif cfg.pe_frac <= 0.0:
        labels = ["pacbio"]
    else:
        labels = ["illumina", "pacbio"]
    pseq = np.array([EXPERIMENTS[l][0] for l in labels])
    lam = np.array([EXPERIMENTS[l][1] for l in labels])
    E = len(labels)
    frag_exp = np.where(frag_is_pe, labels.index("illumina") if "illumina" in labels else 0,
                        labels.index("pacbio") if "pacbio" in labels else 0).astype(np.int32)
    A = int(nal.sum())
    frag_aln_ptr = np.concatenate([[0], np.cumsum(nal)])
    aln_frag = np.repeat(np.arange(F), nal)
    aln_idx = _ragged_arange(nal)
    true_idx = (rng.random(F) * nal).astype(np.int64)
    aln_true = aln_idx == true_idx[aln_frag]
    aln_pe = frag_is_pe[aln_frag]
    aln_exp = frag_exp[aln_frag]
    ln = np.clip(np.exp(rng.normal(np.log(5000.0), 0.5, A)), 1000, 30000).astype(np.int64)
    aln_len = np.where(aln_pe, 200, ln).astype(np.int32)
    p_err = pseq[aln_exp] * np.where(aln_true, 1.0, 1.5)
    aln_numerrors = rng.binomial(aln_len, np.minimum(p_err, 0.99)).astype(np.int32)
    nesp = np.where(aln_pe, 1, rng.choice([1, 2, 3], A, p=[0.75, 0.20, 0.05])).astype(np.int64)
    K = int(nesp.sum())
    aln_esp_ptr = np.concatenate([[0], np.cumsum(nesp)])
    esp_aln = np.repeat(np.arange(A), nesp)
    esp_sub = frag_sub[aln_frag[esp_aln]]

    # clusters: per subproblem t_s "true SV" clusters (one per ~5 fragments) + d_s decoy clusters
    t_s = np.maximum(1, np.rint(sizes / 5.0)).astype(np.int64)
    d_s = np.maximum(1, sizes).astype(np.int64)
    raw_ptr = np.concatenate([[0], np.cumsum(t_s + d_s)])
    to_true = aln_true[esp_aln] | (rng.random(K) < 0.5)
    pick = rng.random(K)
    raw = np.where(to_true, raw_ptr[esp_sub] + (pick * t_s[esp_sub]).astype(np.int64),
                   raw_ptr[esp_sub] + t_s[esp_sub] + (pick * d_s[esp_sub]).astype(np.int64))
    # an alignment must not list the same ESP twice; distinct leaves per entry make that automatic.
    used = np.zeros(raw_ptr[-1], bool)
    used[raw] = True
    remap = np.cumsum(used) - 1
    esp_cluster = remap[raw].astype(np.int32)
    C = int(used.sum())
    raw_sub = np.repeat(np.arange(S), t_s + d_s)
    cl_sub = raw_sub[used]
    sub_cluster_ptr = np.concatenate([[0], np.cumsum(np.bincount(cl_sub, minlength=S))]).astype(np.int64)

    # every alignment-ESP entry is its own leaf, owned by its fragment
    order = np.argsort(esp_cluster, kind="stable")
    nleaf = np.bincount(esp_cluster, minlength=C).astype(np.int64)
    cluster_leaf_ptr = np.concatenate([[0], np.cumsum(nleaf)])
    leaf_slot_sorted = _ragged_arange(nleaf)
    aln_esp_leaf = np.empty(K, np.int32)
    aln_esp_leaf[order] = leaf_slot_sorted.astype(np.int32)
    leaf_owner = aln_frag[esp_aln[order]].astype(np.int32)
    max_support = int(nleaf.max())
    cluster_hist_ptr = np.concatenate([[0], np.cumsum(nleaf + 1)])

    # AlterSV lists (MultiBreakSV.java:748-771, 1303-1330)
    single_leaf = nal[leaf_owner] == 1
    leaf_cluster = np.repeat(np.arange(C), nleaf)
    uniq_counts = np.bincount(leaf_cluster[single_leaf], minlength=C)
    eligible = uniq_counts > 1
    sel = single_leaf & eligible[leaf_cluster]
    sc, sf = leaf_cluster[sel], leaf_owner[sel]
    keep = np.ones(len(sc), bool)
    keep[1:] = (sc[1:] != sc[:-1]) | (sf[1:] != sf[:-1])
    sv_cluster = np.nonzero(eligible)[0].astype(np.int32)
    sv_frag = sf[keep].astype(np.int32)
    sv_frag_ptr = np.concatenate([[0], np.cumsum(np.bincount(sc[keep], minlength=C)[eligible])]).astype(np.int32)
    sv_numchanged = uniq_counts[eligible].astype(np.int32)
    sub_sv_ptr = np.concatenate([[0], np.cumsum(np.bincount(cl_sub[eligible], minlength=S))]).astype(np.int32)

    supports_order = np.arange(C, dtype=np.int32)
    if shuffle_orders:
        key = rng.random(C) + cl_sub          # permute inside each subproblem
        supports_order = np.argsort(key, kind="stable").astype(np.int32)

    cluster_kind = np.zeros(C, np.uint8)
    cluster_root_ptr = np.arange(C + 1, dtype=np.int32)
    root_leaf_ptr = cluster_leaf_ptr.astype(np.int32)
    root_leaf = leaf_slot_sorted.astype(np.int32)
    if conncomp_frac > 0:
        cand = np.nonzero((nleaf >= 3) & (rng.random(C) < conncomp_frac))[0]
        cluster_kind[cand] = 1
        is_cc = cluster_kind == 1
        crp, rlp, rl = [0], [0], []
        for c in range(C):
            nl = int(nleaf[c])
            if not is_cc[c]:
                rl.extend(range(nl)); rlp.append(len(rl)); crp.append(len(rlp) - 1)
                continue
            nroot = int(rng.integers(2, 5))
            member = rng.random((nroot, nl)) < 0.55
            member[rng.integers(0, nroot, nl), np.arange(nl)] = True       # every leaf is in some sub-cluster
            for r in range(nroot):
                idx = np.nonzero(member[r])[0]
                if len(idx) == 0:
                    idx = np.array([int(rng.integers(0, nl))])
                rl.extend(idx.tolist()); rlp.append(len(rl))
            crp.append(len(rlp) - 1)
        cluster_root_ptr = np.array(crp, np.int32)
        root_leaf_ptr = np.array(rlp, np.int32)
        root_leaf = np.array(rl, np.int32)

    T = np.zeros((E, max_support + 1), np.float64)
    for x in range(E):
        T[x] = logprobcov_table(float(lam[x]), max_support)

    arrays = dict(
        sub_frag_ptr=sub_frag_ptr.astype(np.int32), sub_cluster_ptr=sub_cluster_ptr.astype(np.int32),
        sub_sv_ptr=sub_sv_ptr, frag_aln_ptr=frag_aln_ptr.astype(np.int32), aln_numerrors=aln_numerrors,
        aln_len=aln_len, aln_exp=aln_exp.astype(np.int32), aln_esp_ptr=aln_esp_ptr.astype(np.int32),
        aln_esp_cluster=esp_cluster, aln_esp_leaf=aln_esp_leaf, cluster_kind=cluster_kind,
        supports_order=supports_order, cluster_hist_ptr=cluster_hist_ptr.astype(np.int32),
        cluster_leaf_ptr=cluster_leaf_ptr.astype(np.int32), leaf_owner=leaf_owner,
        cluster_root_ptr=cluster_root_ptr, root_leaf_ptr=root_leaf_ptr,
        root_leaf=root_leaf, sv_cluster=sv_cluster, sv_frag_ptr=sv_frag_ptr, sv_frag=sv_frag,
        sv_numchanged=sv_numchanged, exp_pseq=pseq, exp_lambda=lam, logprobcov=T.reshape(-1))
    scalars = dict(abi_version=1, n_sub=S, n_frag=F, n_aln=A, n_aln_esp=K, n_cluster=C, n_exp=E,
                   max_support=max_support, n_sv=len(sv_cluster), n_sv_frag=len(sv_frag), n_leaf=K,
                   n_root=len(root_leaf_ptr) - 1, n_root_leaf=len(root_leaf), n_hist=int(cluster_hist_ptr[-1]))
    p = PackedProblem(scalars, arrays)
    p.exp_labels = labels
    p.config = cfg
    return p


def subproblem_weights(p: PackedProblem, n_chains: int) -> np.ndarray:
    """LPT weight of SURVEY §8e: chains x (fragment x alignment count) of the subproblem."""
    a = p.arrays
    fptr = a["sub_frag_ptr"].astype(np.int64)
    nal_ptr = a["frag_aln_ptr"].astype(np.int64)
    return n_chains * (nal_ptr[fptr[1:]] - nal_ptr[fptr[:-1]])


def write_reference_files(p: PackedProblem, outdir: str, frag_numbers: np.ndarray | None = None) -> dict:
    """Write assignments / experiments / clusters text files for a (small) generated problem."""
    os.makedirs(outdir, exist_ok=True)
    a = p.arrays
    F, C = p.scalars["n_frag"], p.scalars["n_cluster"]
    nums = np.arange(1, F + 1) * 7 + 100000 if frag_numbers is None else frag_numbers
    labels = p.exp_labels
    esp_name = {}
    lines = ["FragmentID\tDiscordantPairs\tNumErrors\tBasesInAlignment(TargetEnd-TargetStart)\tExperimentLabel"]
    for f in range(F):
        for ai, al in enumerate(range(a["frag_aln_ptr"][f], a["frag_aln_ptr"][f + 1])):
            esps = []
            for seg, j in enumerate(range(a["aln_esp_ptr"][al], a["aln_esp_ptr"][al + 1])):
                nm = f"longread_{nums[f]}_{seg}.{ai}-{seg + 1}.{ai}"
                esp_name[j] = nm
                esps.append(nm)
            lines.append(f"longread_{nums[f]}\t{','.join(esps)}\t{a['aln_numerrors'][al]}\t{a['aln_len'][al]}\t"
                         f"{labels[a['aln_exp'][al]]}")
    paths = {k: os.path.join(outdir, v) for k, v in
             dict(assignments="assignments.txt", experiments="experiments.txt", clusters="gasv.clusters").items()}
    with open(paths["assignments"], "w") as fh:
        fh.write("\n".join(lines) + "\n")
    with open(paths["experiments"], "w") as fh:
        fh.write("ExperimentLabel\tPseq\tLambdaD\n")
        for x, l in enumerate(labels):
            fh.write(f"{l}\t{float(a['exp_pseq'][x])!r}\t{float(a['exp_lambda'][x])!r}\n")
    members = [[] for _ in range(C)]
    for j, c in enumerate(a["aln_esp_cluster"]):
        members[c].append(esp_name[j])
    # leaf slot -> ESP name (leaves are the entries sorted by cluster, in entry order)
    order = np.argsort(a["aln_esp_cluster"], kind="stable")
    with open(paths["clusters"], "w") as fh:
        for c in range(C):
            names = [esp_name[int(j)] for j in order[a["cluster_leaf_ptr"][c]:a["cluster_leaf_ptr"][c + 1]]]
            if a["cluster_kind"][c] == 0:
                fh.write(f"c{c + 1}\t{len(names)}\t100.0\tD\t{', '.join(names)}\t1\t1\t1, 2, 3, 4\n")
            else:
                fh.write(f"c{c + 1}\t{len(names)}\t-1\tD\t{', '.join(names)}\t1\t1\t1, 2, 3, 4\n")
                for k, r in enumerate(range(a["cluster_root_ptr"][c], a["cluster_root_ptr"][c + 1])):
                    sub = [names[int(l)] for l in a["root_leaf"][a["root_leaf_ptr"][r]:a["root_leaf_ptr"][r + 1]]]
                    fh.write(f"c{c + 1}.{k + 1}\t{len(sub)}\t90.0\tD\t{', '.join(sub)}\t1\t1\t1, 2, 3, 4\n")
    return paths
