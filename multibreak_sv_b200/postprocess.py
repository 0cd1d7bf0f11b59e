"""Cluster scores from the sampler's output: mirror of the reference's src/quick-post-process.py.

Two entry points with the same arithmetic:

* ``quick_post_process(clustersfile, finalbreakpointfile, finalassignmentfile, minsupport, outfile)`` -- text files in,
  text file out; options, columns, number formats and error messages of quick-post-process.py:33-131.
* ``scores_from_tallies(...)`` -- the same two scores straight from the integer tallies of ``mbsv_run`` (no text
  round trip); used when the caller never writes the ``.final*`` files.

Scores (quick-post-process.py:5-10):
  SVFreq     = (number of long-window samples in which the cluster had >= minsupport supporting alignments) / NumIters
               -- the sum of the Support columns from `minsupport` on (:84-87);
  AvgAssign  = mean over the cluster's discordant pairs of the pair's summed AvgAssignment (:96-107, :122-127).

Row order: the reference iterates a Python 2 dict (:112), i.e. an order that depends on the interpreter's string hash;
this mirror writes clusters in the order of the clusters file (what the same script does under Python 3).  Compare as
sets against a Python 2 run.
"""
from __future__ import annotations

import sys

import numpy as np

HEADER = "#ClusterID\tSVFreq\tAvgAssign\tSVType\tStartChr\tStartLocRange\tEndChr\tEndLocRange\tNumPairs\tPairs\n"


def read_clusters(path):
    """quick-post-process.py:52-68: every row of the clusters file (dotted maximal-cluster rows included, no header
    handling -- a header row fails on int() exactly as it does there)."""
    info = {}
    with open(path) as fin:
        for line in fin:
            row = line.strip().split("\t")
            locs = row[7].split(", ")
            starts = [int(v) for v in locs[0::2]]
            ends = [int(v) for v in locs[1::2]]
            info[row[0]] = [row[3], int(row[5]), [min(starts), max(starts)], int(row[6]), [min(ends), max(ends)],
                            row[4].split(", ")]
    return info


def read_frequencies(path, minsupport):
    """:73-88 -- columns are ClusterID, NumIters, Support0, Support1, ..."""
    freq = {}
    with open(path) as fin:
        for line in fin:
            if "ClusterID" in line:
                continue
            row = line.strip().split("\t")
            numiters = int(row[1])
            total = 0.0
            for i in range(minsupport + 2, len(row)):
                total += float(row[i])
            freq[row[0]] = total / numiters
    return freq


def read_pair_assignments(path):
    """:96-107 -- a pair is selected at most once per mapping, so its AvgAssignments over alignments add up."""
    pairs = {}
    with open(path) as fin:
        for line in fin:
            if "FragmentID" in line:
                continue
            row = line.strip().split("\t")
            if row[3] == "error":
                continue
            for pair in row[3].split(","):
                pairs[pair] = pairs.get(pair, 0.0) + float(row[2])
    return pairs


def format_row(cid, svfreq, avgval, info):
    return "%s\t%.4f\t%.4f\t%s\t%d\t%d-%d\t%d\t%d-%d\t%d\t%s\n" % (
        cid, svfreq, avgval, info[0], info[1], info[2][0], info[2][1], info[3], info[4][0], info[4][1], len(info[5]),
        ",".join(info[5]))


def quick_post_process(clustersfile, finalbreakpointfile, finalassignmentfile, minsupport=2, outfile="out.txt"):
    info = read_clusters(clustersfile)
    freq = read_frequencies(finalbreakpointfile, minsupport)
    pairs = read_pair_assignments(finalassignmentfile)
    with open(outfile, "w") as out:
        out.write(HEADER)
        for origcid, this in info.items():
            # a dotted id is one of the maximal clusters of a connected component; the sampler reports the component (:113-118)
            cid = origcid if origcid in freq else origcid.split(".")[0]
            if cid not in freq:
                sys.exit('ERROR: cid="%s" (parsed to "%s") does not appear in finalbreakpoints file.' % (origcid, cid))
            val = 0
            for pair in this[5]:
                if pair not in pairs:
                    sys.exit('ERROR: end-sequence pair "%s" does not appear in finalassignments file.' % (pair))
                val += pairs[pair]
            out.write(format_row(origcid, freq[cid], val / len(this[5]), this))
    return outfile


def scores_from_tallies(cluster_ids, cluster_hist_ptr, cluster_hist, frag_order, frag_aln_ptr, aln_esp_ptr, esp_names,
                        aln_count, burnin, clusters, minsupport=2):
    """SVFreq / AvgAssign of every row of `clusters` (the dict of read_clusters) from the run's integer tallies.

    cluster_hist[cluster_hist_ptr[c] + k] = samples of the long window with support k (k >= 1; bin 0 is never
    tallied, the writer derives it: MultiBreakSV.java:1754-1770), aln_count[a] = samples with alignment a selected,
    burnin = window length (the NumIters column).  `frag_order` lists fragments in the writer's sorted order
    (getFragmentIDs :1703-1709) so that the floating-point sums run in the order the text path adds them.
    Returns {cluster row id: (svfreq, avgassign)}.
    """
    cluster_hist = np.asarray(cluster_hist, np.float64)
    freq = {}
    for c, cid in enumerate(cluster_ids):
        lo, hi = int(cluster_hist_ptr[c]), int(cluster_hist_ptr[c + 1])
        bins = cluster_hist[lo:hi].copy()
        if hi > lo:
            bins[0] = burnin - bins[1:].sum()
        total = 0.0
        for v in bins[minsupport:]:
            total += float(v)
        freq[cid] = total / burnin
    pairs = {}
    for f in frag_order:
        for a in range(int(frag_aln_ptr[f]), int(frag_aln_ptr[f + 1])):
            avg = float(aln_count[a]) / burnin
            for j in range(int(aln_esp_ptr[a]), int(aln_esp_ptr[a + 1])):
                pairs[esp_names[j]] = pairs.get(esp_names[j], 0.0) + avg
    out = {}
    for origcid, this in clusters.items():
        cid = origcid if origcid in freq else origcid.split(".")[0]
        if cid not in freq:
            raise KeyError('cid="%s" (parsed to "%s") does not appear in the tallies' % (origcid, cid))
        val = 0
        for pair in this[5]:
            if pair not in pairs:
                raise KeyError('end-sequence pair "%s" does not appear in the tallies' % pair)
            val += pairs[pair]
        out[origcid] = (freq[cid], val / len(this[5]))
    return out


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(prog="quick-post-process",
                                 description="Outputs a file with Cluster information, SVFrequency, and AverageProbability.")
    ap.add_argument("--clustersfile")
    ap.add_argument("--finalbreakpointfile")
    ap.add_argument("--finalassignmentfile")
    ap.add_argument("--minsupport", type=int, default=2)
    ap.add_argument("--outfile", default="out.txt")
    o = ap.parse_args(argv)
    if o.clustersfile is None or o.finalbreakpointfile is None or o.finalassignmentfile is None:
        sys.exit("Error: script requires clusters file, finalbreakpoints file, and finalassignments file.\n")
    quick_post_process(o.clustersfile, o.finalbreakpointfile, o.finalassignmentfile, o.minsupport, o.outfile)
    print("Wrote to %s" % o.outfile)


if __name__ == "__main__":
    main()
