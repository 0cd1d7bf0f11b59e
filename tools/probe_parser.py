"""Tuning aid (CPU only): time of the native parser / packer and the partition on generated reference-format files."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multibreak_sv_b200 import synth, host
t0 = time.time()
g = synth.generate("cfg4", n_frag=500_000, seed=1)
print("generate %.1f s" % (time.time() - t0), g.scalars["n_sub"], g.scalars["n_aln"], g.scalars["n_cluster"])
t0 = time.time()
paths = synth.write_reference_files(g, "/tmp/parse_bench")
print("write files %.1f s" % (time.time() - t0), {k: os.path.getsize(v) >> 20 for k, v in paths.items()}, "MB")
for cf in (False, True):
    t0 = time.time()
    sv = host.MultiBreakSV(["--clusterfile", paths["clusters"], "--assignmentfile", paths["assignments"], "--experimentfile",
                            paths["experiments"], "--numiterations", "1000", "--perr", "0.001"] + (["--readclustersfirst"] if cf else []))
    sv.load()
    t1 = time.time()
    p = sv.packed()
    print("clusters_first", cf, "load %.2f s" % (t1 - t0), p.scalars["n_frag"], p.scalars["n_aln"], p.scalars["n_cluster"])
    sv.close()
t0 = time.time()
n, rows = host.generateIndependentSubproblems(paths["clusters"], None)
print("partition %.2f s -> %d subsets" % (time.time() - t0, n))
t0 = time.time()
info = {}
p, skipped = host.load_partitioned(paths["clusters"], paths["assignments"], paths["experiments"], info=info)
print(info)
print("load_partitioned %.2f s -> %d subproblems (%d skipped), %d fragments" % (time.time() - t0, p.scalars["n_sub"], skipped, p.scalars["n_frag"]))
