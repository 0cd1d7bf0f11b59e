"""Whole workflow on generated reference-format files: write files -> partition + load in memory -> GPU sampling.
Prints the wall time of each stage (the file writer is Python and only there to have input)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import multibreak_sv_b200 as m
from multibreak_sv_b200 import host, synth

n_frag = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
chains = int(sys.argv[3]) if len(sys.argv) > 3 else 64
out = "/tmp/mbsv_demo"
os.makedirs(out, exist_ok=True)
g = synth.generate("cfg4", n_frag=n_frag, seed=1)
t0 = time.time()
paths = synth.write_reference_files(g, out)
print("wrote files %.1f s: %s MB" % (time.time() - t0, {k: os.path.getsize(v) >> 20 for k, v in paths.items()}), flush=True)
t0 = time.time()
b = host.PartitionedBatch(paths["clusters"], paths["assignments"], paths["experiments"])
p, info = b.packed, {"skipped": b.skipped, "treeified": b.treeified}
t_load = time.time() - t0
print("partition + load: %.2f s -> %d subproblems, %d fragments, %d alignments %s" % (t_load, p.scalars["n_sub"], p.scalars["n_frag"], p.scalars["n_aln"], info), flush=True)
t0 = time.time()
dp = m.DeviceProblem(p)
t_create = time.time() - t0
t0 = time.time()
r = dp.run(m.MODE_FAST, iters, n_chains=chains, want_state=False, per_chain=False)
t_run = time.time() - t0
print("create %.2f s, run %.2f s (kernel %.0f ms): %d chains x %d iterations, %.3e reassignments -> %.3e /s end to end from files"
      % (t_create, t_run, r.kernel_ms, chains, iters, float(r.totals[4]), float(r.totals[4]) / (t_load + t_create + t_run)))
t0 = time.time()
burnin = chains * m.reference_burnin(iters)
b.writeFinalFiles(os.path.join(out, "batch"), burnin, r.aln_count, r.frag_err_count, r.cluster_hist)
print("wrote %s.final{breakpoints,assignment}.longburnin in %.2f s" % (os.path.join(out, "batch"), time.time() - t0))
