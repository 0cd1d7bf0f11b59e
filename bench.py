#!/usr/bin/env python
"""Benchmark of the MCMC sampler hot path (BASELINE.json: "MCMC fragment-reassignments/sec").

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference path on the host cores (CPU)

Workload: BASELINE.json configs[3] — mixed paired-end + long-read, 5M fragments PER GPU (weak scaling: the
global skeleton has N x 5M fragments and is LPT-sharded over the N ranks), power-law subproblem sizes,
64 chains per subproblem.  One step = one pass of the hot path (mbsv_run: init + `--iters` MH iterations +
tally epilogue for every chain) over that batch.  A fragment-reassignment is one non-lazy proposal
(Naive = 1, AlterSV = numchanged).

Timed regions
  value : K steps of mbsv_run_device, CSR and result buffers resident in HBM; barrier + synchronize on both
          sides; max over ranks.
  e2e   : K steps of mbsv_problem_create (host validation/packing + H2D of the CSR) + mbsv_run with HOST result
          buffers (D2H of the tallies) + destroy, i.e. the reference-facing C-ABI call sequence.
  roofline.achieved : algorithmic bytes per launch ((52+6k) per proposal + (28+26k) per accepted
          move) / the library's own CUDA-event time of the sampler launches.
  cpu_baseline : the oracle's incremental restatement of the Java sampler on all host cores, on a bounded
          sample of the same batch (N=1, rank 0 only).  Java itself cannot run here (no JDK in the image).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

# the sampler overlaps one kernel per state-size class: more hardware queues than CUDA's default 8 (see mbsv_core.cu);
# must be in the environment before torch creates the CUDA context
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def slice_subproblems(p, s0: int, s1: int):
    """Sub-batch [s0, s1) of a PackedProblem (re-based indices); used for the bounded CPU sample."""
    from multibreak_sv_b200 import PackedProblem
    a = p.arrays
    f0, f1 = int(a["sub_frag_ptr"][s0]), int(a["sub_frag_ptr"][s1])
    c0, c1 = int(a["sub_cluster_ptr"][s0]), int(a["sub_cluster_ptr"][s1])
    v0, v1 = int(a["sub_sv_ptr"][s0]), int(a["sub_sv_ptr"][s1])
    a0, a1 = int(a["frag_aln_ptr"][f0]), int(a["frag_aln_ptr"][f1])
    k0, k1 = int(a["aln_esp_ptr"][a0]), int(a["aln_esp_ptr"][a1])
    l0, l1 = int(a["cluster_leaf_ptr"][c0]), int(a["cluster_leaf_ptr"][c1])
    r0, r1 = int(a["cluster_root_ptr"][c0]), int(a["cluster_root_ptr"][c1])
    rl0, rl1 = int(a["root_leaf_ptr"][r0]), int(a["root_leaf_ptr"][r1])
    q0, q1 = int(a["sv_frag_ptr"][v0]), int(a["sv_frag_ptr"][v1])
    h0 = int(a["cluster_hist_ptr"][c0])
    ec = a["aln_esp_cluster"][k0:k1]
    lo = a["leaf_owner"][l0:l1]
    arrays = dict(
        sub_frag_ptr=a["sub_frag_ptr"][s0:s1 + 1] - f0, sub_cluster_ptr=a["sub_cluster_ptr"][s0:s1 + 1] - c0,
        sub_sv_ptr=a["sub_sv_ptr"][s0:s1 + 1] - v0, frag_aln_ptr=a["frag_aln_ptr"][f0:f1 + 1] - a0,
        aln_numerrors=a["aln_numerrors"][a0:a1], aln_len=a["aln_len"][a0:a1], aln_exp=a["aln_exp"][a0:a1],
        aln_esp_ptr=a["aln_esp_ptr"][a0:a1 + 1] - k0, aln_esp_cluster=np.where(ec >= 0, ec - c0, -1).astype(np.int32),
        aln_esp_leaf=a["aln_esp_leaf"][k0:k1], cluster_kind=a["cluster_kind"][c0:c1],
        supports_order=a["supports_order"][c0:c1] - c0, cluster_hist_ptr=a["cluster_hist_ptr"][c0:c1 + 1] - h0,
        cluster_leaf_ptr=a["cluster_leaf_ptr"][c0:c1 + 1] - l0, leaf_owner=np.where(lo >= 0, lo - f0, -1).astype(np.int32),
        cluster_root_ptr=a["cluster_root_ptr"][c0:c1 + 1] - r0, root_leaf_ptr=a["root_leaf_ptr"][r0:r1 + 1] - rl0,
        root_leaf=a["root_leaf"][rl0:rl1], sv_cluster=a["sv_cluster"][v0:v1] - c0,
        sv_frag_ptr=a["sv_frag_ptr"][v0:v1 + 1] - q0, sv_frag=a["sv_frag"][q0:q1] - f0,
        sv_numchanged=a["sv_numchanged"][v0:v1], exp_pseq=a["exp_pseq"], exp_lambda=a["exp_lambda"],
        logprobcov=a["logprobcov"])
    sc = dict(p.scalars)
    sc.update(n_sub=s1 - s0, n_frag=f1 - f0, n_aln=a1 - a0, n_aln_esp=k1 - k0, n_cluster=c1 - c0, n_sv=v1 - v0,
              n_sv_frag=q1 - q0, n_leaf=l1 - l0, n_root=r1 - r0, n_root_leaf=rl1 - rl0,
              n_hist=int(a["cluster_hist_ptr"][c1]) - h0)
    return PackedProblem(sc, arrays)


def stratified_sample(p, n_chains: int, iters: int, target_iterations: float, threads: int, seed: int = 20260199):
    """A bounded sample of the batch's subproblems, stratified by size.

    Every subproblem runs the same number of iterations, but a proposal on a large subproblem costs more (cache misses,
    O(F) set-up), and a power-law batch holds its large subproblems in a tiny share of the count: a prefix or uniform sample
    almost surely misses them.  Strata = subproblem sizes by powers of two ([1], [2], [3,4], [5,8], ...); stratum k holds N_k
    subproblems, of which n_k = min(N_k, max(2 x threads, its proportional share of the budget)) are drawn at random.
    Returns [(N_k, n_k, lo, hi, sub-batch)], ascending size."""
    a = p.arrays
    sizes = np.diff(a["sub_frag_ptr"].astype(np.int64))
    S = len(sizes)
    budget = max(1, int(round(target_iterations / (n_chains * max(iters, 1)))))
    rng = np.random.default_rng(seed)
    cls = np.ceil(np.log2(np.maximum(sizes, 1))).astype(np.int64)          # 1 -> 0, 2 -> 1, 3..4 -> 2, 5..8 -> 3, ...
    out = []
    for k in np.unique(cls):
        idx = np.nonzero(cls == k)[0]
        n_k = int(min(len(idx), max(2 * threads, round(budget * len(idx) / S))))
        pick = np.sort(rng.choice(idx, n_k, replace=False)) if n_k < len(idx) else idx
        out.append((len(idx), n_k, int(sizes[idx].min()), int(sizes[idx].max()), gather_subproblems(p, pick)))
    return out


def gather_subproblems(p, subs):
    """The sub-batch holding the subproblems `subs` (ascending indices), re-based; contiguous runs are sliced and joined."""
    from multibreak_sv_b200 import PackedProblem
    subs = np.asarray(subs, np.int64)
    if len(subs) == 0:
        raise ValueError("empty sample")
    # runs of consecutive subproblems
    cut = np.nonzero(np.diff(subs) != 1)[0] + 1
    runs = [(int(r[0]), int(r[-1]) + 1) for r in np.split(subs, cut)]
    parts = [slice_subproblems(p, s0, s1) for s0, s1 in runs]
    if len(parts) == 1:
        return parts[0]
    return concat_problems(parts)


# (name, the *_ptr array that indexes it or None, what its values index: scalar holding that base's size or None)
_PTRS = ("sub_frag_ptr", "sub_cluster_ptr", "sub_sv_ptr", "frag_aln_ptr", "aln_esp_ptr", "cluster_hist_ptr",
         "cluster_leaf_ptr", "cluster_root_ptr", "root_leaf_ptr", "sv_frag_ptr")
_VALUE_BASE = {"sub_frag_ptr": "n_frag", "sub_cluster_ptr": "n_cluster", "sub_sv_ptr": "n_sv", "frag_aln_ptr": "n_aln",
               "aln_esp_ptr": "n_aln_esp", "cluster_hist_ptr": "n_hist", "cluster_leaf_ptr": "n_leaf",
               "cluster_root_ptr": "n_root", "root_leaf_ptr": "n_root_leaf", "sv_frag_ptr": "n_sv_frag",
               "aln_esp_cluster": "n_cluster", "supports_order": "n_cluster", "leaf_owner": "n_frag",
               "sv_cluster": "n_cluster", "sv_frag": "n_frag"}


def concat_problems(parts):
    """Join independent sub-batches into one PackedProblem (indices shifted by the sizes of the parts in front)."""
    from multibreak_sv_b200 import PackedProblem
    keys = list(parts[0].arrays.keys())
    base = {k: 0 for k in set(_VALUE_BASE.values())}
    cols = {k: [] for k in keys}
    for q in parts:
        for k in keys:
            v = np.asarray(q.arrays[k])
            if k in ("exp_pseq", "exp_lambda", "logprobcov"):
                continue
            if k in _VALUE_BASE:
                off = base[_VALUE_BASE[k]]
                if k in _PTRS:
                    v = v + off
                    if cols[k]:
                        v = v[1:]                      # the first offset repeats the previous part's last one
                else:
                    v = np.where(v >= 0, v + off, v)
            cols[k].append(v)
        for name in base:
            base[name] += int(q.scalars[name])
    arrays = {k: (np.concatenate(cols[k]).astype(np.asarray(parts[0].arrays[k]).dtype) if cols[k] else parts[0].arrays[k])
              for k in keys}
    sc = dict(parts[0].scalars)
    for name in ("n_sub", "n_frag", "n_aln", "n_aln_esp", "n_cluster", "n_sv", "n_sv_frag", "n_leaf", "n_root",
                 "n_root_leaf", "n_hist"):
        sc[name] = int(sum(q.scalars[name] for q in parts))
    return PackedProblem(sc, arrays)


def run_cpu(sample, n_chains, iters, mode, threads):
    from oracle import oracle as orc
    op = orc.Problem(sample.scalars, sample.arrays)
    t0 = time.perf_counter()
    r = orc.run(op, mode, iters, n_chains=n_chains, java_int_wrap=False, nthreads=threads)
    dt = time.perf_counter() - t0
    return int(r.totals[4]), dt


def cpu_estimate(strata, n_chains, iters, mode, threads):
    """Whole-batch CPU throughput from the stratified sample: every stratum is timed on all host threads and scaled by
    N_k / n_k.  Returns (reassignments/s, sampled reassignments, sampled seconds, description)."""
    r_hat = t_hat = r_s = t_s = 0.0
    desc = []
    for N_k, n_k, lo, hi, sub in strata:
        r, dt = run_cpu(sub, n_chains, iters, mode, threads)
        r_hat += r * N_k / n_k
        t_hat += dt * N_k / n_k
        r_s += r
        t_s += dt
        desc.append(f"{n_k}/{N_k} of size {lo}" + (f"-{hi}" if hi != lo else ""))
    return r_hat / t_hat, r_s, t_s, "; ".join(desc)


def algorithmic_bytes(p, proposals: float, accepted: float) -> float:
    """(52 + 6k) B per proposal + (28 + 26k) B per accepted move, k = ESPs in old + new alignment."""
    k = 2.0 * p.scalars["n_aln_esp"] / p.scalars["n_aln"]
    return proposals * (52 + 6 * k) + accepted * (28 + 26 * k)


def build_workload(args, rank, world):
    from multibreak_sv_b200 import sharding, synth
    cfg = synth.CONFIGS[args.config]
    base = args.n_frag if args.n_frag else cfg.n_frag
    if cfg.sub_sizes == "single":
        # giant component (BASELINE configs[4]): cannot be split spatially; every rank holds the same subproblem and
        # runs its share of the chains (mbsv_split_chains), tallies are summed with one ncclReduce
        return cfg, synth.generate(cfg, n_frag=base)
    # weak: `base` fragments PER GPU (the global skeleton has world x base); strong: `base` fragments in total
    total = base * world if args.scaling == "weak" else base
    sk = synth.draw_skeleton(cfg, n_frag=total)
    if world > 1:
        part = sharding.lpt_shard(sk.weights * args.chains, world)
        mine = np.nonzero(part == rank)[0]
        sk = sk.select(mine)
    p = synth.generate(cfg, skeleton=sk, seed=cfg.seed + 1000 * rank)
    return cfg, p


def config_of(cfg, args):
    """The workload description: identical in both arms (`--impl ours` and `--impl reference`)."""
    return {"workload": cfg.name, "fragments": args.n_frag or cfg.n_frag,
            "fragments_are": "per GPU" if (args.scaling == "weak" and cfg.sub_sizes != "single") else "total",
            "chains": args.chains, "iters_per_step": args.iters, "mode": args.mode,
            "l2": "inputs larger than L2 (CSR + chain state >> 126 MB); no flush"}


def reference_arm(args, rank, world):
    """The reference's CPU path on the host cores: the oracle port (Java cannot be built or run in this image),
    all host threads, on a bounded stratified sample of the same workload."""
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.use_native_build()
    cfg, p = build_workload(args, 0, 1)
    threads = os.cpu_count() or 1
    iters = args.iters
    strata = stratified_sample(p, args.chains, iters, args.cpu_iterations, threads)
    mode = args.ref_variant
    for _ in range(args.warmup):
        run_cpu(strata[0][4], args.chains, max(1, iters // 10), mode, threads)
    vals, tsum, rsum = [], 0.0, 0.0
    for _ in range(args.steps):
        v, r_s, t_s, desc = cpu_estimate(strata, args.chains, iters, mode, threads)
        vals.append(v)
        tsum += t_s
        rsum += r_s
    val = float(np.mean(vals))
    sample = (f"stratified by subproblem size, {sum(n for _, n, *_ in strata)} of {p.scalars['n_sub']} subproblems x "
              f"{args.chains} chains x {iters} iterations per step ({desc}); every stratum timed on all threads and scaled "
              f"by its share of the batch; oracle {mode} variant, -O3 -march=native")
    print(json.dumps({
        "impl": "reference", "metric": "MCMC fragment-reassignments/sec", "value": val,
        "unit": "fragment-reassignments/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tsum / max(args.steps, 1), "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_of(cfg, args),
        "reference_impl": f"C++ restatement of the Java sampler ({mode}), not Java: no JDK in this image",
        "cpu_baseline": {"value": val, "unit": "fragment-reassignments/s", "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": val, "unit": "fragment-reassignments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def load_json(path):
    try:
        with open(path) as fh:
            return json.load(fh)
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg4", choices=["cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the config's fragment count PER GPU; strong: in total, sharded over the GPUs")
    ap.add_argument("--n-frag", type=int, default=0, help="fragments (default: the config's)")
    ap.add_argument("--chains", type=int, default=0)
    ap.add_argument("--iters", type=int, default=0,
                    help="MH iterations per chain per step (default 10000; 1000000 for the giant component, SURVEY 8d; a full "
                         "reference run of the batch configs is 1e5: setup costs amortise with it)")
    ap.add_argument("--mode", default="fast", choices=["fast", "replay"])
    ap.add_argument("--cpu-iterations", type=float, default=2.0e9,
                    help="chain-iterations in the bounded CPU sample (about 10-30 s on the host cores)")
    ap.add_argument("--ref-variant", default="incremental", choices=["incremental", "faithful"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    from multibreak_sv_b200 import synth
    cfg0 = synth.CONFIGS[args.config]
    if not args.chains:
        args.chains = cfg0.n_chains
    if not args.iters:
        args.iters = 1_000_000 if cfg0.sub_sizes == "single" else 10_000
    giant = cfg0.sub_sizes == "single"
    if giant:
        args.scaling = "strong"          # the chains of ONE component are split over the GPUs: total work is fixed
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import multibreak_sv_b200 as mbsv
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the sampler has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(dev)

    cfg, p = build_workload(args, rank, world)
    sc = p.scalars
    mode = mbsv.MODE_FAST if args.mode == "fast" else mbsv.MODE_REPLAY_EXACT
    dp = mbsv.DeviceProblem(p, device=local_rank)
    # ONE contiguous tally buffer [aln_count | frag_err_count | cluster_hist]: what mbsv_reduce_tallies sums with one ncclReduce
    n_words = sc["n_aln"] + sc["n_frag"] + sc["n_hist"]
    tallies = torch.zeros(max(n_words, 1), dtype=torch.int32, device=dev)
    totals_t = torch.zeros(6, dtype=torch.int64, device=dev)
    base = tallies.data_ptr()
    ptrs = {"aln_count": base, "frag_err_count": base + 4 * sc["n_aln"],
            "cluster_hist": base + 4 * (sc["n_aln"] + sc["n_frag"]), "totals": totals_t.data_ptr()}

    my_first, my_chains = 0, args.chains
    comm = None
    if giant:
        from multibreak_sv_b200 import sharding
        my_first, my_chains = sharding.split_chains(args.chains, world, rank)
        if world > 1:
            def exchange(ident):          # rank 0's NCCL id to every rank (host-side plumbing)
                t = torch.from_numpy(ident.copy()).to(dev)
                dist.broadcast(t, src=0)
                return t.cpu().numpy()
            comm = mbsv.Comm(world, rank, local_rank, exchange)

    reduce_ms = []

    def step():
        st, ms = dp.run_device(ptrs, mode=mode, num_iters=args.iters, n_chains=my_chains, seed=123456 + my_first)
        if comm is not None:
            # the library's one ncclReduce of the integer tallies onto rank 0 (NVLink)
            reduce_ms.append(comm.reduce_tallies(base, n_words, 0))
        return ms

    for _ in range(args.warmup):
        step()
    reduce_ms.clear()
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    t0 = time.perf_counter()
    kernel_ms = 0.0
    launches = 0
    for _ in range(args.steps):
        kernel_ms += step()
        launches += int(dp.lib.mbsv_last_launch_count()) + (1 if comm is not None else 0)
    barrier()
    dt = time.perf_counter() - t0
    clk = clocks.stop()
    launch_info = dp.launch_info()
    if not clk.get("samples"):
        # the timed region was shorter than nvidia-smi's sampling period (giant component: tens of ms per step): sample
        # the clocks while the same step repeats, untimed, for about a second
        clocks = ClockSampler(local_rank)
        clocks.start()
        n_rep = torch.tensor([int(1.2 / max(dt / args.steps, 1e-3)) + 1], device=dev)
        if world > 1:
            dist.all_reduce(n_rep, op=dist.ReduceOp.MAX)      # the step holds a collective: every rank repeats it equally often
        for _ in range(int(n_rep.item())):
            step()
        barrier()
        clk = clocks.stop()
        clk["sampled"] = "untimed repeats of the step right after the timed region (the region is shorter than the sampling period)"
        reduce_ms[:] = reduce_ms[:args.steps]
    totals = totals_t.cpu().numpy().astype(np.int64)      # of the last step (every step is identical)
    if int(totals[5]) != 0:
        raise SystemExit(f"a chain reported status {int(totals[5])}")
    proposals, accepted, reassign = int(totals[0] + totals[2]), int(totals[1] + totals[3]), int(totals[4])

    stats = torch.tensor([dt, kernel_ms / 1e3, float(reassign), float(proposals), float(accepted)],
                         dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dt_max, kern_max = float(mx[0]), float(mx[1])
        reassign_all, prop_all, acc_all = float(sm[2]), float(sm[3]), float(sm[4])
    else:
        dt_max, kern_max, reassign_all, prop_all, acc_all = dt, kernel_ms / 1e3, reassign, proposals, accepted
    value = reassign_all * args.steps / dt_max

    # ---- end to end through the C ABI with host buffers: create (validation, packing, H2D) + run (D2H) + destroy ----
    e2e = None
    if not args.no_e2e:
        h2d = p.host_bytes
        d2h = 0
        barrier()
        t0 = time.perf_counter()
        e2e_launches = 0
        for _ in range(args.steps):
            d2 = mbsv.DeviceProblem(p, device=local_rank)
            out = d2.run(mode, args.iters, n_chains=my_chains, seed=123456 + my_first, want_state=False, per_chain=False)
            e2e_launches += out.launches
            d2h = out.result_bytes
            re2e = int(out.totals[4])
            d2.close()
            if comm is not None:
                # host tallies of every rank pooled on rank 0: H2D of the tallies + the library's reduce + D2H on rank 0
                t = torch.from_numpy(np.concatenate([out.aln_count, out.frag_err_count, out.cluster_hist]).view(np.int32)).to(dev)
                comm.reduce_tallies(t.data_ptr(), n_words, 0)
                if rank == 0:
                    t.cpu()
                e2e_launches += 1
        barrier()
        et = time.perf_counter() - t0
        es = torch.tensor([et, float(re2e)], dtype=torch.float64, device=dev)
        if world > 1:
            emx = es.clone(); dist.all_reduce(emx, op=dist.ReduceOp.MAX)
            esm = es.clone(); dist.all_reduce(esm, op=dist.ReduceOp.SUM)
            et, re2e = float(emx[0]), float(esm[1])
        e2e = {"value": re2e * args.steps / et, "unit": "fragment-reassignments/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * et / args.steps, "steps": args.steps}
        launches += e2e_launches

    # ---- roofline.  The chain state of this path is SM-resident (shared memory / L1 / L2), so HBM does not bind; what binds
    # is the SM's issue slots times the share of lanes doing useful work.  The issue-slot figures come from the ncu capture of
    # THIS workload (profiles/r02_issue_<config>.json, written by profiles/summarize_ncu.py from the .ncu-rep); the HBM
    # figure from SURVEY 8d's algorithmic bytes is kept beside it as `hbm_notional`. ----
    peaks = load_json(os.path.join(ROOT, "MEASURED_PEAKS.json"))
    if peaks:
        peak, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    alg = algorithmic_bytes(p, proposals, accepted)                      # this rank, per step
    achieved = alg / (kernel_ms / 1e3 / args.steps) / 1e9
    issue = load_json(os.path.join(ROOT, "profiles", f"r02_issue_{args.config}.json"))
    per_class = []
    tot_ms = sum(li["ms"] for li in launch_info) or 1.0
    for li in launch_info:
        kind = ("memo" if li["rows"] < 0 else "window" if (li["rows"] == 0 and li["smem_bytes"] > 32768)
                else "global" if li["rows"] == 0 else "smem")
        per_class.append({"kernel": kind, "rows": abs(li["rows"]), "warps": li["warps"], "ms": li["ms"],
                          "reassignments": li["reassignments"], "accepted": li["accepted"],
                          "reassignments_per_s_alone": li["reassignments"] / (li["ms"] / 1e3) if li["ms"] > 0 else None})
    if issue and issue.get("dominant"):
        d = issue["dominant"]
        roofline = {"bound": "sm-issue", "achieved": d["issue_active_pct"] * d["active_lanes"] / 32.0, "peak": 100.0,
                    "unit": "% of issue slots x active lanes / 32", "frac": d["issue_active_pct"] * d["active_lanes"] / 3200.0,
                    "traffic": d.get("dram_bytes_per_launch"), "kernel": d["kernel"], "kernels": issue.get("kernels"),
                    "source": issue.get("source")}
    else:
        roofline = {"bound": "sm-issue", "achieved": None, "peak": 100.0, "unit": "% of issue slots x active lanes / 32",
                    "frac": None, "traffic": None, "source": "no ncu capture of this config committed"}
    roofline["hbm_notional"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                                "peak_source": peak_src, "algorithmic_bytes_per_step": alg,
                                "kernel_ms_per_step": kernel_ms / args.steps,
                                "note": "SURVEY 8d algorithmic bytes / the library's CUDA-event time of the sampler launches; the "
                                        "bytes are served from shared memory / L1 / L2, measured DRAM traffic is `traffic`"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as orc
        orc.use_native_build()
        threads = os.cpu_count() or 1
        strata = stratified_sample(p, args.chains, args.iters, args.cpu_iterations, threads)
        v, r_s, t_s, desc = cpu_estimate(strata, args.chains, args.iters, "incremental", threads)
        cpu_baseline = {"value": v, "unit": "fragment-reassignments/s", "cores": threads, "kind": "port",
                        "sample": f"stratified by subproblem size, {sum(n for _, n, *_ in strata)} of {sc['n_sub']} subproblems "
                                  f"x {args.chains} chains x {args.iters} iterations ({int(r_s)} reassignments in {t_s:.1f} s: "
                                  f"{desc}); strata scaled by their share of the batch; oracle incremental variant, "
                                  f"-O3 -march=native"}

    if rank == 0:
        line = {
            "metric": "MCMC fragment-reassignments/sec", "value": value, "unit": "fragment-reassignments/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt_max / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(cfg, args),
            "sharding": ("chains split over the GPUs + one ncclReduce of the tallies (mbsv_reduce_tallies)" if giant else "lpt") if world > 1 else "none",
            "rank0": {"subproblems": sc["n_sub"], "fragments": sc["n_frag"], "alignments": sc["n_aln"], "chains": my_chains},
            "reduce_ms_per_step": (sum(reduce_ms) / len(reduce_ms)) if reduce_ms else None,
            "reduce_bytes": 4 * n_words if comm is not None else None,
            "iterations_per_s": (sc["n_sub"] * args.chains * args.iters) * args.steps / dt_max if world == 1 else None,
            "accepted_per_s": acc_all * args.steps / dt_max, "proposals_per_s": prop_all * args.steps / dt_max,
            "clocks": clk, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "per_class": per_class}
        print(json.dumps(line))
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
