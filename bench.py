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


def cpu_sample(p, n_chains: int, iters: int, target_iterations: float):
    """A prefix of the batch's subproblems holding about target_iterations chain-iterations."""
    S = p.scalars["n_sub"]
    n = int(min(S, max(1, round(target_iterations / (n_chains * max(iters, 1))))))
    return slice_subproblems(p, 0, n), n


def run_cpu(sample, n_chains, iters, mode, threads):
    from oracle import oracle as orc
    op = orc.Problem(sample.scalars, sample.arrays)
    t0 = time.perf_counter()
    r = orc.run(op, mode, iters, n_chains=n_chains, java_int_wrap=False, nthreads=threads)
    dt = time.perf_counter() - t0
    return int(r.totals[4]), dt


def algorithmic_bytes(p, proposals: float, accepted: float) -> float:
    """(52 + 6k) B per proposal + (28 + 26k) B per accepted move, k = ESPs in old + new alignment."""
    k = 2.0 * p.scalars["n_aln_esp"] / p.scalars["n_aln"]
    return proposals * (52 + 6 * k) + accepted * (28 + 26 * k)


def build_workload(args, rank, world):
    from multibreak_sv_b200 import sharding, synth
    cfg = synth.CONFIGS[args.config]
    per_gpu = args.n_frag if args.n_frag else cfg.n_frag
    if cfg.sub_sizes == "single":
        # giant component (BASELINE configs[4]): cannot be split spatially; every rank holds the same subproblem and
        # runs its share of the chains (sharding.split_chains), tallies are summed with one reduce
        return cfg, synth.generate(cfg, n_frag=per_gpu)
    sk = synth.draw_skeleton(cfg, n_frag=per_gpu * world)
    if world > 1:
        part = sharding.lpt_shard(sk.weights * args.chains, world)
        mine = np.nonzero(part == rank)[0]
        sk = sk.select(mine)
    p = synth.generate(cfg, skeleton=sk, seed=cfg.seed + 1000 * rank)
    return cfg, p


def reference_arm(args, rank, world):
    """The reference's CPU path on the host cores: the oracle port (Java cannot be built or run in this image),
    all host threads, on a bounded sample of the same workload."""
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.build()
    cfg, p = build_workload(args, 0, 1)
    threads = os.cpu_count() or 1
    iters = args.iters
    sample, nsub = cpu_sample(p, args.chains, iters, args.cpu_iterations)
    mode = args.ref_variant
    for _ in range(args.warmup):
        run_cpu(sample, args.chains, max(1, iters // 10), mode, threads)
    tot, tsum = 0, 0.0
    for _ in range(args.steps):
        n, dt = run_cpu(sample, args.chains, iters, mode, threads)
        tot += n
        tsum += dt
    val = tot / tsum
    desc = (f"first {nsub} of {p.scalars['n_sub']} subproblems of {cfg.name} x {args.chains} chains x {iters} "
            f"iterations per step")
    print(json.dumps({
        "impl": "reference", "metric": "MCMC fragment-reassignments/sec", "value": val,
        "unit": "fragment-reassignments/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tsum / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg.name, "chains": args.chains, "iters_per_step": iters,
                   "reference_impl": f"C++ restatement of the Java sampler ({mode}), not Java: no JDK in this image"},
        "cpu_baseline": {"value": val, "unit": "fragment-reassignments/s", "cores": threads, "kind": "port",
                         "sample": desc},
        "e2e": {"value": val, "unit": "fragment-reassignments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg4")
    ap.add_argument("--n-frag", type=int, default=0, help="fragments per GPU (default: the config's)")
    ap.add_argument("--chains", type=int, default=0)
    ap.add_argument("--iters", type=int, default=10000,
                    help="MH iterations per chain per step (a full reference run is 1e5; setup costs amortise with it)")
    ap.add_argument("--mode", default="fast", choices=["fast", "replay"])
    ap.add_argument("--cpu-iterations", type=float, default=2.0e9,
                    help="chain-iterations in the bounded CPU sample (about 10-30 s on the host cores)")
    ap.add_argument("--ref-variant", default="incremental", choices=["incremental", "faithful"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    from multibreak_sv_b200 import synth
    if not args.chains:
        args.chains = synth.CONFIGS[args.config].n_chains
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
    t_u32 = lambda n: torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
    bufs = {"aln_count": t_u32(sc["n_aln"]), "frag_err_count": t_u32(sc["n_frag"]), "cluster_hist": t_u32(sc["n_hist"]),
            "totals": torch.zeros(6, dtype=torch.int64, device=dev)}
    ptrs = {k: v.data_ptr() for k, v in bufs.items()}

    giant = cfg.sub_sizes == "single"
    my_first, my_chains = 0, args.chains
    if giant:
        from multibreak_sv_b200 import sharding
        my_first, my_chains = sharding.split_chains(args.chains, world, rank)

    def step():
        st, ms = dp.run_device(ptrs, mode=mode, num_iters=args.iters, n_chains=my_chains, seed=123456 + my_first)
        if giant and world > 1:
            # one reduce of the integer tallies onto rank 0 (NCCL over NVLink)
            for k in ("aln_count", "frag_err_count", "cluster_hist"):
                dist.reduce(bufs[k], dst=0, op=dist.ReduceOp.SUM)
        return ms

    for _ in range(args.warmup):
        step()
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    t0 = time.perf_counter()
    kernel_ms = 0.0
    launches = 0
    for _ in range(args.steps):
        kernel_ms += step()
        launches += int(dp.lib.mbsv_last_launch_count())
    barrier()
    dt = time.perf_counter() - t0
    clk = clocks.stop()
    launch_info = dp.launch_info()
    totals = bufs["totals"].cpu().numpy().astype(np.int64)      # of the last step (every step is identical)
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

    # ---- end to end through the C ABI with host buffers ----
    e2e = None
    if not args.no_e2e:
        h2d = p.host_bytes
        d2h = 0
        barrier()
        t0 = time.perf_counter()
        e2e_launches = 0
        e2e_steps = max(1, min(args.steps, 3))      # every step repeats the same create + run + destroy; 3 bound the run time
        for _ in range(e2e_steps):
            d2 = mbsv.DeviceProblem(p, device=local_rank)
            out = d2.run(mode, args.iters, n_chains=my_chains, seed=123456 + my_first, want_state=False, per_chain=False)
            e2e_launches += out.launches
            d2h = out.result_bytes
            if int(out.totals[5]) != 0:
                raise SystemExit("e2e: chain status != 0")
            re2e = int(out.totals[4])
            d2.close()
        barrier()
        et = time.perf_counter() - t0
        es = torch.tensor([et, float(re2e)], dtype=torch.float64, device=dev)
        if world > 1:
            emx = es.clone(); dist.all_reduce(emx, op=dist.ReduceOp.MAX)
            esm = es.clone(); dist.all_reduce(esm, op=dist.ReduceOp.SUM)
            et, re2e = float(emx[0]), float(esm[1])
        e2e = {"value": re2e * e2e_steps / et, "unit": "fragment-reassignments/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * et / e2e_steps, "steps": e2e_steps}
        launches += e2e_launches

    # ---- roofline of the sampler launches (one kernel template, one launch per state-size class) ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    alg = algorithmic_bytes(p, proposals, accepted)                      # this rank, per step
    achieved = alg / (kernel_ms / 1e3 / args.steps) / 1e9
    # DRAM traffic comes from an `ncu --set full` capture (profiles/traffic.json); it only fills `traffic` when the
    # capture was taken on this same workload, otherwise it is reported beside it with its own configuration.
    traffic, traffic_other = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            if tj.get("n_frag") == sc["n_frag"] and tj.get("iters") == args.iters and tj.get("chains") == args.chains:
                traffic = tj.get("dram_bytes_per_launch")
            else:
                traffic_other = tj
        except Exception:
            pass
    issue = None
    ipath = os.path.join(ROOT, "profiles", "issue.json")
    if os.path.exists(ipath):
        try:
            issue = json.load(open(ipath))
        except Exception:
            issue = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_reduced_config": traffic_other, "peak_source": peak_src,
                "kernel": "mbsv_memo_kernel + mbsv_chain_kernel<FAST> (one launch per state-size class, concurrent streams)",
                "kernel_ms_per_step": kernel_ms / args.steps, "algorithmic_bytes_per_step": alg,
                "binding": issue,
                "note": "algorithmic bytes are served from shared memory / L1 / L2 (chain state is SM-resident), so "
                        "HBM is not the binding roofline; `binding` holds the ncu issue-slot figures of the dominant "
                        "launch (profiles/), see DESIGN.md section 4"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as orc
        orc.build()
        threads = os.cpu_count() or 1
        sample, nsub = cpu_sample(p, args.chains, args.iters, args.cpu_iterations)
        n, cdt = run_cpu(sample, args.chains, args.iters, "incremental", threads)
        cpu_baseline = {"value": n / cdt, "unit": "fragment-reassignments/s", "cores": threads, "kind": "port",
                        "sample": f"first {nsub} of {sc['n_sub']} subproblems x {args.chains} chains x {args.iters} "
                                  f"iterations ({n} reassignments in {cdt:.1f} s), oracle incremental variant"}

    if rank == 0:
        line = {
            "metric": "MCMC fragment-reassignments/sec", "value": value, "unit": "fragment-reassignments/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg.name, "fragments_per_gpu": sc["n_frag"] if world == 1 else args.n_frag or cfg.n_frag,
                       "subproblems_rank0": sc["n_sub"], "alignments_rank0": sc["n_aln"], "chains": args.chains,
                       "iters_per_step": args.iters, "mode": args.mode,
                       "sharding": ("chains split + reduce" if giant else "lpt") if world > 1 else "none",
                       "l2": "inputs larger than L2 (CSR + chain state >> 126 MB); no flush"},
            "iterations_per_s": (sc["n_sub"] * args.chains * args.iters) * args.steps / dt_max if world == 1 else None,
            "accepted_per_s": acc_all * args.steps / dt_max, "proposals_per_s": prop_all * args.steps / dt_max,
            "clocks": clk, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "launch_info": launch_info}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
