#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full) into the small text summaries kept under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof_r01b.ncu-rep profiles/r01_ncu_full [launch-index]

writes <out>_metrics.csv (one row per captured launch), <out>_lines.txt (CUDA source lines ranked by executed
warp-instructions with their average active lanes) and prints the dominant kernel's DRAM traffic per launch.
"""
import collections
import csv
import json
import subprocess
import sys

METRICS = [
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "gpu__time_duration.sum",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__sass_average_branch_targets_threads_uniform.pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__average_warp_latency_per_inst_issued.ratio",
]
STALLS = ["wait", "no_instruction", "branch_resolving", "long_scoreboard", "short_scoreboard", "not_selected",
          "math_pipe_throttle", "dispatch_stall", "lg_throttle", "mio_throttle", "barrier", "selected"]


def state_rows(name, shm):
    """Rows of a chain's state tile from the launch's dynamic shared memory: 4 warps x rows x 32 state words of 4 bytes (1 byte in the kern_fast8.cu instantiations); the general kernel's
    SMTOT instantiations (last template argument) keep 2560 + 4096 * MAXE bytes of counters and totals behind the tiles."""
    if "giant" in name or not shm:
        return 0
    if "mbsv_chain_kernel<" in name:
        args = [a.strip() for a in name.split("mbsv_chain_kernel<")[1].split(">")[0].split(",")]
        if len(args) >= 8 and args[7] in ("1", "true"):
            shm -= 2560 + 4096 * int(args[3])
        if args[2] in ("0", "false"):
            return 0
        if len(args) >= 9 and "char" in args[8]:          # 8-bit state words (kern_fast8.cu)
            return int(shm // 128)
    return int(shm // 512)


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = ["Kernel Name"] + [m for m in METRICS if m in idx] + \
           [f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio" for s in STALLS
            if f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio" in idx]
    with open(out + "_metrics.csv", "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(cols)
        w.writerow([units[idx[c]] for c in cols])
        for r in data:
            w.writerow([r[idx[c]] for c in cols])
    # ---- per-kernel table of the binding bound (SM issue slots x lane utilisation), consumed by bench.py as roofline ----
    # python profiles/summarize_ncu.py <rep> <out> --issue profiles/r02_issue_<config>.json "<what was captured>"
    if "--issue" in sys.argv:
        ia = sys.argv.index("--issue")
        issue_path, what = sys.argv[ia + 1], (sys.argv[ia + 2] if len(sys.argv) > ia + 2 else "")
        del sys.argv[ia:ia + 3]

        def num(r, m):
            try:
                v = float(r[idx[m]])
                return None if v != v else v          # ncu prints n/a -> nan for counters it could not collect
            except (KeyError, ValueError):
                return None

        def to_b(r, m):
            return float(r[idx[m]]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[idx[m]]]
        dur_unit = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
        du = dur_unit.get(units[idx["gpu__time_duration.sum"]], 1.0)
        ks = []
        for r in data:
            name = r[idx["Kernel Name"]]
            if "mbsv_" not in name:
                continue
            shm = num(r, "launch__shared_mem_per_block_dynamic") or 0.0
            shm *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6}.get(units[idx["launch__shared_mem_per_block_dynamic"]].split("/")[0], 1)
            issue, lanes = num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"), num(r, "smsp__thread_inst_executed_per_inst_executed.ratio")
            ks.append({"kernel": name.replace("void ", "").split("(")[0], "grid": int(num(r, "launch__grid_size")),
                       "block": int(num(r, "launch__block_size")), "smem_bytes": int(shm),
                       "state_rows": state_rows(name, shm),
                       "ms": num(r, "gpu__time_duration.sum") * du, "issue_active_pct": issue, "active_lanes": lanes,
                       "frac": issue * lanes / 3200.0 if issue is not None and lanes is not None else None, "warps_active_per_smsp": num(r, "smsp__warps_active.avg.per_cycle_active"),
                       "long_scoreboard_per_issue": num(r, "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
                       "barrier_per_issue": num(r, "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
                       "registers": int(num(r, "launch__registers_per_thread")),
                       "dram_bytes_per_launch": to_b(r, "dram__bytes_read.sum") + to_b(r, "dram__bytes_write.sum")})
        tot = sum(k["ms"] for k in ks) or 1.0
        for k in ks:
            k["share_of_kernel_time"] = k["ms"] / tot
        dom = max((k for k in ks if k["frac"] is not None), key=lambda k: k["ms"])
        with open(issue_path, "w") as fh:
            json.dump({"source": f"ncu --set full --clock-control none ({what}); serialised cold-cache launches, shares by gpu__time_duration; "
                                 f"written by profiles/summarize_ncu.py", "bound": "SM issue slots x active lanes / 32",
                       "dominant": dom, "kernels": ks}, fh, indent=1)
    # dominant launch = the one that executed the most warp-instructions (the bulk of the chains)
    inst = [float(r[idx["smsp__inst_executed.sum"]]) for r in data]
    k = max(range(len(data)), key=lambda i: inst[i])
    if len(sys.argv) > 3:          # source lines of another captured launch (e.g. one of the general-kernel classes)
        k = int(sys.argv[3])

    def to_bytes(v, u):
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    traffic = sum(to_bytes(data[k][idx[m]], units[idx[m]]) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    print(json.dumps({"dominant_launch": k, "kernel": data[k][idx["Kernel Name"]], "grid": data[k][idx["launch__grid_size"]],
                      "duration": data[k][idx["gpu__time_duration.sum"]] + " " + units[idx["gpu__time_duration.sum"]],
                      "dram_bytes_per_launch": traffic}))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(k),
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    agg = collections.OrderedDict()
    cur_file = cur_line = cur_src = None
    h = None
    for r in csv.reader(src.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            h = r
            iE, iT, iS = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
            continue
        if h is None:
            continue
        if r[0] != "":
            cur_line, cur_src = r[0], r[1]
        try:
            e, t, s = int(r[iE] or 0), int(r[iT] or 0), int(r[iS] or 0)
        except (ValueError, IndexError):
            continue
        a = agg.setdefault((cur_file, cur_line), [0, 0, 0, cur_src])
        a[0] += e; a[1] += t; a[2] += s
    tot = sum(a[0] for a in agg.values()) or 1
    tott = sum(a[1] for a in agg.values())
    tots = sum(a[2] for a in agg.values()) or 1
    with open(out + "_lines.txt", "w") as fh:
        fh.write(f"{'dominant ' if len(sys.argv) <= 3 else ''}launch #{k}: {data[k][idx['Kernel Name']]} grid {data[k][idx['launch__grid_size']]}\n")
        fh.write(f"warp-instructions {tot}  thread-instructions {tott}  avg active lanes {tott / tot:.2f}\n")
        fh.write("share-of-warp-inst  share-of-stall-samples  avg-active-lanes  file:line  source\n")
        for (f, l), (e, t, s, sline) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:60]:
            fh.write(f"{100 * e / tot:5.1f}%  {100 * s / tots:5.1f}%  {t / max(e, 1):5.1f}  {f}:{l}  {sline.strip()[:110]}\n")


if __name__ == "__main__":
    main()
