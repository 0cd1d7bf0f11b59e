# Final round-2 profiling pass (run under gpurun from the repo root) for the kernels as committed.  Every ncu command runs only
# after the same command line exited 0 without ncu.  Summaries go to gpurun_out/ and are copied to profiles/ by hand.
set -x
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/final_pytest.log 2>&1; tail -3 $O/final_pytest.log
# 1. the default bench (cfg4-mixed-5M, 64 chains, 10 000 iterations per step) and its launch list
python bench.py > $O/r02_bench_cfg4_1gpu.json 2> $O/r02_bench_cfg4_1gpu.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r02_launches_cfg4.csv python bench.py --no-e2e --no-cpu-baseline > $O/ncu_launches.log 2>&1
# 2. one step of the SAME batch (5M fragments, 64 chains) with 1000 iterations per chain under ncu --set full
#    (launches of one step: table build + 12 size classes; the warm-up step has one more, the alignment-slot records)
SMALL="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --iters 1000"
$SMALL > $O/plain_small.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mbsv_ -s 14 -c 13 -o $O/prof_r02_cfg4 $SMALL > $O/ncu_full_cfg4.log 2>&1
python profiles/summarize_ncu.py $O/prof_r02_cfg4.ncu-rep $O/r02_ncu_cfg4 --issue $O/r02_issue_cfg4.json "one step of bench.py --iters 1000 on the full cfg4-mixed-5M batch, 64 chains" > $O/r02_ncu_cfg4_dominant.json
python - <<'PY'
import json, subprocess
d = json.load(open("gpurun_out/r02_issue_cfg4.json"))
ks = d["kernels"]
gl = [i for i, k in enumerate(ks) if "chain_kernel" in k["kernel"] and k["state_rows"] == 0]
sm = [i for i, k in enumerate(ks) if "chain_kernel" in k["kernel"] and k["state_rows"] > 0]
if gl:
    subprocess.run(["python", "profiles/summarize_ncu.py", "gpurun_out/prof_r02_cfg4.ncu-rep", "gpurun_out/r02_ncu_cfg4_global", str(gl[0])])
if sm:
    big = max(sm, key=lambda i: ks[i]["ms"])
    subprocess.run(["python", "profiles/summarize_ncu.py", "gpurun_out/prof_r02_cfg4.ncu-rep", "gpurun_out/r02_ncu_cfg4_smem", str(big)])
PY
rm -f $O/r02_ncu_cfg4_global_metrics.csv $O/r02_ncu_cfg4_smem_metrics.csv $O/prof_r02_cfg4.ncu-rep
# 3. the other configurations
python bench.py --config cfg5 > $O/r02_bench_cfg5_1gpu.json 2> $O/r02_bench_cfg5_1gpu.err
python bench.py --config cfg2 --iters 100000 > $O/r02_bench_cfg2_1gpu.json 2> $O/r02_bench_cfg2.err
python bench.py --config cfg3 --iters 100000 > $O/r02_bench_cfg3_1gpu.json 2> $O/r02_bench_cfg3.err
for f in $O/r02_bench_*.json; do echo $f; cut -c1-330 $f; done
tail -n 2 $O/ncu_full_cfg4.log | cut -c1-200
