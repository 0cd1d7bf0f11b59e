# Round-2 profiling recipe (run under gpurun from the repo root).  Every ncu command runs only after the same command line
# exited 0 without ncu.  Summaries go to gpurun_out/ and are copied to profiles/ by hand.
set -x
mkdir -p gpurun_out
O=gpurun_out
# 1. the default bench (cfg4-mixed-5M, 64 chains, 10 000 iterations per step) and its launch list
python bench.py > $O/r02_bench_cfg4_1gpu.json 2> $O/r02_bench_cfg4_1gpu.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r02_launches_cfg4.csv python bench.py --no-e2e --no-cpu-baseline > $O/ncu_launches.log 2>&1
# 2. one step of the SAME batch (5M fragments, 64 chains) with 1000 iterations per chain under ncu --set full
SMALL="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --iters 1000"
$SMALL > $O/plain_small.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mbsv_ -s 13 -c 13 -o $O/prof_r02_cfg4 $SMALL > $O/ncu_full_cfg4.log 2>&1
python profiles/summarize_ncu.py $O/prof_r02_cfg4.ncu-rep $O/r02_ncu_cfg4 --issue $O/r02_issue_cfg4.json "one step of bench.py --iters 1000 on the full cfg4-mixed-5M batch, 64 chains" > $O/r02_ncu_cfg4_dominant.json
# the launches of the general kernel with the largest shared-memory class and of the global-workspace class
python - <<'PY'
import json, subprocess
d = json.load(open("gpurun_out/r02_issue_cfg4.json"))
ks = d["kernels"]
gl = [i for i, k in enumerate(ks) if "chain_kernel" in k["kernel"] and k["smem_bytes"] == 0]
sm = [i for i, k in enumerate(ks) if "chain_kernel" in k["kernel"] and k["smem_bytes"] > 0]
if gl:
    subprocess.run(["python", "profiles/summarize_ncu.py", "gpurun_out/prof_r02_cfg4.ncu-rep", "gpurun_out/r02_ncu_cfg4_global", str(gl[0])])
if sm:
    big = max(sm, key=lambda i: ks[i]["ms"])
    subprocess.run(["python", "profiles/summarize_ncu.py", "gpurun_out/prof_r02_cfg4.ncu-rep", "gpurun_out/r02_ncu_cfg4_smem", str(big)])
PY
rm -f $O/r02_ncu_cfg4_global_metrics.csv $O/r02_ncu_cfg4_smem_metrics.csv $O/prof_r02_cfg4.ncu-rep
# 3. the giant component (cfg5): bench, then the speculative-window kernel under ncu (100 000 iterations)
python bench.py --config cfg5 > $O/r02_bench_cfg5_1gpu.json 2> $O/r02_bench_cfg5_1gpu.err
G="python bench.py --config cfg5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --iters 100000"
$G > $O/plain_giant.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mbsv_giant_kernel -s 1 -c 1 -o $O/prof_r02_cfg5 $G > $O/ncu_full_cfg5.log 2>&1
python profiles/summarize_ncu.py $O/prof_r02_cfg5.ncu-rep $O/r02_ncu_cfg5 --issue $O/r02_issue_cfg5.json "one step of bench.py --config cfg5 --iters 100000 (64 chains, clusters of 2 CTAs)" > $O/r02_ncu_cfg5_dominant.json
rm -f $O/prof_r02_cfg5.ncu-rep
# 4. the other configurations and the reference arm
python bench.py --config cfg2 --iters 100000 > $O/r02_bench_cfg2_1gpu.json 2> $O/r02_bench_cfg2.err
python bench.py --config cfg3 --iters 100000 > $O/r02_bench_cfg3_1gpu.json 2> $O/r02_bench_cfg3.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_reference_cfg4.json 2> $O/r02_reference_cfg4.err
python bench.py --impl reference --steps 1 --warmup 0 --ref-variant faithful --cpu-iterations 2e8 > $O/r02_reference_cfg4_faithful.json 2> $O/r02_reference_cfg4_faithful.err
python bench.py --impl reference --config cfg5 --steps 2 --warmup 1 > $O/r02_reference_cfg5.json 2> $O/r02_reference_cfg5.err
for f in $O/r02_bench_*.json $O/r02_reference_*.json; do echo $f; cut -c1-330 $f; done
tail -2 $O/ncu_full_cfg4.log $O/ncu_full_cfg5.log | cut -c1-200
