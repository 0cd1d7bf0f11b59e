# Profiling recipe used for profiles/ (run under gpurun from the repo root): default bench, ncu launch list of the same
# command, then ncu --set full on a reduced size; summaries: python profiles/summarize_ncu.py gpurun_out/prof_r01_final.ncu-rep profiles/<name>
set -x
mkdir -p gpurun_out
SMALL="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --n-frag 1000000 --iters 1000"
python bench.py > gpurun_out/bench_default.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py > gpurun_out/ncu_launches.log 2>&1
$SMALL > gpurun_out/plain_small.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mbsv_ -s 13 -c 13 -o gpurun_out/prof_r01_final $SMALL > gpurun_out/ncu_full.log 2>&1
# summarise on the box: the report (~60 MB) would push gpurun_out/ over what is copied back
python profiles/summarize_ncu.py gpurun_out/prof_r01_final.ncu-rep gpurun_out/r01_ncu_full_final > gpurun_out/r01_ncu_dominant.json
python profiles/summarize_ncu.py gpurun_out/prof_r01_final.ncu-rep gpurun_out/r01_ncu_full_final_general 5 > /dev/null
rm -f gpurun_out/r01_ncu_full_final_general_metrics.csv gpurun_out/prof_r01_final.ncu-rep
tail -1 gpurun_out/bench_default.log | cut -c1-400
tail -2 gpurun_out/ncu_full.log | cut -c1-200
