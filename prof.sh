set -x
mkdir -p gpurun_out
SMALL="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --n-frag 1000000 --iters 500"
$SMALL > gpurun_out/plain_small.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mbsv_ -s 12 -c 12 -o gpurun_out/prof_r01d $SMALL > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log | cut -c1-300
