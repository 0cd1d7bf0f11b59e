# Round-2 multi-GPU lines (run under `gpurun --gpus 8` from the repo root): the giant component with its chains split over the
# GPUs and ONE ncclReduce inside the library, and the 5M-fragment batch as STRONG scaling (5M in total, LPT-sharded).
set -x
O=gpurun_out
mkdir -p $O
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
NCCL_DEBUG=INFO $TR bench.py --gpus $N --config cfg5 --no-cpu-baseline > $O/r02_cfg5_${N}gpu.json 2> $O/r02_cfg5_${N}gpu.err
grep -E "NVLS|NCCL INFO (Connected|comm 0x)" $O/r02_cfg5_${N}gpu.json $O/r02_cfg5_${N}gpu.err | head -20 > $O/r02_cfg5_${N}gpu_nccl.txt
$TR bench.py --gpus $N --scaling strong --no-cpu-baseline > $O/r02_cfg4_strong_${N}gpu.json 2> $O/r02_cfg4_strong_${N}gpu.err
for f in $O/r02_cfg5_${N}gpu.json $O/r02_cfg4_strong_${N}gpu.json; do echo $f; grep '^{' $f | cut -c1-700; done
for f in $O/*_${N}gpu.err; do tail -n 3 $f; done
