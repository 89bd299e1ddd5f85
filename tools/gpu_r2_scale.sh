#!/bin/bash
# round 2: the scaling lines on one box with $1 GPUs (default 8): the driver's command at N = 2, 4, (8); a steady-state line at the largest N;
# configs[3] as BASELINE names it (64 spp split across the GPUs = strong scaling, reduce of the 132.7 MB FP32 image included) next to the same job on 1 GPU
MAXN=${1:-8}
mkdir -p gpurun_out
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n "$@"; }
for n in 2 4 8; do
  [ $n -le $MAXN ] || continue
  run $n --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2h_bench_${n}gpu_s20.json 2> gpurun_out/r2h_bench_${n}gpu_s20.err
done
run $MAXN --steps 128 --warmup 8 --no-cpu-baseline > gpurun_out/r2h_bench_${MAXN}gpu_s128.json 2> gpurun_out/r2h_bench_${MAXN}gpu_s128.err
run $MAXN --config instanced --scaling strong --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/r2h_bench_instanced_${MAXN}gpu_strong64.json 2> gpurun_out/r2h_bench_instanced_${MAXN}gpu_strong64.err
python bench.py --config instanced --scaling strong --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/r2h_bench_instanced_1gpu_strong64.json 2> gpurun_out/r2h_bench_instanced_1gpu_strong64.err
for f in gpurun_out/r2h_bench_*gpu*.json; do echo $f; head -c 260 $f; echo; done
tail -n 3 gpurun_out/r2h_bench_${MAXN}gpu_s20.err
