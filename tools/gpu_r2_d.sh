#!/bin/bash
# round 2, call D (2 GPUs): the N > 1 paths of bench.py (weak, strong, e2e with root-only read-back, reduce_check) + the multi-GPU check scripts
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2d_bench_2gpu_s20.json 2> gpurun_out/r2d_bench_2gpu_s20.err
$TR bench.py --gpus 2 --steps 128 --warmup 8 --no-cpu-baseline > gpurun_out/r2d_bench_2gpu_s128.json 2> gpurun_out/r2d_bench_2gpu_s128.err
$TR bench.py --gpus 2 --steps 256 --warmup 8 --no-cpu-baseline --scaling strong > gpurun_out/r2d_bench_2gpu_strong256.json 2> gpurun_out/r2d_bench_2gpu_strong256.err
python bench.py --gpus 1 --steps 256 --warmup 8 --no-cpu-baseline --scaling strong > gpurun_out/r2d_bench_1gpu_strong256.json 2> gpurun_out/r2d_bench_1gpu_strong256.err
$TR tests/check_reduce_multigpu.py > gpurun_out/r2d_check_reduce.log 2>&1
$TR tests/check_probes_sharded_multigpu.py > gpurun_out/r2d_check_probes.log 2>&1
for f in gpurun_out/r2d_bench_*.json; do echo $f; head -c 600 $f; echo; done
tail -3 gpurun_out/r2d_check_reduce.log gpurun_out/r2d_check_probes.log
tail -5 gpurun_out/r2d_bench_2gpu_s20.err
