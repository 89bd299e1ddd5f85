#!/bin/bash
# round 2, call A: GPU tests, driver-flag bench, steady-state bench, ncu launch list + --set full of the traversal/shade kernels of one wave
set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2a_gpu_tests.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_s20.json 2> gpurun_out/r2a_bench_s20.err
python bench.py --steps 128 --warmup 8 --no-cpu-baseline > gpurun_out/r2a_bench_s128.json 2> gpurun_out/r2a_bench_s128.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2a_ref.json 2> gpurun_out/r2a_ref.err
python bench.py --steps 16 --warmup 16 --device-only > gpurun_out/r2a_device_only_s16.json 2> gpurun_out/r2a_device_only_s16.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 16 --warmup 16 --device-only > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_trace_spec|k_shade" -s 21 -c 21 -o /tmp/r2a_kernels python bench.py --steps 16 --warmup 16 --device-only > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/r2a_kernels.ncu-rep --page raw --csv > gpurun_out/r2a_raw.csv 2> gpurun_out/ncu_export.err
ls -la /tmp/r2a_kernels.ncu-rep
tail -3 gpurun_out/r2a_gpu_tests.log
cat gpurun_out/r2a_bench_s20.json | head -c 3000
