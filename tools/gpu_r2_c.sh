#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2c_gpu_tests.log 2>&1
tail -15 gpurun_out/r2c_gpu_tests.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench_s20.json 2> gpurun_out/r2c_bench_s20.err
head -c 1500 gpurun_out/r2c_bench_s20.json
