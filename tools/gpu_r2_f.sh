#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r2f_variants.jsonl
BPT_SPECIALISE=0 python tools/quick_bench.py --spp 64 --tag general 2>>gpurun_out/r2f.err | tee -a gpurun_out/r2f_variants.jsonl
python tools/quick_bench.py --spp 64 --tag specialised 2>>gpurun_out/r2f.err | tee -a gpurun_out/r2f_variants.jsonl
python tools/quick_bench.py --spp 64 --tag specialised_again 2>>gpurun_out/r2f.err | tee -a gpurun_out/r2f_variants.jsonl
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/bench_configs.py 0_ 2_ 3_ 4_ > gpurun_out/r2f_configs.log 2>&1; cp gpurun_out/configs.json gpurun_out/r2f_configs.json; tail -5 gpurun_out/r2f_configs.log | cut -c1-400
