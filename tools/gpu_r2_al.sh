#!/bin/bash
# round 2, capture AL: wide kernels (merged mode): exact leaf box first, triangle fetched only behind it
mkdir -p gpurun_out; rm -f gpurun_out/r2al_variants.jsonl
python tools/quick_bench.py --config atrium --spp 64 --tag base 2>>gpurun_out/r2al.err | tee -a gpurun_out/r2al_variants.jsonl
python tools/quick_bench.py --config atrium --spp 64 --tag leafseq --lib bisemutum-engine_b200/csrc/_exp/libbpt_leafseq.so 2>>gpurun_out/r2al.err | tee -a gpurun_out/r2al_variants.jsonl
