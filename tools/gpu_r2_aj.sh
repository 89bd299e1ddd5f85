#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r2aj_variants.jsonl
python tools/quick_bench.py --config atrium --spp 64 --tag base 2>>gpurun_out/r2aj.err | tee -a gpurun_out/r2aj_variants.jsonl
python tools/quick_bench.py --config atrium --spp 64 --tag pktrim --lib bisemutum-engine_b200/csrc/_exp/libbpt_pktrim.so 2>>gpurun_out/r2aj.err | tee -a gpurun_out/r2aj_variants.jsonl
