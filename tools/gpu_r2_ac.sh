#!/bin/bash
# round 2, capture AC: which bounces walk the wide tree, re-measured on the final kernels (run-time switches)
mkdir -p gpurun_out; rm -f gpurun_out/r2ac_variants.jsonl
python tools/quick_bench.py --config atrium --spp 64 --tag base_e2_c2 2>>gpurun_out/r2ac.err | tee -a gpurun_out/r2ac_variants.jsonl
BPT_WIDE_FROM_BOUNCE=1 python tools/quick_bench.py --config atrium --spp 64 --tag e1_c2 2>>gpurun_out/r2ac.err | tee -a gpurun_out/r2ac_variants.jsonl
BPT_WIDE_CONNECT_FROM_BOUNCE=1 python tools/quick_bench.py --config atrium --spp 64 --tag e2_c1 2>>gpurun_out/r2ac.err | tee -a gpurun_out/r2ac_variants.jsonl
BPT_WIDE_FROM_BOUNCE=1 BPT_WIDE_CONNECT_FROM_BOUNCE=1 python tools/quick_bench.py --config atrium --spp 64 --tag e1_c1 2>>gpurun_out/r2ac.err | tee -a gpurun_out/r2ac_variants.jsonl
BPT_WIDE_FROM_BOUNCE=3 python tools/quick_bench.py --config atrium --spp 64 --tag e3_c2 2>>gpurun_out/r2ac.err | tee -a gpurun_out/r2ac_variants.jsonl
BPT_WIDE_CONNECT_FROM_BOUNCE=3 python tools/quick_bench.py --config atrium --spp 64 --tag e2_c3 2>>gpurun_out/r2ac.err | tee -a gpurun_out/r2ac_variants.jsonl
BPT_WIDE=0 python tools/quick_bench.py --config atrium --spp 64 --tag binary_only 2>>gpurun_out/r2ac.err | tee -a gpurun_out/r2ac_variants.jsonl
