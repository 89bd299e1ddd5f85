#!/bin/bash
# round 2, capture AM: resident blocks of the two-level wide kernels without any-hit instances (8 = 64 registers today) on the instanced scene
mkdir -p gpurun_out; rm -f gpurun_out/r2am_variants.jsonl
python tools/quick_bench.py --config instanced --spp 16 --tag 2l8 2>>gpurun_out/r2am.err | tee -a gpurun_out/r2am_variants.jsonl
for v in 2l9 2l7; do
  python tools/quick_bench.py --config instanced --spp 16 --tag $v --lib bisemutum-engine_b200/csrc/_exp/libbpt_$v.so 2>>gpurun_out/r2am.err | tee -a gpurun_out/r2am_variants.jsonl
done
