#!/bin/bash
# A/B of kernel variants on configs[1] (tools/quick_bench.py): shared-memory short stack sizes, in-block octant sort of the next rays
mkdir -p gpurun_out
for v in base s8 s12 s16 s24 oct oct_s16; do
  python tools/quick_bench.py --spp 64 --tag $v --lib bisemutum-engine_b200/csrc/_exp/libbpt_$v.so 2>>gpurun_out/r2b.err | tee -a gpurun_out/r2b_variants.jsonl
done
