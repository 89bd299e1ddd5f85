#!/bin/bash
# round 2, capture U: prefetch of the pushed nodes in the wide step (L2 / L1), atrium + instanced
mkdir -p gpurun_out; rm -f gpurun_out/r2u_variants.jsonl
for cfg in atrium instanced; do
  spp=64; [ $cfg = instanced ] && spp=16
  python tools/quick_bench.py --config $cfg --spp $spp --tag base 2>>gpurun_out/r2u.err | tee -a gpurun_out/r2u_variants.jsonl
  for v in pfl2 pfl1; do
    python tools/quick_bench.py --config $cfg --spp $spp --tag $v --lib bisemutum-engine_b200/csrc/_exp/libbpt_$v.so 2>>gpurun_out/r2u.err | tee -a gpurun_out/r2u_variants.jsonl
  done
done
