#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r2j_variants.jsonl
python tools/quick_bench.py --config instanced --spp 16 --tag base 2>>gpurun_out/r2j.err | tee -a gpurun_out/r2j_variants.jsonl
for v in n20r24 n24r28 n20r20 n16r16 n28r30 n16r28 n24r24; do
  python tools/quick_bench.py --config instanced --spp 16 --tag $v --lib bisemutum-engine_b200/csrc/_exp/libbpt_$v.so 2>>gpurun_out/r2j.err | tee -a gpurun_out/r2j_variants.jsonl
done
tail -3 gpurun_out/r2j.err
python tools/quick_bench.py --accel two_level --spp 32 --tag atrium_two_level 2>>gpurun_out/r2j.err | tee -a gpurun_out/r2j_variants.jsonl
