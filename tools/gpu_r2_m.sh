#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r2m_variants.jsonl
python tools/quick_bench.py --config instanced --spp 16 --tag p16r20 2>>gpurun_out/r2m.err | tee -a gpurun_out/r2m_variants.jsonl
for v in p12r16 p20r24 p16r24 p8r12 p24r28 p12r24; do
  python tools/quick_bench.py --config instanced --spp 16 --tag $v --lib bisemutum-engine_b200/csrc/_exp/libbpt_$v.so 2>>gpurun_out/r2m.err | tee -a gpurun_out/r2m_variants.jsonl
done
