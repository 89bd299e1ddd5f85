#!/bin/bash
# round 2, capture AB: lazy register top-of-stack for the wide kernels; streaming (evict-first) ray loads / hit stores
mkdir -p gpurun_out; rm -f gpurun_out/r2ab_variants.jsonl
python tools/quick_bench.py --config atrium --spp 64 --tag base 2>>gpurun_out/r2ab.err | tee -a gpurun_out/r2ab_variants.jsonl
for v in wtos stream wtos_stream; do
  python tools/quick_bench.py --config atrium --spp 64 --tag $v --lib bisemutum-engine_b200/csrc/_exp/libbpt_$v.so 2>>gpurun_out/r2ab.err | tee -a gpurun_out/r2ab_variants.jsonl
done
python tools/quick_bench.py --config instanced --spp 16 --tag base 2>>gpurun_out/r2ab.err | tee -a gpurun_out/r2ab_variants.jsonl
for v in wtos stream; do
  python tools/quick_bench.py --config instanced --spp 16 --tag $v --lib bisemutum-engine_b200/csrc/_exp/libbpt_$v.so 2>>gpurun_out/r2ab.err | tee -a gpurun_out/r2ab_variants.jsonl
done
