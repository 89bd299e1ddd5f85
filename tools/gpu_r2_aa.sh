#!/bin/bash
# round 2, capture AA: shared-memory short stack re-measured now that the wide kernels are L1-request-bound (predicated pushes, 256-bit loads)
mkdir -p gpurun_out; rm -f gpurun_out/r2aa_variants.jsonl
python tools/quick_bench.py --config atrium --spp 64 --tag base 2>>gpurun_out/r2aa.err | tee -a gpurun_out/r2aa_variants.jsonl
for v in smem4 smem8 smem12 smem16; do
  python tools/quick_bench.py --config atrium --spp 64 --tag $v --lib bisemutum-engine_b200/csrc/_exp/libbpt_$v.so 2>>gpurun_out/r2aa.err | tee -a gpurun_out/r2aa_variants.jsonl
done
python tools/quick_bench.py --config instanced --spp 16 --tag base 2>>gpurun_out/r2aa.err | tee -a gpurun_out/r2aa_variants.jsonl
for v in smem8 smem16; do
  python tools/quick_bench.py --config instanced --spp 16 --tag $v --lib bisemutum-engine_b200/csrc/_exp/libbpt_$v.so 2>>gpurun_out/r2aa.err | tee -a gpurun_out/r2aa_variants.jsonl
done
