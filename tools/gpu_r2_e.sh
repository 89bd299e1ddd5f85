#!/bin/bash
# A/B: compile-time specialisation (no any-hit instances / no rect lights) and resident-block counts on top of it
mkdir -p gpurun_out
rm -f gpurun_out/r2e_variants.jsonl
BPT_SPECIALISE=0 python tools/quick_bench.py --spp 64 --tag general 2>>gpurun_out/r2e.err | tee -a gpurun_out/r2e_variants.jsonl
python tools/quick_bench.py --spp 64 --tag specialised 2>>gpurun_out/r2e.err | tee -a gpurun_out/r2e_variants.jsonl
for v in sb9 sb10 sb12 tb12; do
  python tools/quick_bench.py --spp 64 --tag $v --lib bisemutum-engine_b200/csrc/_exp/libbpt_$v.so 2>>gpurun_out/r2e.err | tee -a gpurun_out/r2e_variants.jsonl
done
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
