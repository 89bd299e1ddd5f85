#!/bin/bash
# round 2, capture X: predicated stack pushes in the wide step (only the kept children reach the L1)
mkdir -p gpurun_out; rm -f gpurun_out/r2x_variants.jsonl
for cfg in atrium instanced; do
  spp=64; [ $cfg = instanced ] && spp=16
  python tools/quick_bench.py --config $cfg --spp $spp --tag base 2>>gpurun_out/r2x.err | tee -a gpurun_out/r2x_variants.jsonl
  python tools/quick_bench.py --config $cfg --spp $spp --tag pushpred --lib bisemutum-engine_b200/csrc/_exp/libbpt_pushpred.so 2>>gpurun_out/r2x.err | tee -a gpurun_out/r2x_variants.jsonl
done
python tools/quick_bench.py --config atrium --accel two_level --spp 32 --tag base2l 2>>gpurun_out/r2x.err | tee -a gpurun_out/r2x_variants.jsonl
python tools/quick_bench.py --config atrium --accel two_level --spp 32 --tag pushpred2l --lib bisemutum-engine_b200/csrc/_exp/libbpt_pushpred.so 2>>gpurun_out/r2x.err | tee -a gpurun_out/r2x_variants.jsonl
