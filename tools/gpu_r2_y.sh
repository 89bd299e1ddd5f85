#!/bin/bash
# round 2, capture Y: 256-bit loads for BVH nodes and exact leaf boxes (base = product build with them; noldg256 = four / two 128-bit loads)
mkdir -p gpurun_out; rm -f gpurun_out/r2y_variants.jsonl
for cfg in atrium instanced; do
  spp=64; [ $cfg = instanced ] && spp=16
  python tools/quick_bench.py --config $cfg --spp $spp --tag ldg256 2>>gpurun_out/r2y.err | tee -a gpurun_out/r2y_variants.jsonl
  python tools/quick_bench.py --config $cfg --spp $spp --tag noldg256 --lib bisemutum-engine_b200/csrc/_exp/libbpt_noldg256.so 2>>gpurun_out/r2y.err | tee -a gpurun_out/r2y_variants.jsonl
done
python tools/quick_bench.py --config atrium --accel two_level --spp 32 --tag ldg256_2l 2>>gpurun_out/r2y.err | tee -a gpurun_out/r2y_variants.jsonl
python tools/quick_bench.py --config atrium --accel two_level --spp 32 --tag noldg256_2l --lib bisemutum-engine_b200/csrc/_exp/libbpt_noldg256.so 2>>gpurun_out/r2y.err | tee -a gpurun_out/r2y_variants.jsonl
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
