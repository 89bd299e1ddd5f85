#!/bin/bash
# round 2, capture V: octant bins (shade appends next rays' slots to per-octant lists; extend fetches through them)
mkdir -p gpurun_out; rm -f gpurun_out/r2v_variants.jsonl
for cfg in atrium instanced; do
  spp=64; [ $cfg = instanced ] && spp=16
  python tools/quick_bench.py --config $cfg --spp $spp --tag base 2>>gpurun_out/r2v.err | tee -a gpurun_out/r2v_variants.jsonl
  BPT_OCTANT_BINS=1 python tools/quick_bench.py --config $cfg --spp $spp --tag octbins 2>>gpurun_out/r2v.err | tee -a gpurun_out/r2v_variants.jsonl
done
BPT_OCTANT_BINS=1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
