#!/bin/bash
# round 2, capture AK: packets of 64 camera rays (two rays per lane, BPT_PACKET=9) against 32 (BPT_PACKET=1): parity + A/B
mkdir -p gpurun_out; rm -f gpurun_out/r2ak_variants.jsonl
BPT_PACKET=9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for m in 1 9; do
  BPT_PACKET=$m python tools/quick_bench.py --config atrium --spp 64 --tag packet$m 2>>gpurun_out/r2ak.err | tee -a gpurun_out/r2ak_variants.jsonl
done
