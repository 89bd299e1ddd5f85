#!/bin/bash
# round 2, capture AD: packet traversal of the camera rays (and of their hit points' shadow rays): parity + A/B (BPT_PACKET bit 0 / bit 1)
mkdir -p gpurun_out; rm -f gpurun_out/r2ad_variants.jsonl
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for m in 0 1 3; do
  BPT_PACKET=$m python tools/quick_bench.py --config atrium --spp 64 --tag packet$m 2>>gpurun_out/r2ad.err | tee -a gpurun_out/r2ad_variants.jsonl
done
for m in 0 1 3; do
  BPT_PACKET=$m python tools/quick_bench.py --config mixed --spp 8 --tag mixed_packet$m 2>>gpurun_out/r2ad.err | tee -a gpurun_out/r2ad_variants.jsonl
done
