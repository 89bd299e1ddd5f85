#!/bin/bash
# round 2, capture AN: two-level packet traversal of the camera rays (BPT_PACKET bit 4): parity + A/B on the instanced scene and the two-level atrium
mkdir -p gpurun_out; rm -f gpurun_out/r2an_variants.jsonl
BPT_PACKET=17 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for m in 1 17; do
  BPT_PACKET=$m python tools/quick_bench.py --config instanced --spp 8 --tag inst_packet$m 2>>gpurun_out/r2an.err | tee -a gpurun_out/r2an_variants.jsonl
done
for m in 1 17; do
  BPT_PACKET=$m python tools/quick_bench.py --config atrium --accel two_level --spp 32 --tag atrium2l_packet$m 2>>gpurun_out/r2an.err | tee -a gpurun_out/r2an_variants.jsonl
done
