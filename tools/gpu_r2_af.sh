#!/bin/bash
# round 2, capture AF: frustum traversal of the camera rays (BPT_PACKET=4) against per-lane packets (1) and k_trace_spec (0): parity + A/B
mkdir -p gpurun_out; rm -f gpurun_out/r2af_variants.jsonl
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for m in 0 1 4; do
  BPT_PACKET=$m python tools/quick_bench.py --config atrium --spp 64 --tag packet$m 2>>gpurun_out/r2af.err | tee -a gpurun_out/r2af_variants.jsonl
done
BPT_PACKET=4 python tools/quick_bench.py --config mixed --spp 8 --tag mixed_packet4 2>>gpurun_out/r2af.err | tee -a gpurun_out/r2af_variants.jsonl
