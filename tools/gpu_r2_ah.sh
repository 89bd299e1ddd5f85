#!/bin/bash
# round 2, capture AH: packet kernel with octant-specialised box tests: parity + A/B against k_trace_spec (BPT_PACKET=0)
mkdir -p gpurun_out; rm -f gpurun_out/r2ah_variants.jsonl
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for m in 0 1; do
  BPT_PACKET=$m python tools/quick_bench.py --config atrium --spp 64 --tag packet$m 2>>gpurun_out/r2ah.err | tee -a gpurun_out/r2ah_variants.jsonl
done
BPT_PACKET=1 python tools/quick_bench.py --config mixed --spp 8 --tag mixed_packet1 2>>gpurun_out/r2ah.err | tee -a gpurun_out/r2ah_variants.jsonl
