#!/bin/bash
# round 2, capture AI: packet kernel: resident blocks, shadow-ray packets with the octant-specialised tests
mkdir -p gpurun_out; rm -f gpurun_out/r2ai_variants.jsonl
python tools/quick_bench.py --config atrium --spp 64 --tag pk10 2>>gpurun_out/r2ai.err | tee -a gpurun_out/r2ai_variants.jsonl
for v in pk8 pk12; do
  python tools/quick_bench.py --config atrium --spp 64 --tag $v --lib bisemutum-engine_b200/csrc/_exp/libbpt_$v.so 2>>gpurun_out/r2ai.err | tee -a gpurun_out/r2ai_variants.jsonl
done
BPT_PACKET=3 python tools/quick_bench.py --config atrium --spp 64 --tag packet3 2>>gpurun_out/r2ai.err | tee -a gpurun_out/r2ai_variants.jsonl
BPT_PACKET=3 python tools/quick_bench.py --config mixed --spp 8 --tag mixed_packet3 2>>gpurun_out/r2ai.err | tee -a gpurun_out/r2ai_variants.jsonl
