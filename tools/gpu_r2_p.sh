#!/bin/bash
# round 2, capture P: tapered e2e wave schedule (bench.py) + traversal variants (triangle-phase loop, third postponed leaf, wide-kernel thresholds)
mkdir -p gpurun_out; rm -f gpurun_out/r2p_variants.jsonl
python bench.py --steps 20 --warmup 5 > gpurun_out/r2p_bench_s20.json 2> gpurun_out/r2p_bench_s20.err
python tools/quick_bench.py --tag base 2>>gpurun_out/r2p.err | tee -a gpurun_out/r2p_variants.jsonl
for v in triloop leaf3 leaf3tl w6r12 w12r16 w8r16 w8r8 w12r20; do
  python tools/quick_bench.py --tag $v --lib bisemutum-engine_b200/csrc/_exp/libbpt_$v.so 2>>gpurun_out/r2p.err | tee -a gpurun_out/r2p_variants.jsonl
done
python tools/quick_bench.py --tag base_again 2>>gpurun_out/r2p.err | tee -a gpurun_out/r2p_variants.jsonl
python -c "
import json
d=json.load(open('gpurun_out/r2p_bench_s20.json'))
print('value',d['value'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],d['e2e'].get('wave_schedule'),'steady',d['e2e'].get('steady_state_128_steps'))
"
