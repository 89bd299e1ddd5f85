#!/bin/bash
# round 2, capture Q: wide-kernel phase thresholds (second sweep) + e2e wave schedules of a 20-frame job
mkdir -p gpurun_out; rm -f gpurun_out/r2q_variants.jsonl gpurun_out/r2q_e2e.jsonl
python tools/quick_bench.py --tag base 2>>gpurun_out/r2q.err | tee -a gpurun_out/r2q_variants.jsonl
for v in w12r20 w16r20 w16r24 w12r24 w20r24 w10r20 w12r20b12 w12r20b6 w12r20b8r8; do
  python tools/quick_bench.py --tag $v --lib bisemutum-engine_b200/csrc/_exp/libbpt_$v.so 2>>gpurun_out/r2q.err | tee -a gpurun_out/r2q_variants.jsonl
done
for w in 20 16,4 14,6 12,8 10,10 18,2 17,3 8,8,4; do
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-waves $w 2>>gpurun_out/r2q.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']
print(json.dumps({'waves':e['wave_schedule'],'e2e':e['value'],'ms_per_step':e['ms_per_step'],'value':d['value']}))" | tee -a gpurun_out/r2q_e2e.jsonl
done
