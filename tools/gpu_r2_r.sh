#!/bin/bash
# round 2, capture R: combined thresholds; where the e2e timed region goes (per-frame GPU events)
mkdir -p gpurun_out; rm -f gpurun_out/r2r_variants.jsonl
python tools/quick_bench.py --tag base 2>>gpurun_out/r2r.err | tee -a gpurun_out/r2r_variants.jsonl
for v in w16r24b8r8 w16r24b8r10 w16r24b6r8 w16r24b8r6; do
  python tools/quick_bench.py --tag $v --lib bisemutum-engine_b200/csrc/_exp/libbpt_$v.so 2>>gpurun_out/r2r.err | tee -a gpurun_out/r2r_variants.jsonl
done
for w in 20 16,4; do
  BENCH_E2E_TRACE=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-waves $w > gpurun_out/r2r_e2e_$w.json 2> gpurun_out/r2r_e2e_trace_$w.txt
done
