#!/bin/bash
mkdir -p gpurun_out
for k in 0 1 0 1; do
  BENCH_E2E_PRETOUCH=$k BENCH_E2E_TRACE=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-waves 16,4 2>&1 >/dev/null | grep -E "e2e tail|total|frame  19|frame 127" 
done > gpurun_out/r2s_tail.txt
cat gpurun_out/r2s_tail.txt
