#!/bin/bash
# round 2, capture T: GPU tests on the re-tuned thresholds + fused colour target, bench lines
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2t_gpu_tests.log; cat gpurun_out/r2t_gpu_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2t_bench_s20.json 2> gpurun_out/r2t_bench_s20.err
python -c "
import json
d=json.load(open('gpurun_out/r2t_bench_s20.json'))
print('value',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],d['e2e'].get('wave_schedule'),'steady',d['e2e'].get('steady_state_128_steps'), d['kernel_ms_per_step'])
"
