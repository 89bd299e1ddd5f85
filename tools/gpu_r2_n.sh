#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_trace_spec" -s 14 -c 2 -o /tmp/r2n_inst python bench.py --config instanced --steps 8 --warmup 8 --device-only > gpurun_out/ncu_n.log 2>&1
python tools/summarize_ncu.py source /tmp/r2n_inst.ncu-rep > gpurun_out/r2n_instanced_source.md 2>&1
ncu -i /tmp/r2n_inst.ncu-rep --page raw --csv > gpurun_out/r2n_inst_raw.csv 2>/dev/null
python tools/summarize_ncu.py full gpurun_out/r2n_inst_raw.csv > gpurun_out/r2n_instanced_kernels.md 2>&1
head -45 gpurun_out/r2n_instanced_source.md | cut -c1-210
