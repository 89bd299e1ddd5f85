#!/bin/bash
# source-level view of the first two shade launches and the first two extend launches of a 16-sample wave (summarised on the box: the reports are too large to bring back)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_shade" -s 7 -c 2 -o /tmp/r2k_shade python bench.py --steps 16 --warmup 16 --device-only > gpurun_out/ncu_k.log 2>&1
python tools/summarize_ncu.py source /tmp/r2k_shade.ncu-rep > gpurun_out/r2k_shade_source.md 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_trace_spec" -s 14 -c 3 -o /tmp/r2k_trace python bench.py --steps 16 --warmup 16 --device-only >> gpurun_out/ncu_k.log 2>&1
python tools/summarize_ncu.py source /tmp/r2k_trace.ncu-rep > gpurun_out/r2k_trace_source.md 2>&1
head -50 gpurun_out/r2k_shade_source.md | cut -c1-200
