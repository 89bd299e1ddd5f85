#!/bin/bash
# round 2, call G: ncu of the current kernels on configs[1] (launch list + --set full of one wave) and on configs[3] (instanced, --set full of the
# traversal launches of one wave), bench lines of both configs
mkdir -p gpurun_out
python bench.py --steps 16 --warmup 16 --device-only > gpurun_out/r2g_device_only_s16.json 2> gpurun_out/r2g_device_only.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g_launches.csv python bench.py --steps 16 --warmup 16 --device-only > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_trace_spec|k_shade" -s 21 -c 21 -o /tmp/r2g_kernels python bench.py --steps 16 --warmup 16 --device-only > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/r2g_kernels.ncu-rep --page raw --csv > gpurun_out/r2g_raw.csv 2> gpurun_out/ncu_export.err
# instanced: 8 samples per wave at 4K (2^26 / pixels); warm-up 8 = one wave, the timed 8 = the second wave: 14 traversal launches each
python bench.py --config instanced --steps 8 --warmup 8 --device-only > gpurun_out/r2g_inst_device_only_s8.json 2> gpurun_out/r2g_inst_device_only.err
ncu --set full --clock-control none -k regex:"k_trace_spec" -s 14 -c 14 -o /tmp/r2g_inst python bench.py --config instanced --steps 8 --warmup 8 --device-only > gpurun_out/ncu_inst.log 2>&1
ncu -i /tmp/r2g_inst.ncu-rep --page raw --csv > gpurun_out/r2g_inst_raw.csv 2>> gpurun_out/ncu_export.err
python bench.py --config instanced --steps 16 --warmup 8 > gpurun_out/r2g_bench_instanced.json 2> gpurun_out/r2g_bench_instanced.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench_s20.json 2> gpurun_out/r2g_bench_s20.err
ls -la /tmp/*.ncu-rep; head -c 400 gpurun_out/r2g_bench_instanced.json; echo; tail -3 gpurun_out/r2g_bench_instanced.err
