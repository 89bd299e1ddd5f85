#!/bin/bash
# round 2, capture F (FINAL kernels of the round: + predicated wide pushes, 256-bit node loads, streaming ray accesses, packet traversal of the camera rays):
# GPU tests, ncu launch list + --set full of one wave on configs[1], --set full of the traversal launches on configs[3] (instanced),
# per-ray counters for bench.py's roofline, then the bench lines (driver command, 128 steps, instanced, reference arm)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2f_gpu_tests.log; cat gpurun_out/r2f_gpu_tests.log
python bench.py --steps 16 --warmup 16 --device-only > gpurun_out/r2f_device_only_s16.json 2> gpurun_out/r2f_device_only.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 16 --warmup 16 --device-only > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_trace_spec|k_trace_packet|k_shade" -s 21 -c 21 -o /tmp/r2f_kernels python bench.py --steps 16 --warmup 16 --device-only > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/r2f_kernels.ncu-rep --page raw --csv > gpurun_out/r2f_raw.csv 2> gpurun_out/ncu_export.err
ncu --set full --clock-control none --import-source on -k regex:"k_trace_spec|k_trace_packet" -s 14 -c 3 -o /tmp/r2f_trace python bench.py --steps 16 --warmup 16 --device-only >> gpurun_out/ncu_full.log 2>&1
python tools/summarize_ncu.py source /tmp/r2f_trace.ncu-rep > gpurun_out/r2f_trace_source.md 2>&1
python bench.py --config instanced --steps 8 --warmup 8 --device-only > gpurun_out/r2f_inst_device_only_s8.json 2> gpurun_out/r2f_inst_device_only.err
ncu --set full --clock-control none -k regex:"k_trace_spec" -s 14 -c 14 -o /tmp/r2f_inst python bench.py --config instanced --steps 8 --warmup 8 --device-only > gpurun_out/ncu_inst.log 2>&1
ncu -i /tmp/r2f_inst.ncu-rep --page raw --csv > gpurun_out/r2f_inst_raw.csv 2>> gpurun_out/ncu_export.err
python tools/summarize_ncu.py counters gpurun_out/r2f_raw.csv gpurun_out/r2f_device_only_s16.json profiles/r2_extend_counters_atrium.json
python tools/summarize_ncu.py counters gpurun_out/r2f_inst_raw.csv gpurun_out/r2f_inst_device_only_s8.json profiles/r2_extend_counters_instanced.json
cp profiles/r2_extend_counters_atrium.json profiles/r2_extend_counters_instanced.json gpurun_out/
python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench_s20.json 2> gpurun_out/r2f_bench_s20.err
python bench.py > gpurun_out/r2f_bench_default.json 2> gpurun_out/r2f_bench_default.err
python bench.py --config instanced --steps 16 --warmup 8 > gpurun_out/r2f_bench_instanced.json 2> gpurun_out/r2f_bench_instanced.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err
python __graft_entry__.py smoke > gpurun_out/r2f_smoke.log 2>&1 || python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1
tail -2 gpurun_out/r2f_smoke.log
for f in r2f_bench_s20 r2f_bench_default r2f_bench_instanced; do python -c "
import json,sys
d=json.load(open('gpurun_out/$f.json')); r=d['roofline']
print('$f', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'frac', round(r['frac'],3), r['bound'], 'lanes', round(r['ceilings']['issue']['active_threads_per_warp'],2), 'inst/ray', round(r['ceilings']['issue']['warp_inst_per_ray'],1), 'hbm', round(r['ceilings']['hbm']['frac'],3))
"; done
