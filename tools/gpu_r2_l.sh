#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r2l_variants.jsonl
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/quick_bench.py --config instanced --spp 16 --tag park 2>>gpurun_out/r2l.err | tee -a gpurun_out/r2l_variants.jsonl
python tools/quick_bench.py --config instanced --spp 16 --tag nopark --lib bisemutum-engine_b200/csrc/_exp/libbpt_nopark.so 2>>gpurun_out/r2l.err | tee -a gpurun_out/r2l_variants.jsonl
python tools/quick_bench.py --accel two_level --spp 32 --tag atrium2l_park 2>>gpurun_out/r2l.err | tee -a gpurun_out/r2l_variants.jsonl
python tools/quick_bench.py --accel two_level --spp 32 --tag atrium2l_nopark --lib bisemutum-engine_b200/csrc/_exp/libbpt_nopark.so 2>>gpurun_out/r2l.err | tee -a gpurun_out/r2l_variants.jsonl
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_smoke.py > gpurun_out/r2_final_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2_final_sanitizer_memcheck.log
tail -4 gpurun_out/r2_final_sanitizer_memcheck.log
