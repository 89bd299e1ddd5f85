#!/bin/bash
# round 2 final, 1 GPU: tests, smoke, the driver's bench command + reference arm, steady-state line, instanced line, sanitizer
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_final_gpu_tests.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_bench_reference.json 2> gpurun_out/r2_final_ref.err
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_bench_s20.json 2> gpurun_out/r2_final_bench_s20.err
python bench.py > gpurun_out/r2_final_bench_default.json 2> gpurun_out/r2_final_bench_default.err
python bench.py --config instanced --steps 16 --warmup 8 > gpurun_out/r2_final_bench_instanced.json 2> gpurun_out/r2_final_bench_instanced.err
python tools/bench_configs.py 0_ 1_ 2_ 4_ > gpurun_out/r2_final_configs.log 2>&1; cp gpurun_out/configs.json gpurun_out/r2_final_configs.json
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_smoke.py > gpurun_out/r2_final_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2_final_sanitizer_memcheck.log
tail -3 gpurun_out/r2_final_gpu_tests.log; cat gpurun_out/r2_final_smoke.log | tail -2; tail -4 gpurun_out/r2_final_sanitizer_memcheck.log
head -c 300 gpurun_out/r2_final_bench_s20.json; echo
