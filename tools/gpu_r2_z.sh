#!/bin/bash
# round 2, capture Z: 256-bit node / leaf-box loads in the merged-mode kernels only; GPU tests
mkdir -p gpurun_out; rm -f gpurun_out/r2z_variants.jsonl
python tools/quick_bench.py --config atrium --spp 64 --tag final 2>>gpurun_out/r2z.err | tee -a gpurun_out/r2z_variants.jsonl
python tools/quick_bench.py --config instanced --spp 16 --tag final 2>>gpurun_out/r2z.err | tee -a gpurun_out/r2z_variants.jsonl
python tools/quick_bench.py --config atrium --accel two_level --spp 32 --tag final_2l 2>>gpurun_out/r2z.err | tee -a gpurun_out/r2z_variants.jsonl
python tools/quick_bench.py --config mixed --spp 8 --tag final_mixed 2>>gpurun_out/r2z.err | tee -a gpurun_out/r2z_variants.jsonl
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
