#!/bin/bash
# round 2, capture AO: automatic choice of the packet kernels (instanced triangles <= 2 x pixels): parity + the three scenes
mkdir -p gpurun_out; rm -f gpurun_out/r2ao_variants.jsonl
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
python tools/quick_bench.py --config atrium --spp 32 --tag auto 2>>gpurun_out/r2ao.err | tee -a gpurun_out/r2ao_variants.jsonl
python tools/quick_bench.py --config atrium --accel two_level --spp 32 --tag auto_2l 2>>gpurun_out/r2ao.err | tee -a gpurun_out/r2ao_variants.jsonl
python tools/quick_bench.py --config instanced --spp 8 --tag auto_inst 2>>gpurun_out/r2ao.err | tee -a gpurun_out/r2ao_variants.jsonl
