#!/bin/bash
# Builds an experimental libbpt variant: tools/build_variant.sh <tag> <extra nvcc flags...>  ->  bisemutum-engine_b200/csrc/_exp/libbpt_<tag>.so
set -e
cd "$(dirname "$0")/../bisemutum-engine_b200/csrc"
tag=$1; shift
mkdir -p _exp
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-fvisibility=hidden "$@" \
    -shared -o _exp/libbpt_$tag.so bpt_api.cu bvh_build.cu render.cu post.cu ibl.cu lighttex.cu reblur.cu -cudart static
