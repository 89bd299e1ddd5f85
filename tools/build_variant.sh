#!/bin/bash
# Builds an experimental libbpt variant: tools/build_variant.sh <tag> <extra nvcc flags...>  ->  bisemutum-engine_b200/csrc/_exp/libbpt_<tag>.so
# Only render.cu (the traversal / shade kernels, where the tunables live) is recompiled; the other objects come from the product build (make).
set -e
cd "$(dirname "$0")/../bisemutum-engine_b200/csrc"
tag=$1; shift
mkdir -p _exp
make -s >/dev/null
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-fvisibility=hidden "$@" \
    -c render.cu -o _exp/render_$tag.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o _exp/libbpt_$tag.so _exp/render_$tag.o \
    _obj/bpt_api.o _obj/bvh_build.o _obj/post.o _obj/ibl.o _obj/lighttex.o _obj/reblur.o -cudart static
rm -f _exp/render_$tag.o
