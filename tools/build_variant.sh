#!/bin/bash
# A/B build of the C-ABI library with extra -D flags on ONE source, for RQB200_LIB=<path> runs.
#   tools/build_variant.sh p3 encode_tc3.cu -DT3_PREFETCH_N=3   →  ai_education_generative_recommendation_b200/librqvae_b200_p3.so
# Needs the regular build first (python -c 'import __graft_entry__ as g; g.build()'): the other objects are reused.
set -e
name=$1; src=$2; shift 2
root=$(cd "$(dirname "$0")/.." && pwd)
csrc=$root/ai_education_generative_recommendation_b200/csrc
obj=$csrc/build/${src%.cu}_$name.o
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --fmad=false -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
     "$@" -c "$csrc/$src" -o "$obj"
# link the regular objects (one per .cu) of the other sources plus this variant; objects of earlier variants are skipped
regular=""
for o in "$csrc"/build/*.o; do
    b=$(basename "$o" .o)
    [ -f "$csrc/$b.cu" ] && [ "$b.cu" != "$src" ] && regular="$regular $o"
done
out=$root/ai_education_generative_recommendation_b200/librqvae_b200_$name.so
nvcc -shared -o "$out" $regular "$obj" -lcudart -lcuda
echo "$out"
