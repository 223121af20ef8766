#!/bin/bash
# dev: build variants/libptb200_<name>.so with extra nvcc flags applied to every translation unit
# usage: scripts/build_variant.sh mb9 "-DPT_TRACE_MINBLOCKS=9" [ptxas-info]
set -e
NAME=$1; EXTRA=$2
cd "$(dirname "$0")/../optix-path-tracer_b200/csrc"
mkdir -p ../../variants/obj_$NAME
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -Xcompiler -w --expt-relaxed-constexpr"
for f in api bvh_build env_alias integrator multi; do
  $NV $EXTRA ${3:+-Xptxas -v} -c $f.cu -o ../../variants/obj_$NAME/$f.o 2> ../../variants/obj_$NAME/$f.log &
done
wait
$NV -shared -o ../../variants/libptb200_$NAME.so ../../variants/obj_$NAME/*.o -ldl
[ -n "$3" ] && grep -E "Compiling entry|registers" ../../variants/obj_$NAME/integrator.log | paste - - | sed -E 's/.*(k_trace_rays|k_trace|k_shadow|k_shade|k_miss|k_generate)(I[A-Za-z0-9]*E)?.*Used ([0-9]+) registers(.*)/\1\2 \3 \4/' | grep -E "k_trace|k_shadow" | sort -u
echo "built variants/libptb200_$NAME.so"
