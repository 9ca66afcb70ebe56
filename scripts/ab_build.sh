#!/bin/bash
# A/B helper: build a second copy of the library with extra nvcc defines into ab/<name>.so
#   scripts/ab_build.sh nospill -DSSS_AB_NO_SPILL
# and run it on the GPU box by copying it over sessionsimilaritysearch_b200/libsss_b200.so inside the gpurun command.
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $root/ab/obj_$name
objs=""
for s in $root/sessionsimilaritysearch_b200/csrc/*.cu; do
  o=$root/ab/obj_$name/$(basename ${s%.cu}).o
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c $s -o $o &
  objs="$objs $o"
done
wait
/usr/local/cuda/bin/nvcc -shared -o $root/ab/$name.so $objs -gencode arch=compute_100a,code=sm_100a -cudart static
rm -rf $root/ab/obj_$name
echo $root/ab/$name.so
