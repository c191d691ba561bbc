#!/bin/bash
# build an experimental variant of the library: tools/build_variant.sh <name> <extra nvcc flags...>
# -> partgs_b200/variants/lib<name>.so  (select with PARTGS_B200_LIB=...)
name=$1; shift
d=partgs_b200/variants; mkdir -p $d/obj_$name
for f in partgs_b200/csrc/*.cu; do
  b=$(basename $f .cu)
  if [ "$b" == "render" ] || [ ! -f partgs_b200/csrc/build/$b.o ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC -Xptxas -v "$@" -c $f -o $d/obj_$name/$b.o 2> $d/obj_$name/$b.log || { cat $d/obj_$name/$b.log; exit 1; }
  else
    cp partgs_b200/csrc/build/$b.o $d/obj_$name/$b.o
  fi
done
nvcc -shared -o $d/lib$name.so $d/obj_$name/*.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -lcudart
grep -A2 "render_fwd_kernelILb0\|render_bwd_kernelILb0" $d/obj_$name/render.log | grep -E "Used|spill" 
