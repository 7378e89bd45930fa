#!/bin/bash
# tools/build_variant.sh NAME "-DFLAG=.. ..."  -> build/variants/libsqmp_NAME.so  (kernel A/B builds; dev aid)
set -e
name=$1; shift
cd "$(dirname "$0")/../squirrel_motion_planner_b200/csrc"
out=../../build/variants; mkdir -p $out/obj_$name
for f in capi validity validity_fast fk nn; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -ccbin /usr/bin/g++ -Xcompiler -fPIC,-O2 -Xptxas -v --expt-relaxed-constexpr "$@" -c $f.cu -o $out/obj_$name/$f.o 2> $out/obj_$name/$f.ptxas.log &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -shared -o $out/libsqmp_$name.so $out/obj_$name/*.o -lcudart
grep -A2 "k1_fastILb0ELb0" $out/obj_$name/validity_fast.ptxas.log | grep -E "spill|registers" | head -3
