#!/bin/bash
# Builds libpcd_b200.so variants that differ only in the macros given for csrc/sor_resident.cu (A/B timing on the GPU box:
# PCD_LIB=poisson_caustic_design_b200/variants/libpcd_<name>.so python tools/res_time.py ...).
#   tools/build_variants.sh name1:"-DX=0 -DY=1" name2:"..."
set -e
cd "$(dirname "$0")/../poisson_caustic_design_b200"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -I ../include -I csrc"
mkdir -p variants/obj
for f in csrc/*.cu; do
  b=$(basename $f .cu)
  [ $b = sor_resident ] && continue
  stale=0
  for d in $f csrc/*.h csrc/*.cuh ../include/pcd.h; do [ $d -nt variants/obj/$b.o ] && stale=1; done
  if [ ! -f variants/obj/$b.o ] || [ $stale = 1 ]; then nvcc $FLAGS -c $f -o variants/obj/$b.o & fi
done
wait
for spec in "$@"; do
  name=${spec%%:*}; defs=${spec#*:}
  ( nvcc $FLAGS $defs -c csrc/sor_resident.cu -o variants/obj/sor_resident_$name.o &&
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o variants/libpcd_$name.so variants/obj/sor_resident_$name.o $(ls variants/obj/*.o | grep -v sor_resident_) ) &
done
wait
ls -la variants/*.so
