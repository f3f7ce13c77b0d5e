#!/bin/bash
# build tuning variants of K1 (decim1.cu) into habdec_b200/variant_<name>.so; the other objects are shared
# usage: tools/build_variants.sh "name:-DHBD_K1_STAGES=4 -DHBD_K1_PSB_MIN=6" ...
set -e
cd "$(dirname "$0")/../habdec_b200/csrc"
make -j8 >/dev/null
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
for spec in "$@"; do
  name=${spec%%:*}; defs=${spec#*:}
  mkdir -p build/var_$name
  $NVCC $ARCH -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $defs -c decim1.cu -o build/var_$name/decim1.o
  extra=""
  for f in tail api; do   # these see decim1.cuh / tail tunables too
    $NVCC $ARCH -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $defs -c $f.cu -o build/var_$name/$f.o
  done
  objs=$(ls build/*.o | grep -v -e build/decim1.o -e build/tail.o -e build/api.o)
  $NVCC $ARCH -shared -o ../variant_$name.so build/var_$name/decim1.o build/var_$name/tail.o build/var_$name/api.o $objs
  echo "built variant_$name.so ($defs)"
done
