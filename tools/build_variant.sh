#!/bin/bash
# Build an experimental variant of the library next to the real one: tools/build_variant.sh NAME -DFLAG=... [file.cu ...]
# (objects of the files not named are reused from stereoanywhere_b200/build).  Use with SA_B200_LIB=<path>.
set -e
name=$1; shift
flags=(); files=()
for a in "$@"; do case "$a" in -D*) flags+=("$a");; *) files+=("$a");; esac; done
cd "$(dirname "$0")/../stereoanywhere_b200"
mkdir -p lib/variants build/variants/$name
objs=()
for o in build/*.o; do
  base=$(basename $o .o); skip=0
  for f in "${files[@]}"; do [ "$f" == "$base.cu" ] && skip=1; done
  [ $skip == 0 ] && objs+=("$o")
done
for f in "${files[@]}"; do
  base=$(basename $f .cu)
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr "${flags[@]}" -c csrc/$f -o build/variants/$name/$base.o
  objs+=("build/variants/$name/$base.o")
done
nvcc -shared -o lib/variants/libsa_b200_$name.so "${objs[@]}" -gencode arch=compute_100a,code=sm_100a -lcudart_static
echo lib/variants/libsa_b200_$name.so
