#!/bin/bash
# tools/build_variant.sh NAME [-DFLAG ...]  → libflate_b200/libb2f_NAME.so (a second build of the library for A/B kernel experiments;
# select it with B2F_LIB=libflate_b200/libb2f_NAME.so)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
tmp=/tmp/b2f_variant_$name; rm -rf $tmp; mkdir -p $tmp
rm -f libflate_b200/libb2f_$name.so
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -cudart static"
objs=""; pids=""
for s in b2f_api encode_kernels decode_kernels spec_kernels checksum_kernels; do
  /usr/local/cuda/bin/nvcc $F "$@" -c libflate_b200/csrc/$s.cu -o $tmp/$s.o &
  pids="$pids $!"
  objs="$objs $tmp/$s.o"
done
for p in $pids; do wait $p; done
/usr/local/cuda/bin/nvcc $F -shared -o libflate_b200/libb2f_$name.so $objs
echo libflate_b200/libb2f_$name.so
