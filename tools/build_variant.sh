#!/bin/bash
# tools/build_variant.sh NAME [-DKNOB=VALUE ...] — another build of libbronko_b200.so with different compile-time
# knobs, for A/B runs on the GPU box:  BRONKO_B200_LIB=bronko_b200/csrc/variants/NAME.so python bench.py --no-e2e ...
set -e
cd "$(dirname "$0")/../bronko_b200/csrc"
name=$1; shift
mkdir -p variants
make -s bk_host_index.o bk_io.o
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-Wall,-Wno-unused-function \
     -Xptxas -v "$@" -c -o variants/$name.o bk_device.cu 2> variants/$name.ptxas.log || (cat variants/$name.ptxas.log; false)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/$name.so variants/$name.o bk_host_index.o bk_io.o -lz -ldl -cudart static
rm -f variants/$name.o
echo "built variants/$name.so"
