#!/bin/bash
# Kernel experiments: builds libsparsex_b200.so variants with extra -D flags into sparsex_b200/variants/ (git-ignored,
# travels to the GPU box); select one at run time with SPARSEX_B200_LIB=<path>.
# Usage: tools/variants.sh name1="-DFOO=1" name2="-DFOO=2 -DBAR" ...
set -e
cd "$(dirname "$0")/../sparsex_b200/csrc"
make -s -j8 > /dev/null
mkdir -p ../variants
for spec in "$@"; do
  name="${spec%%=*}"; flags="${spec#*=}"
  (
    /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v $flags \
      -c engine.cu -o ../variants/engine_$name.o 2> ../variants/ptxas_$name.log
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/lib_$name.so encoder.o mmf.o gpu_layout.o \
      csx_tools.o rcm.o ../variants/engine_$name.o api.o -Xcompiler -pthread
    rm -f ../variants/engine_$name.o
    echo "built $name ($flags)"
  ) &
done
wait
