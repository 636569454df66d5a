#!/bin/bash
# Builds A/B variants of the graph-search kernels as alternate libraries with the same ABI:
#   leann_rs_b200/alt/libleann_cuda_k1_<name>.so, selected at run time with LEANN_CUDA_LIB=<path>.
# Usage: benchmarks/k1_variants.sh l2pol1 l2pol3 ...   (l2pol<N>: -DLEANN_K1_L2POL=N, spec<N>: -DLEANN_K1_SPEC=N; see graph_device.cuh)
set -e
cd "$(dirname "$0")/../leann_rs_b200"
make -j8 -s
mkdir -p alt build_alt
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-pthread --expt-relaxed-constexpr -I../include"
build() { # name defines...
  name=$1; shift
  nvcc $FLAGS "$@" -x cu -c csrc/graph_search.cu -o build_alt/graph_search_$name.o
  objs=$(ls build/*.o | grep -v "build/graph_search.cu.o")
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o alt/libleann_cuda_k1_$name.so $objs build_alt/graph_search_$name.o -cudart shared -Xlinker -rpath,/usr/local/cuda/lib64 -lpthread -ldl
}
pids=""
for v in "$@"; do
  case $v in
    l2pol*) build $v -DLEANN_K1_L2POL=${v#l2pol} & pids="$pids $!" ;;
    spec*) build $v -DLEANN_K1_SPEC=${v#spec} & pids="$pids $!" ;;
  esac
done
for p in $pids; do wait $p; done
ls alt/
