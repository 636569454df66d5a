#!/bin/bash
# Builds A/B variants of the BM25 query kernel as alternate libraries with the same ABI:
#   leann_rs_b200/alt/libleann_cuda_<name>.so, selected at run time with LEANN_CUDA_LIB=<path> (benchmarks/k3_probe.py).
# Measured on the C5 corpus (10k queries, top-50), kernel ms: default geometry (8192-document tiles, 256 threads, 3 CTAs per SM)
# 28.8; 16384 / 256 / 2: 35.2; 16384 / 512 / 2: 31.8; 8192 / 512 / 2: 33.0 (profiles/r2_k3_variants.json).
set -e
cd "$(dirname "$0")/../leann_rs_b200"
make -j8 -s
mkdir -p alt build_alt
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-pthread --expt-relaxed-constexpr"
build() { # name defines...
  name=$1; shift
  nvcc $FLAGS "$@" -x cu -c csrc/bm25.cu -o build_alt/bm25_$name.o
  objs=$(ls build/*.o | grep -v "build/bm25.cu.o")
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o alt/libleann_cuda_$name.so $objs build_alt/bm25_$name.o -cudart shared -Xlinker -rpath,/usr/local/cuda/lib64 -lpthread -ldl
}
for v in "$@"; do
  case $v in
    nopipe) build nopipe -DLEANN_BM_PIPE=0 ;;
    t16k_256x2) build t16k_256x2 -DLEANN_BM_TILE=16384 -DLEANN_BM_MINB=2 -DLEANN_BM_PIPE=0 ;;
    t16k_512x2) build t16k_512x2 -DLEANN_BM_TILE=16384 -DLEANN_BM_THREADS=512 -DLEANN_BM_CAP=4096 -DLEANN_BM_MINB=2 -DLEANN_BM_PIPE=0 ;;
    t4k_256x4) build t4k_256x4 -DLEANN_BM_TILE=4096 -DLEANN_BM_MINB=4 ;;
    t8k_512x2) build t8k_512x2 -DLEANN_BM_THREADS=512 -DLEANN_BM_CAP=4096 -DLEANN_BM_MINB=2 -DLEANN_BM_PIPE=0 ;;
  esac
done
ls alt/
