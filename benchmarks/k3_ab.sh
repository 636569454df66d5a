#!/bin/bash
# A/B session of the BM25 query kernel on one box: the corpus is generated once (K3_CORPUS_CACHE), then one process per
# (library, dense-row threshold) pair. Usage: benchmarks/k3_ab.sh "<lib or -> <frac> [<max>]" ...
export K3_CORPUS_CACHE=/dev/shm/k3corpus.npz
cd "$(dirname "$0")/.."
for v in "$@"; do
  set -- $v
  lib=$1; frac=$2; mx=${3:-64}
  if [ "$lib" = "-" ]; then unset LEANN_CUDA_LIB; else export LEANN_CUDA_LIB=$PWD/leann_rs_b200/alt/libleann_cuda_$lib.so; fi
  LEANN_CUDA_BM25_DENSE_FRAC=$frac LEANN_CUDA_BM25_DENSE_MAX=$mx python benchmarks/k3_probe.py
done
