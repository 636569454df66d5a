#!/bin/bash
# A/B session of the short-row graph kernel on one box: the C4 shard index is built once (K1_INDEX_CACHE), then one process
# per alternate library (benchmarks/k1_variants.sh). Usage: benchmarks/k1_ab.sh - l2pol1 l2pol3 ...   ("-" = the default library)
export K1_INDEX_CACHE=/dev/shm/k1c4
cd "$(dirname "$0")/.."
for v in "$@"; do
  if [ "$v" = "-" ]; then unset LEANN_CUDA_LIB; else export LEANN_CUDA_LIB=$PWD/leann_rs_b200/alt/libleann_cuda_k1_$v.so; fi
  python benchmarks/k1_tune.py --variants default --steps 8
done
