#!/bin/bash
# A/B session of the shared-memory visited tables (smv) of the short-row graph kernel on one box: the C4 shard index is built
# once (K1_INDEX_CACHE), then one process per setting. Settings are environment switches of the default library:
#   hyb     LEANN_CUDA_SMV=2                      shared-memory first level + q16 overflow level (default)
#   smv     LEANN_CUDA_SMV=1                      stand-alone shared-memory tables, 3 CTAs per SM, unroll 3
#           (profiles/r2_k1_smv_ab.log was taken when LEANN_CUDA_SMV_U still selected unroll 2 / 3 / 4)
#   q16     LEANN_CUDA_SMV=0                      the L2-resident q16 tables alone
# Usage: benchmarks/k1_smv_ab.sh hyb q16 smv hyb
export K1_INDEX_CACHE=/dev/shm/k1c4
cd "$(dirname "$0")/.."
for v in "$@"; do
  unset LEANN_CUDA_SMV
  case $v in
    hyb) export LEANN_CUDA_SMV=2 ;;
    smv*) export LEANN_CUDA_SMV=1 ;;
    q16) export LEANN_CUDA_SMV=0 ;;
  esac
  echo "== $v"
  python benchmarks/k1_tune.py --variants default --steps 8
done
