#!/bin/bash
# A/B session of the shared-memory row ring (bulk async copies) of the short-row graph kernel on one box: the C4 shard index
# is built once (K1_INDEX_CACHE), then one process per setting of LEANN_CUDA_RING (0 = rows in registers, 1 = ring).
# (profiles/r2_k1_ring_ab.log was taken when the switch still selected 6 or 7 CTAs per SM.) Usage: benchmarks/k1_ring_ab.sh 1 0 1 0
export K1_INDEX_CACHE=/dev/shm/k1c4
cd "$(dirname "$0")/.."
for v in "$@"; do
  export LEANN_CUDA_RING=$v
  echo "== ring $v"
  python benchmarks/k1_tune.py --variants default --steps 8
done
