#!/bin/bash
# A/B inside one box: build kernels with 128-thread (head) vs 64-thread CTAs x visited-table size (B200HNSW_BUILD_HASH)
cd "$(dirname "$0")/.."
export B200HNSW_BUILD_PROFILE=1
for h in 4096 2048; do
  echo "== head (128 threads) hash $h"; B200HNSW_BUILD_HASH=$h python scripts/probe_build_only.py 2>&1 | grep -E "rep|profile"
done
for h in 4096 2048 1024; do
  echo "== 64 threads hash $h"; B200HNSW_LIB=$PWD/research_new_hnsw_b200/_variants/libb200hnsw_bt64.so B200HNSW_BUILD_HASH=$h python scripts/probe_build_only.py 2>&1 | grep -E "rep|profile"
done
