#!/bin/bash
# C4: sampled bound pass stride x candidate-buffer depth (one process per setting, same box)
cd "$(dirname "$0")/.."
export B200HNSW_BF_PROFILE=1
for sc in "4 1536" "8 3072" "8 2048" "16 6144" "16 4096"; do
  set -- $sc
  echo "== stride $1 cap $2"
  B200HNSW_BF_SAMPLE=$1 B200HNSW_BF_CAP=$2 python scripts/probe_bf_c4.py 10000 2>&1 | grep -E "nq|tensor ==|profile|cand" | tail -4
done
