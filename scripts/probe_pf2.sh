#!/bin/bash
# second A/B round: early prefetch, team sizes, e2e chunking (same box, same graph)
cd "$(dirname "$0")/.."
python scripts/probe_pf.py 28 > /dev/null 2>&1
for v in "PF=11" "PF=67" "PF=75" "PF=64" "PF=65"; do
  env B200HNSW_$v python scripts/probe_pf.py 28 64 128 2>&1 | grep "ms/step"
done
for t in 32 128; do
  env B200HNSW_PF=11 B200HNSW_TEAM=$t python scripts/probe_pf.py 28 2>&1 | grep "ms/step"
  env B200HNSW_PF=67 B200HNSW_TEAM=$t python scripts/probe_pf.py 28 2>&1 | grep "ms/step"
done
for hb in 10 12; do
  env B200HNSW_PF=11 B200HNSW_HASH_BITS=$hb python scripts/probe_pf.py 28 64 2>&1 | grep "ms/step"
done
env B200HNSW_PF=11 PROBE_BF16=1 python scripts/probe_pf.py 28 2>&1 | grep "ms/step"
env B200HNSW_PF=67 PROBE_BF16=1 python scripts/probe_pf.py 28 2>&1 | grep "ms/step"
