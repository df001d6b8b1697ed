#!/bin/bash
# one gpurun call: A/B of the search kernel's prefetch / PDL policies on the same box and graph
cd "$(dirname "$0")/.."
python scripts/probe_pf.py 28 > /dev/null 2>&1   # build + cache the graph
for v in "PF=0 PDL=0" "PF=0 PDL=1" "PF=1 PDL=1" "PF=3 PDL=1" "PF=7 PDL=1" "PF=11 PDL=1" "PF=19 PDL=1" "PF=51 PDL=1" "PF=23 PDL=1" "PF=55 PDL=1" "PF=3 PDL=0"; do
  set -- $v
  env B200HNSW_$1 B200HNSW_$2 python scripts/probe_pf.py 28 64 2>&1 | grep "ms/step"
done
for v in "PF=0" "PF=3" "PF=19" "PF=55"; do
  env B200HNSW_$v PROBE_BF16=1 python scripts/probe_pf.py 28 2>&1 | grep "ms/step"
done
env B200HNSW_PF=3 B200HNSW_PDL=1 PROBE_NULLSTREAM=1 python scripts/probe_pf.py 28 2>&1 | grep "ms/step"
