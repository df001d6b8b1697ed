#!/bin/bash
# usage: scripts/probe_variants.sh lib1.so lib2.so ...  -- same box, same graph, one bench line per library build
cd "$(dirname "$0")/.."
for lib in "$@"; do
  B200HNSW_LIB=$PWD/$lib B200HNSW_BENCH_SKIP_BUILD=1 python bench.py --cpu-seconds 0.1 --steps 40 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); r=j['roofline']
        print('$lib ef=%d value=%.3f MQPS e2e=%.3f MQPS kernel=%.4f ms frac=%.3f' % (j['config']['ef'], j['value']/1e6, j['e2e']['value']/1e6, r['kernel_ms'], r['frac']))
"
done
