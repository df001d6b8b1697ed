#!/bin/bash
cd "$(dirname "$0")/.."
export B200HNSW_BENCH_SKIP_BUILD=1
for c in "$@"; do
  B200HNSW_CHUNKS=$c python bench.py --cpu-seconds 0.1 --steps 60 2>/dev/null | C=$c python -c "
import json,sys,os
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('chunks %s: e2e %.3f MQPS (%.3f ms/step) device %.3f MQPS' % (os.environ['C'], j['e2e']['value']/1e6, j['e2e']['ms_per_step'], j['value']/1e6))
"
done
