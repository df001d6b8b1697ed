#!/bin/bash
# usage: scripts/probe_bench.sh "<make EXTRA flags>" [bench args...]   -- rebuild the library with flags, run the search bench
set -e
cd "$(dirname "$0")/.."
EXTRA="$1"; shift
make -C research_new_hnsw_b200/csrc clean > /dev/null
make -C research_new_hnsw_b200/csrc -j8 EXTRA="$EXTRA" > /dev/null
B200HNSW_BENCH_SKIP_BUILD=1 python bench.py --cpu-seconds 0.1 "$@" | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); r=j['roofline']
        print('EXTRA=[$EXTRA] ef=%d recall=%.4f value=%.3f MQPS e2e=%.3f MQPS kernel=%.4f ms achieved=%.0f GB/s frac=%.3f D=%.0f' % (j['config']['ef'], j['config']['recall_at_10'], j['value']/1e6, j['e2e']['value']/1e6, r['kernel_ms'], r['achieved'], r['frac'], r['per_query']['D']))
"
