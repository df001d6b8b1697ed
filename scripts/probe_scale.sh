#!/bin/bash
# A/B of the N>1 exchange inside one gpurun call.  usage: probe_scale.sh N mode...   (modes: pipe serial ctas2 ll)
N=${1:-2}; shift
for mode in ${@:-pipe serial}; do
  unset B200HNSW_BENCH_INFLIGHT B200HNSW_BENCH_NO_PIPELINE NCCL_MAX_CTAS NCCL_MIN_CTAS NCCL_PROTO TORCH_NCCL_HIGH_PRIORITY B200HNSW_PIPE_DEPTH
  case $mode in
    d4) export B200HNSW_PIPE_DEPTH=4;;
    hp_d4) export B200HNSW_PIPE_DEPTH=4 TORCH_NCCL_HIGH_PRIORITY=1;;
    hp_d4_ctas2) export B200HNSW_PIPE_DEPTH=4 TORCH_NCCL_HIGH_PRIORITY=1 NCCL_MAX_CTAS=2 NCCL_MIN_CTAS=1;;
    hp) export TORCH_NCCL_HIGH_PRIORITY=1;;
    serial) export B200HNSW_BENCH_NO_PIPELINE=1;;
    unbounded) export B200HNSW_BENCH_INFLIGHT=0;;
    inflight3) export B200HNSW_BENCH_INFLIGHT=3;;
    ctas2) export NCCL_MAX_CTAS=2 NCCL_MIN_CTAS=1;;
    ll) export NCCL_PROTO=LL;;
  esac
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
    bench.py --gpus $N --steps 50 --warmup 5 2> gpurun_out/scale_${mode}_n$N.err | grep '^{' > gpurun_out/scale_${mode}_n$N.json
  python - <<PY
import json
d=json.loads(open('gpurun_out/scale_${mode}_n$N.json').read())
print('$mode N=$N value %.2fM ms/step %.3f kernel_ms %.3f ef %d e2e %.2fM'%(d['value']/1e6,d['ms_per_step'],d['roofline']['kernel_ms'],d['config']['ef'],d['e2e']['value']/1e6), d.get('per_rank'))
PY
done
