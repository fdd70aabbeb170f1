#!/bin/bash
# The bench exactly as the driver launches it for N > 1 (records, e2e, speed-aware shards, peer exchange on), at the per-rank
# batch of the 8-GPU run:   gpurun --gpus 2 -- bash tools/r2_default_2gpu.sh
N=${1:-2}
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --sequences $((512 * N)) --steps 4 --warmup 3 --no-cpu \
  > gpurun_out/r2_bench_default_${N}gpu.json 2> gpurun_out/r2_bench_default_${N}gpu.err
echo "rc=$?"; tail -3 gpurun_out/r2_bench_default_${N}gpu.err
python - <<PY
import json
d = json.load(open('gpurun_out/r2_bench_default_${N}gpu.json'))
c = d['config']
print(round(d['value']), 'tok/s', round(d['ms_per_step'], 1), 'ms/step | e2e', d['e2e'] and round(d['e2e']['value']), '| records', [(r['micro_batch_sequences'], round(r['value'])) for r in c['records']], '|', c['dw_exchange'][:60], c['speed_aware_shards'], c['by_rank'])
PY
