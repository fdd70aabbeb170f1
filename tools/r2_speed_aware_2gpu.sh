#!/bin/bash
# Speed-aware shards on N GPUs: small C3 / C5 batches, with and without (gpurun --gpus N -- bash tools/r2_speed_aware_2gpu.sh N)
N=${1:-2}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
S=$((256 * N))
$RUN --master-port 29531 bench.py --gpus $N --sequences $S --steps 4 --warmup 4 --no-cpu > gpurun_out/r2_sa_c3_${N}gpu.json 2> gpurun_out/r2_sa_c3_${N}gpu.err
cut -c1-250 gpurun_out/r2_sa_c3_${N}gpu.json; tail -3 gpurun_out/r2_sa_c3_${N}gpu.err
$RUN --master-port 29532 bench.py --gpus $N --sequences $S --steps 4 --warmup 4 --no-cpu --no-e2e --no-records --no-speed-aware > gpurun_out/r2_sa_c3_${N}gpu_equal.json 2> gpurun_out/r2_sa_c3_${N}gpu_equal.err
cut -c1-250 gpurun_out/r2_sa_c3_${N}gpu_equal.json; tail -3 gpurun_out/r2_sa_c3_${N}gpu_equal.err
$RUN --master-port 29533 bench.py --gpus $N --config c5 --sequences $((64 * N)) --steps 3 --warmup 4 --no-cpu --no-e2e --no-records > gpurun_out/r2_sa_c5_${N}gpu.json 2> gpurun_out/r2_sa_c5_${N}gpu.err
cut -c1-250 gpurun_out/r2_sa_c5_${N}gpu.json; tail -3 gpurun_out/r2_sa_c5_${N}gpu.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_sa_*gpu*.json')):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, 'ERR', e); continue
    c = d['config']
    print(f, round(d['value']), round(d['ms_per_step'], 1), 'e2e', d['e2e'] and round(d['e2e']['value']), c.get('speed_aware_shards'), c['by_rank'], c['records'][:1])
PY
