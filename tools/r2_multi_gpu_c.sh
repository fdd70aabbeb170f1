#!/bin/bash
# Round 2, third 8-GPU pass: speed-aware shards (default) against equal shards on the same box, C3 and C5
N=${1:-8}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$RUN --master-port 29541 bench.py --gpus $N --steps 8 --warmup 5 --no-cpu > gpurun_out/r2_bench_c3_${N}gpu_v3.json 2> gpurun_out/r2_bench_c3_${N}gpu_v3.err
cut -c1-200 gpurun_out/r2_bench_c3_${N}gpu_v3.json; tail -2 gpurun_out/r2_bench_c3_${N}gpu_v3.err
$RUN --master-port 29542 bench.py --gpus $N --steps 6 --warmup 3 --no-cpu --no-e2e --no-records --no-speed-aware > gpurun_out/r2_bench_c3_${N}gpu_v3_equal.json 2> gpurun_out/r2_bench_c3_${N}gpu_v3_equal.err
cut -c1-200 gpurun_out/r2_bench_c3_${N}gpu_v3_equal.json; tail -2 gpurun_out/r2_bench_c3_${N}gpu_v3_equal.err
$RUN --master-port 29543 bench.py --gpus $N --config c5 --steps 3 --warmup 4 --no-e2e --no-cpu --no-records > gpurun_out/r2_bench_c5_${N}gpu_v3.json 2> gpurun_out/r2_bench_c5_${N}gpu_v3.err
cut -c1-200 gpurun_out/r2_bench_c5_${N}gpu_v3.json; tail -2 gpurun_out/r2_bench_c5_${N}gpu_v3.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_bench_c*gpu_v3*.json')):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, 'ERR', e); continue
    c = d['config']
    print(f, round(d['value']), round(d['ms_per_step'], 1), 'e2e', d['e2e'] and round(d['e2e']['value']), c.get('speed_aware_shards'), c['by_rank'], c['records'][:1])
PY
