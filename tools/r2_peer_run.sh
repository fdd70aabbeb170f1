#!/bin/bash
# Peer-mapped dW exchange on N GPUs: check + timing against NCCL, then the bench step with it on / off
# (gpurun --gpus N -- bash tools/r2_peer_run.sh N [bench: 1|0] [extra bench flags; default --no-speed-aware])
N=${1:-2}
BENCH=${2:-1}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
GRPO_PEER_LOG=gpurun_out/r2_peer_check_${N}gpu.log timeout 170 $RUN --master-port 29541 tools/gpu_peer_check.py --iters 6 2> gpurun_out/r2_peer_check_${N}gpu.err | grep -v "^\[rank [1-9]" 
tail -3 gpurun_out/r2_peer_check_${N}gpu.err
[ "$BENCH" = "1" ] || exit 0
S=$((512 * N))   # 512 sequences per rank = the per-rank share of config C3 on 8 GPUs
EXTRA=${3:---no-speed-aware}
for MODE in on off; do
  timeout 200 $RUN --master-port 2955$N bench.py --gpus $N --sequences $S --steps 4 --warmup 3 --no-cpu --no-e2e --no-records $EXTRA --peer $MODE \
    > gpurun_out/r2_peer_bench_${N}gpu_$MODE.json 2> gpurun_out/r2_peer_bench_${N}gpu_$MODE.err
  tail -2 gpurun_out/r2_peer_bench_${N}gpu_$MODE.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_peer_bench_*gpu_*.json')):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, 'ERR', e); continue
    c = d['config']
    print(f, round(d['value']), 'tok/s', round(d['ms_per_step'], 1), 'ms/step', c['dw_exchange'][:40], c['by_rank'], 'outside timers', round(d['roofline']['ms_per_step_outside_phase_timers'], 1))
PY
